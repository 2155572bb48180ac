// sg_augment.cu — training-time augmentation of the input pipeline on a device-resident uint8 image pool
// (SURVEY.md §8f-1). Reference: src/data_loader_signatures.py:154-219 `get_train_transforms`, applied per image by
// PIL workers (`SignatureDataset.__getitem__`, :120-135):
//     RandomRotation(±5°, fill=255) -> RandomAffine(degrees=0, scale=(0.9, 1.1), fill=255) [-> horizontal flip]
//     -> ToTensor -> Normalize(0.5, 0.5)
// Both resamplings are nearest-neighbour on 8-bit pixels, so the result is integer work and must be BIT-EXACT:
//   * rotation  = Pillow `affine_fixed`: 16.16 fixed point, source = ((a5 + y*a4 + x*a3) >> 16, (a2 + y*a1 + x*a0) >> 16);
//   * scaling   = Pillow `ImagingScaleAffine`: a double-precision coordinate advanced by repeated addition,
//                 COORD(v) = v < 0 ? -1 : (int)v, per row and per column;
//   * ToTensor / Normalize = ((float)v / 255 - 0.5) / 0.5 in float32 (a 256-entry table here).
// One CTA handles one image at a time: the 4-16 KB source tile is gathered from the pool into shared memory with 16-byte
// loads, rotated into a second tile, and the scaled / flipped / normalised result is written as float4 rows — 1 byte
// read and 4 bytes written per pixel, nothing else touches HBM. The grid is a multiple of the SM count and strides over
// the batch.
//
// sg_augment_params (host code, no GPU) turns the sampled (angle, scale) of every image into the coefficient tables,
// with the same libm calls and the same decimal rounding as Pillow / torchvision's Python code.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#include "sg_kernels.cuh"

namespace sg {
namespace {

constexpr int kAugThreads = 256;

template <int S>
__global__ void __launch_bounds__(kAugThreads) augment_kernel(const uint8_t* __restrict__ pool,
                                                              const int* __restrict__ index,
                                                              const int* __restrict__ rot,
                                                              const double* __restrict__ sc,
                                                              const uint8_t* __restrict__ flip, int batch,
                                                              float* __restrict__ out) {
    __shared__ __align__(16) uint8_t src[S * S];
    __shared__ __align__(16) uint8_t rotd[S * S];
    __shared__ float lut[256];
    __shared__ short xt[S], yt[S];
    const int tid = threadIdx.x;
    // ToTensor + Normalize of every possible 8-bit value, with IEEE division (no reciprocal shortcuts)
    lut[tid] = __fdiv_rn(__fsub_rn(__fdiv_rn(static_cast<float>(tid), 255.0f), 0.5f), 0.5f);
    for (int img = blockIdx.x; img < batch; img += gridDim.x) {
        const long long src_img = index ? index[img] : img;
        const uint4* g = reinterpret_cast<const uint4*>(pool + src_img * (S * S));
        for (int i = tid; i < S * S / 16; i += kAugThreads) reinterpret_cast<uint4*>(src)[i] = __ldg(g + i);
        // coordinate tables of the scaling step: entry k is the start value advanced k times by `o += a` in double,
        // exactly the sequence Pillow's serial loop goes through (thread k repeats the first k additions)
        if (tid < 2 * S) {
            const int k = tid % S, which = tid / S;  // 0: columns (a0, xo), 1: rows (a4, yo)
            const double a = sc[img * 4 + 2 * which];
            double o = sc[img * 4 + 2 * which + 1];
            for (int j = 0; j < k; ++j) o += a;
            int v = o < 0.0 ? -1 : static_cast<int>(o);
            if (v >= S) v = -1;
            (which ? yt : xt)[k] = static_cast<short>(v);
        }
        const int a0 = rot[img * 6 + 0], a1 = rot[img * 6 + 1], a2 = rot[img * 6 + 2];
        const int a3 = rot[img * 6 + 3], a4 = rot[img * 6 + 4], a5 = rot[img * 6 + 5];
        const bool fl = flip && flip[img];
        __syncthreads();
        for (int p = tid; p < S * S; p += kAugThreads) {
            const int y = p / S, x = p % S;
            const int xin = (a2 + y * a1 + x * a0) >> 16;
            const int yin = (a5 + y * a4 + x * a3) >> 16;
            const bool ok = xin >= 0 && xin < S && yin >= 0 && yin < S;
            rotd[p] = ok ? src[yin * S + xin] : static_cast<uint8_t>(255);
        }
        __syncthreads();
        float4* o4 = reinterpret_cast<float4*>(out + static_cast<long long>(img) * (S * S));
        for (int q = tid; q < S * S / 4; q += kAugThreads) {
            const int y = q / (S / 4), x = (q % (S / 4)) * 4;
            const int yi = yt[y];
            float v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int xi = xt[fl ? S - 1 - (x + j) : x + j];
                v[j] = lut[(yi >= 0 && xi >= 0) ? rotd[yi * S + xi] : 255];
            }
            o4[q] = make_float4(v[0], v[1], v[2], v[3]);
        }
        __syncthreads();  // src / rotd / tables are rewritten by the next image
    }
}

// Geometry.c: FIX(v) = FLOOR(v * 65536 + 0.5) with FLOOR(v) = v < 0 ? (int)floor(v) : (int)v
inline int fix16(double v) {
    const double t = v * 65536.0 + 0.5;
    return t < 0.0 ? static_cast<int>(std::floor(t)) : static_cast<int>(t);
}

// Python's round(x, 15) for |x| <= 1: correctly rounded decimal conversion and back (glibc printf / strtod are exact).
inline double round15(double x) {
    char buf[48];
    snprintf(buf, sizeof(buf), "%.15f", x);
    return strtod(buf, nullptr);
}

}  // namespace

int augment_batch(const uint8_t* pool, const int* index, const int* rot, const double* sc, const uint8_t* flip, int batch,
                  int size, float* out, cudaStream_t s) {
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (sms <= 0) sms = 148;
    }
    const int per_sm = size == 64 ? 8 : 4;  // 9 KB / 33 KB of shared memory per CTA
    int grid = sms * per_sm;
    if (grid > batch) grid = batch;
    note_launch();
    if (size == 64)
        augment_kernel<64><<<grid, kAugThreads, 0, s>>>(pool, index, rot, sc, flip, batch, out);
    else if (size == 128)
        augment_kernel<128><<<grid, kAugThreads, 0, s>>>(pool, index, rot, sc, flip, batch, out);
    else
        return -1;
    return 0;
}

void augment_params(const double* angles, const double* scales, int n, int size, int* rot, double* sc) {
    const double c0 = size / 2.0;  // Image.rotate: center = (w / 2, h / 2); F.affine: center = (w * 0.5, h * 0.5)
    const int one = 65536;
    for (int i = 0; i < n; ++i) {
        // ---- PIL.Image.rotate(angle, NEAREST, expand=0, center=None) -> Image.transform(AFFINE) -> affine_fixed
        double ang = std::fmod(angles ? angles[i] : 0.0, 360.0);
        if (ang < 0.0) ang += 360.0;  // Python's float % is non-negative for a positive modulus
        if (ang >= 360.0) ang = 0.0;  // (-tiny) % 360.0 rounds to 360.0 in Python too; 360 deg == identity matrix path
        int* r = rot + 6 * i;
        if (ang == 0.0) {  // fast paths of Image.rotate: exact copies / transposes
            r[0] = one; r[1] = 0; r[2] = one / 2; r[3] = 0; r[4] = one; r[5] = one / 2;
        } else if (ang == 180.0) {
            r[0] = -one; r[1] = 0; r[2] = size * one - one / 2; r[3] = 0; r[4] = -one; r[5] = size * one - one / 2;
        } else if (ang == 90.0) {
            r[0] = 0; r[1] = -one; r[2] = size * one - one / 2; r[3] = one; r[4] = 0; r[5] = one / 2;
        } else if (ang == 270.0) {
            r[0] = 0; r[1] = one; r[2] = one / 2; r[3] = -one; r[4] = 0; r[5] = size * one - one / 2;
        } else {
            const double a = -(ang * (M_PI / 180.0));  // -math.radians(angle)
            double m[6] = {round15(std::cos(a)), round15(std::sin(a)), 0.0, round15(-std::sin(a)), round15(std::cos(a)), 0.0};
            m[2] = m[0] * (-c0) + m[1] * (-c0) + m[2];
            m[5] = m[3] * (-c0) + m[4] * (-c0) + m[5];
            m[2] += c0;
            m[5] += c0;
            r[0] = fix16(m[0]); r[1] = fix16(m[1]); r[2] = fix16(m[2] + m[0] * 0.5 + m[1] * 0.5);
            r[3] = fix16(m[3]); r[4] = fix16(m[4]); r[5] = fix16(m[5] + m[3] * 0.5 + m[4] * 0.5);
        }
        // ---- torchvision F.affine(angle=0, scale) -> _get_inverse_affine_matrix -> ImagingScaleAffine
        const double scale = scales ? scales[i] : 1.0;
        const double a_ = 1.0, b_ = -0.0, c_ = 0.0, d_ = 1.0;  // cos(0)/cos(0), -cos(0)*tan(0)/cos(0) - sin(0), ...
        double q[6] = {d_ / scale, -b_ / scale, 0.0 / scale, -c_ / scale, a_ / scale, 0.0 / scale};
        q[2] += q[0] * (-c0) + q[1] * (-c0);
        q[5] += q[3] * (-c0) + q[4] * (-c0);
        q[2] += c0;
        q[5] += c0;
        double* o = sc + 4 * i;
        o[0] = q[0];
        o[1] = q[2] + q[0] * 0.5;
        o[2] = q[4];
        o[3] = q[5] + q[4] * 0.5;
    }
}

}  // namespace sg
