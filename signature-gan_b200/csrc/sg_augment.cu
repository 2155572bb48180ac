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

constexpr int kAugWorkers = 256;              // threads that move pixels
constexpr int kAugThreads = kAugWorkers + 32;  // + one warp that only builds the scaling coordinate tables

// What the ncu captures showed (profiles/r01_ncu_augment.txt): the first version had every thread k < 2S rebuild table
// entry k with k dependent double additions — 4096 DADD per 64x64 image — and sat on the FP64 pipe (31 us per 4096
// images); with the chains moved to two lanes of a ninth warp it was issue-bound at 37 thread-instructions per pixel
// (28 us), most of them spent on the intermediate rotated tile. The two resamplings are now COMPOSED per output pixel —
// (y, x) -> scaled source (yt[y], xt[x]) -> its fixed-point rotation source — which is the same function, needs no
// second tile and no second barrier, and keeps the per-column products xt[x]*a0, xt[x]*a3 in registers.
template <int S>
__global__ void __launch_bounds__(kAugThreads) augment_kernel(const uint8_t* __restrict__ pool,
                                                              const int* __restrict__ index,
                                                              const int* __restrict__ rot,
                                                              const double* __restrict__ sc,
                                                              const uint8_t* __restrict__ flip, int batch,
                                                              float* __restrict__ out) {
    constexpr int QPR = S / 4;                   // quads (4 pixels) per row
    constexpr int RSTEP = kAugWorkers / QPR;     // rows between two consecutive quads of one thread
    constexpr int NQ = S * S / 4 / kAugWorkers;  // quads per thread per image
    __shared__ __align__(16) uint8_t src[S * S];
    __shared__ float lut[256];
    __shared__ short xt[S], yt[S];
    __shared__ int coef[8];
    const int tid = threadIdx.x;
    const bool worker = tid < kAugWorkers;
    // ToTensor + Normalize of every possible 8-bit value, with IEEE division (no reciprocal shortcuts)
    if (worker) lut[tid] = __fdiv_rn(__fsub_rn(__fdiv_rn(static_cast<float>(tid), 255.0f), 0.5f), 0.5f);
    const int qx = (tid % QPR) * 4, qy = tid / QPR;  // this worker's quad column / first row
    for (int img = blockIdx.x; img < batch; img += gridDim.x) {
        if (worker) {
            const long long src_img = index ? index[img] : img;
            const uint4* g = reinterpret_cast<const uint4*>(pool + src_img * (S * S));
#pragma unroll
            for (int i = 0; i < S * S / 16 / kAugWorkers; ++i)
                reinterpret_cast<uint4*>(src)[tid + i * kAugWorkers] = __ldg(g + tid + i * kAugWorkers);
        } else {
            const int lane = tid - kAugWorkers;
            if (lane < 2) {
                // scaling coordinates: the start value advanced by `o += a` in double once per output column / row — the
                // very sequence of roundings Pillow's serial loop produces; COORD(v) = v < 0 ? -1 : (int)v, -1 for v >= S
                const double a = sc[img * 4 + 2 * lane];  // lane 0: columns (a0, xo), lane 1: rows (a4, yo)
                double o = sc[img * 4 + 2 * lane + 1];
                short* tab = lane ? yt : xt;
                for (int k = 0; k < S; ++k) {
                    int v = o < 0.0 ? -1 : static_cast<int>(o);
                    if (v >= S) v = -1;
                    tab[k] = static_cast<short>(v);
                    o += a;
                }
            } else if (lane < 8) {
                coef[lane - 2] = rot[img * 6 + lane - 2];
            } else if (lane == 8) {
                coef[6] = (flip && flip[img]) ? 1 : 0;
            }
        }
        __syncthreads();
        if (worker) {
            // Pillow affine_fixed: source = ((a5 + y*a4 + x*a3) >> 16, (a2 + y*a1 + x*a0) >> 16), 32-bit integers
            const int a0 = coef[0], a1 = coef[1], a2 = coef[2], a3 = coef[3], a4 = coef[4], a5 = coef[5];
            const bool fl = coef[6] != 0;
            int cx0[4], cx3[4];
            bool xok[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int xi = xt[fl ? S - 1 - (qx + j) : qx + j];
                xok[j] = xi >= 0;
                cx0[j] = xi * a0;
                cx3[j] = xi * a3;
            }
            float4* o4 = reinterpret_cast<float4*>(out + static_cast<long long>(img) * (S * S)) + tid;
#pragma unroll
            for (int i = 0; i < NQ; ++i) {
                const int yi = yt[qy + i * RSTEP];
                const int bx = a2 + yi * a1, by = a5 + yi * a4;
                float v[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const unsigned xin = static_cast<unsigned>((bx + cx0[j]) >> 16);
                    const unsigned yin = static_cast<unsigned>((by + cx3[j]) >> 16);
                    const bool ok = xok[j] && yi >= 0 && xin < S && yin < S;
                    v[j] = ok ? lut[src[yin * S + xin]] : 1.0f;  // fill = 255 -> (255/255 - 0.5) / 0.5 = 1
                }
                o4[i * kAugWorkers] = make_float4(v[0], v[1], v[2], v[3]);
            }
        }
        __syncthreads();  // src / tables / coefficients are rewritten by the next image
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
    const int per_sm = 5;  // resident CTAs per SM (288 threads x 40 registers): one wave, every CTA strides over the batch
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
            // round(cos, 15), round(sin, 15), round(-sin, 15), round(cos, 15): decimal rounding is symmetric, so two
            // conversions serve the four entries (they dominate this function's time)
            const double rc = round15(std::cos(a)), rs = round15(std::sin(a));
            double m[6] = {rc, rs, 0.0, -rs, rc, 0.0};
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
