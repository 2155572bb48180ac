// sg_gfinal.cu — the Generator's tail, Conv3x3 (32 -> 1) + bias + tanh (gen…:153-163), and its backward, written as
// streaming stencils. Both are HBM-bound (256 KB of bf16 activations per 64x64 image against 2.4 MFLOP), so the
// design goal is to touch every activation byte exactly once and keep many bytes in flight:
//   * one CTA per SM walks (image, 64-pixel-wide column strip) units; the strip's rows are streamed through a
//     4-slot shared-memory ring with cp.async.bulk (row segments are contiguous in NHWC), issued three chunks ahead
//     by one elected thread — the depth of the ring, not the number of resident warps, hides HBM latency. (No
//     dedicated producer warp: a ninth warp would put three warps on one scheduler and cap registers at 168.)
//   * a consumer thread owns (pixel column x, 8-channel slice g) and slides down the strip with the 3x3 window in
//     registers, so each staged row is read from shared memory once per horizontal neighbour and never re-fetched;
//   * in training mode the input is the PRE-BatchNorm convolution output and relu(y*scale+shift) is applied while
//     loading: the normalised activation of the last upsample block never exists in HBM, forward or backward;
//   * the backward kernel produces, in the same pass over y: the masked data gradient, the 3x3 weight gradient,
//     the bias gradient and the BatchNorm-backward reductions (sum d, sum d*y) of the last block.
#include "sg_elem.cuh"
#include "sg_kernels.cuh"
#include "sg_umma.cuh"

#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace sg {
namespace {
// LeakyReLU with slope in [0, 1): slope 0 is ReLU and returns +0 for negative inputs (as fmaxf(x, 0) does)
__device__ __forceinline__ float act_leaky(float x, float slope) { return fmaxf(x, 0.f) + slope * fminf(x, 0.f); }


constexpr int kFC = 32;      // channels of the last generator level (gen…:139,149)
constexpr int kStripW = 64;  // pixels per strip row
constexpr int kSlots = 4;
constexpr int kConsumers = kStripW * 4;       // (pixel, 8-channel slice)
constexpr int kThreadsG = kConsumers;
constexpr int kRedFloats = 9 * kFC + 1 + 2 * kFC;  // dW, dbias, (sum d, sum d*y)

template <typename T>
struct StripCfg {
    static constexpr int kPxBytes = kFC * sizeof(T);
    static constexpr int kPitch = (kStripW + 2) * kPxBytes;  // one halo pixel on either side
    static constexpr int kRows = sizeof(T) == 2 ? 8 : 4;     // rows per ring slot
    static constexpr int kActBytes = kRows * kPitch;
    // backward only: rows of dout and out (fp32 images), columns [x0-4, x0+68) so that every copy is 16-byte aligned
    static constexpr int kImgPitch = (kStripW + 8) * 4;
    static constexpr int kImgBytes = kRows * 2 * kImgPitch;
    static constexpr int kSlotBytes = kActBytes + kImgBytes;
    static constexpr int kSmem = kSlots * kSlotBytes + 2 * kSlots * 8 + 128 /*alignment*/;
};

__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ void lds8(const uint8_t* p, float (&f)[8], bf16) {
    const uint4 u = *reinterpret_cast<const uint4*>(p);
    unpack8(u, f);
}
__device__ __forceinline__ void lds8(const uint8_t* p, float (&f)[8], float) {
    const float4 a = reinterpret_cast<const float4*>(p)[0];
    const float4 b = reinterpret_cast<const float4*>(p)[1];
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w;
    f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}

// Chunk q of this CTA = rows [c*kRows, (c+1)*kRows) of unit blockIdx.x + (q / chunks) * gridDim.x. Called by one
// thread, kSlots-1 chunks ahead of the consumers.
template <typename T>
__device__ __forceinline__ void strip_issue(const T* __restrict__ in, uint8_t* smem, uint64_t* full, uint64_t* empty,
                                            uint32_t q, int S, const float* __restrict__ img0 = nullptr,
                                            const float* __restrict__ img1 = nullptr) {
    using Cfg = StripCfg<T>;
    const int strips = S / kStripW, chunks = S / Cfg::kRows;
    const int u = blockIdx.x + static_cast<int>(q / chunks) * gridDim.x, c = static_cast<int>(q % chunks);
    const int n = u / strips, x0 = (u - n * strips) * kStripW;
    const int xs = x0 > 0 ? x0 - 1 : 0;
    const int xe = x0 + kStripW + 1 < S ? x0 + kStripW + 1 : S;
    const uint32_t bytes = static_cast<uint32_t>(xe - xs) * Cfg::kPxBytes;
    const uint32_t dst_off = static_cast<uint32_t>(xs - (x0 - 1)) * Cfg::kPxBytes;
    const int slot = q % kSlots;
    mbar_wait(&empty[slot], ((q / kSlots) & 1) ^ 1);  // every consumer warp has released the slot's previous chunk
    const int is = x0 >= 4 ? x0 - 4 : 0;
    const int ie = x0 + kStripW + 4 < S ? x0 + kStripW + 4 : S;
    const uint32_t ibytes = img0 ? static_cast<uint32_t>(ie - is) * 4 : 0;
    mbar_arrive_expect_tx(&full[slot], (bytes + 2 * ibytes) * Cfg::kRows);
    uint8_t* dst = smem + slot * Cfg::kSlotBytes + dst_off;
    const T* src = in + ((static_cast<size_t>(n) * S + c * Cfg::kRows) * S + xs) * kFC;
#pragma unroll
    for (int r = 0; r < Cfg::kRows; ++r)
        bulk_g2s(dst + r * Cfg::kPitch, src + static_cast<size_t>(r) * S * kFC, bytes, &full[slot]);
    if (img0) {
        uint8_t* idst = smem + slot * Cfg::kSlotBytes + Cfg::kActBytes + static_cast<uint32_t>(is - (x0 - 4)) * 4;
        const size_t ioff = (static_cast<size_t>(n) * S + c * Cfg::kRows) * S + is;
#pragma unroll
        for (int r = 0; r < Cfg::kRows; ++r) {
            bulk_g2s(idst + (2 * r) * Cfg::kImgPitch, img0 + ioff + static_cast<size_t>(r) * S, ibytes, &full[slot]);
            bulk_g2s(idst + (2 * r + 1) * Cfg::kImgPitch, img1 + ioff + static_cast<size_t>(r) * S, ibytes, &full[slot]);
        }
    }
}

template <typename T>
__device__ __forceinline__ uint8_t* strip_setup(uint8_t* smem_raw, uint64_t*& full, uint64_t*& empty) {
    using Cfg = StripCfg<T>;
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
    full = reinterpret_cast<uint64_t*>(smem + kSlots * Cfg::kSlotBytes);
    empty = full + kSlots;
    if (threadIdx.x == 0) {
        for (int s = 0; s < kSlots; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], kConsumers / 32);
        }
        mbar_fence_init();
    }
    __syncthreads();
    return smem;
}

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
template <typename T, bool kAffine>
__global__ void __launch_bounds__(kThreadsG, 1)
gfinal_fwd_kernel(const T* __restrict__ in, const float* __restrict__ scale, const float* __restrict__ shift,
                  const float* __restrict__ w, const float* __restrict__ bias, float* __restrict__ out,
                  uint8_t* __restrict__ out_u8, int B, int S, float act_slope) {
    using Cfg = StripCfg<T>;
    extern __shared__ uint8_t smem_raw[];
    uint64_t *full, *empty;
    uint8_t* smem = strip_setup<T>(smem_raw, full, empty);
    const int tid = threadIdx.x, lane = tid & 31;
    const int xl = tid >> 2, g = tid & 3;
    float wr[9][8], sc[8], sh[8];
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
        for (int j = 0; j < 8; ++j) wr[t][j] = w[(g * 8 + j) * 9 + t];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        sc[j] = kAffine ? scale[g * 8 + j] : 1.f;
        sh[j] = kAffine ? shift[g * 8 + j] : 0.f;
    }
    const float b0 = bias[0];
    const int strips = S / kStripW, units = B * strips, chunks = S / Cfg::kRows;
    const uint32_t my_off = static_cast<uint32_t>(xl * kFC + g * 8) * sizeof(T);  // column x-1 of this thread
    const uint32_t total_q = static_cast<uint32_t>((units - blockIdx.x + gridDim.x - 1) / gridDim.x) * chunks;
    if (tid == 0)
        for (uint32_t q = 0; q < kSlots - 1 && q < total_q; ++q) strip_issue<T>(in, smem, full, empty, q, S);
    uint32_t it = 0;
    for (int u = blockIdx.x; u < units; u += gridDim.x) {
        const int n = u / strips, x0 = (u - n * strips) * kStripW, x = x0 + xl;
        const bool okl = x > 0, okr = x + 1 < S;
        float w0[3][8], w1[3][8], w2[3][8];  // window rows y-1, y, y+1; [column x-1, x, x+1][channel]
#pragma unroll
        for (int d = 0; d < 3; ++d)
#pragma unroll
            for (int j = 0; j < 8; ++j) w0[d][j] = w1[d][j] = w2[d][j] = 0.f;
        float hold = 0.f;
        // `emit` computes output row yo from the window once row yo+1 has been shifted in.
        auto emit = [&](int yo) {
            float a0 = 0.f, a1 = 0.f, a2 = 0.f;
#pragma unroll
            for (int d = 0; d < 3; ++d)
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    a0 = fmaf(w0[d][j], wr[d][j], a0);
                    a1 = fmaf(w1[d][j], wr[3 + d][j], a1);
                    a2 = fmaf(w2[d][j], wr[6 + d][j], a2);
                }
            float acc = a0 + a1 + a2;
            acc += __shfl_xor_sync(0xffffffffu, acc, 1);
            acc += __shfl_xor_sync(0xffffffffu, acc, 2);
            // slice g keeps the sum of row (yo & ~3) + g: one tanh + store per thread every four rows
            if ((yo & 3) == g) hold = acc;
            if ((yo & 3) == 3) {
                const float t = tanhf(hold + b0);
                const size_t o = (static_cast<size_t>(n) * S + (yo - 3 + g)) * S + x;
                out[o] = t;
                if (out_u8) {
                    float q = (t + 1.f) * 127.5f;
                    q = fminf(fmaxf(q, 0.f), 255.f);
                    out_u8[o] = static_cast<uint8_t>(q);  // numpy astype(uint8) truncates (utils/inference.py:129)
                }
            }
        };
        for (int c = 0; c < chunks; ++c, ++it) {
            const int slot = it % kSlots;
            if (tid == 0 && it + kSlots - 1 < total_q) strip_issue<T>(in, smem, full, empty, it + kSlots - 1, S);
            __syncwarp();
            mbar_wait(&full[slot], (it / kSlots) & 1);
            const uint8_t* base = smem + slot * Cfg::kSlotBytes + my_off;
#pragma unroll
            for (int r = 0; r < Cfg::kRows; ++r) {
#pragma unroll
                for (int d = 0; d < 3; ++d)
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        w0[d][j] = w1[d][j];
                        w1[d][j] = w2[d][j];
                    }
#pragma unroll
                for (int d = 0; d < 3; ++d) {
                    float v[8];
                    lds8(base + r * Cfg::kPitch + d * Cfg::kPxBytes, v, T());
                    const bool ok = d == 0 ? okl : (d == 2 ? okr : true);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        float t = kAffine ? act_leaky(fmaf(v[j], sc[j], sh[j]), act_slope) : v[j];
                        w2[d][j] = ok ? t : 0.f;
                    }
                }
                const int y = c * Cfg::kRows + r;  // row just shifted in
                if (y > 0) emit(y - 1);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[slot]);
        }
        // zero row below the image
#pragma unroll
        for (int d = 0; d < 3; ++d)
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                w0[d][j] = w1[d][j];
                w1[d][j] = w2[d][j];
                w2[d][j] = 0.f;
            }
        emit(S - 1);
    }
}

// ------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kThreadsG, 1)
gfinal_bwd_kernel(const T* __restrict__ y, const float* __restrict__ scale, const float* __restrict__ shift,
                  const float* __restrict__ dout, const float* __restrict__ out, const float* __restrict__ w,
                  T* __restrict__ dbn, float* __restrict__ part_w, float* __restrict__ part_bn, int B, int S,
                  float act_slope) {
    using Cfg = StripCfg<T>;
    extern __shared__ uint8_t smem_raw[];
    __shared__ float red[kConsumers / 32][4][kRedFloats / 4 + 1];  // per warp, per slice: 72 dW + 8 s0 + 8 s1 (+ dbias)
    uint64_t *full, *empty;
    uint8_t* smem = strip_setup<T>(smem_raw, full, empty);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int xl = tid >> 2, g = tid & 3;
    float wr[9][8], aw[9][8], sc[8], sh[8], s0[8], s1[8];
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            wr[t][j] = w[(g * 8 + j) * 9 + t];
            aw[t][j] = 0.f;
        }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        sc[j] = scale[g * 8 + j];
        sh[j] = shift[g * 8 + j];
        s0[j] = s1[j] = 0.f;
    }
    float dsum = 0.f;
    const int strips = S / kStripW, units = B * strips, chunks = S / Cfg::kRows;
    const uint32_t my_off = static_cast<uint32_t>((xl + 1) * kFC + g * 8) * sizeof(T);  // this thread's own pixel
    const uint32_t total_q = static_cast<uint32_t>((units - blockIdx.x + gridDim.x - 1) / gridDim.x) * chunks;
    if (tid == 0)
        for (uint32_t q = 0; q < kSlots - 1 && q < total_q; ++q) strip_issue<T>(y, smem, full, empty, q, S, dout, out);
    uint32_t it = 0;
    for (int u = blockIdx.x; u < units; u += gridDim.x) {
        const int n = u / strips, x0 = (u - n * strips) * kStripW, x = x0 + xl;
        // slices 0..2 of a pixel compute d(pre-tanh) of columns x-1, x, x+1 and share them by shuffle
        const int col = x - 1 + g;
        const bool col_ok = g < 3 && col >= 0 && col < S;
        float d0[3] = {0.f, 0.f, 0.f}, d1[3] = {0.f, 0.f, 0.f}, d2[3] = {0.f, 0.f, 0.f};  // d(pre) rows y-1, y, y+1
        float ap[8], yp[8];  // relu(bn(y)) and raw y of the row whose outputs are computed next
#pragma unroll
        for (int j = 0; j < 8; ++j) ap[j] = yp[j] = 0.f;
        // shift d(pre) row `row` in (zeros for row == S) and emit everything for row `row - 1`
        auto step = [&](int row, float mine) {
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                d0[k] = d1[k];
                d1[k] = d2[k];
                d2[k] = __shfl_sync(0xffffffffu, mine, (lane & ~3) + k);
            }
            if (row == 0) return;
            // dv[ky*3+kx] = dpre[yo + 1 - ky][x + 1 - kx]
            const float dv[9] = {d2[2], d2[1], d2[0], d1[2], d1[1], d1[0], d0[2], d0[1], d0[0]};
            float d[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) d[j] = 0.f;
#pragma unroll
            for (int t = 0; t < 9; ++t)
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    d[j] = fmaf(dv[t], wr[t][j], d[j]);
                    aw[t][j] = fmaf(dv[t], ap[j], aw[t][j]);
                }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                d[j] = ap[j] > 0.f ? d[j] : d[j] * act_slope;
                s0[j] += d[j];
                s1[j] = fmaf(d[j], yp[j], s1[j]);
            }
            if (g == 0) dsum += dv[4];
            store8(dbn + ((static_cast<size_t>(n) * S + (row - 1)) * S + x) * kFC + g * 8, d);
        };
        for (int c = 0; c < chunks; ++c, ++it) {
            const int slot = it % kSlots;
            if (tid == 0 && it + kSlots - 1 < total_q)
                strip_issue<T>(y, smem, full, empty, it + kSlots - 1, S, dout, out);
            __syncwarp();
            mbar_wait(&full[slot], (it / kSlots) & 1);
            const uint8_t* base = smem + slot * Cfg::kSlotBytes + my_off;
            const float* img = reinterpret_cast<const float*>(smem + slot * Cfg::kSlotBytes + Cfg::kActBytes) + xl + 3 + g;
#pragma unroll
            for (int r = 0; r < Cfg::kRows; ++r) {
                const int row = c * Cfg::kRows + r;
                float mine = 0.f;
                if (col_ok) {
                    const float dd = img[(2 * r) * (Cfg::kImgPitch / 4)], oo = img[(2 * r + 1) * (Cfg::kImgPitch / 4)];
                    mine = dd * (1.f - oo * oo);
                }
                step(row, mine);
                lds8(base + r * Cfg::kPitch, yp, T());
#pragma unroll
                for (int j = 0; j < 8; ++j) ap[j] = act_leaky(fmaf(yp[j], sc[j], sh[j]), act_slope);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[slot]);
        }
        step(S, 0.f);
    }
    // ---- block reduction: lanes with equal g hold partial sums of the same outputs
#pragma unroll
    for (int o = 4; o < 32; o <<= 1) {
#pragma unroll
        for (int t = 0; t < 9; ++t)
#pragma unroll
            for (int j = 0; j < 8; ++j) aw[t][j] += __shfl_xor_sync(0xffffffffu, aw[t][j], o);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            s0[j] += __shfl_xor_sync(0xffffffffu, s0[j], o);
            s1[j] += __shfl_xor_sync(0xffffffffu, s1[j], o);
        }
        dsum += __shfl_xor_sync(0xffffffffu, dsum, o);
    }
    if (lane < 4) {
        float* r = red[warp][g];
#pragma unroll
        for (int t = 0; t < 9; ++t)
#pragma unroll
            for (int j = 0; j < 8; ++j) r[j * 9 + t] = aw[t][j];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            r[72 + j] = s0[j];
            r[80 + j] = s1[j];
        }
        r[88] = dsum;
    }
    __syncthreads();
    constexpr int NW = kConsumers / 32;
    float* pw = part_w + static_cast<size_t>(blockIdx.x) * (9 * kFC + 1);
    float* pb = part_bn + static_cast<size_t>(blockIdx.x) * 2 * kFC;
    for (int i = tid; i < 9 * kFC; i += kConsumers) {  // i = c*9 + t, c = gg*8 + j
        const int gg = i / 72, k = i - gg * 72;
        float s = 0.f;
#pragma unroll
        for (int wv = 0; wv < NW; ++wv) s += red[wv][gg][k];
        pw[i] = s;
    }
    if (tid < 2 * kFC) {  // [which][c]
        const int which = tid / kFC, c = tid - which * kFC;
        float s = 0.f;
#pragma unroll
        for (int wv = 0; wv < NW; ++wv) s += red[wv][c >> 3][72 + which * 8 + (c & 7)];
        pb[tid] = s;
    }
    if (tid == 2 * kFC) {
        float s = 0.f;
#pragma unroll
        for (int wv = 0; wv < NW; ++wv) s += red[wv][0][88];
        pw[9 * kFC] = s;
    }
}

int sm_count_g() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    }
    return n;
}

template <typename K>
bool set_smem(K kernel, int bytes) {
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes) == cudaSuccess;
}

}  // namespace

// sg_gfinal_mma.cu
void gfinal_fwd_mma(const bf16* in, const float* scale, const float* shift, const float* w, const float* bias, float* out,
                    uint8_t* out_u8, int B, int S, cudaStream_t s, int C = 32);
int gfinal_bwd_mma(const float* dout, const float* out, const bf16* y, const float* scale, const float* shift,
                   const float* w, bf16* dbn, float* part_w, float* part_bn, int B, int S, int mode, const float* mean,
                   const float* rstd, const float* k1, const float* k2, const float* k3, cudaStream_t s, int ld = 32);

// SIGGAN_GFINAL=stencil keeps the bf16 path on the streaming-stencil kernels (A/B comparison in the harness).
static bool use_stencil_bf16() {
    static const bool v = [] {
        const char* e = getenv("SIGGAN_GFINAL");
        return e && !strcmp(e, "stencil");
    }();
    return v;
}

template <typename T>
void final_conv_tanh_stencil(const T* in, const float* scale, const float* shift, const float* w, const float* bias,
                             float* out, uint8_t* out_u8, int B, int S, int C, float act_slope, cudaStream_t s) {
    if (C != kFC || S % kStripW != 0) {  // the streaming stencil is specialised to the reference's 32-channel last level
        note_unsupported("final_conv_tanh (stencil)", C);
        return;
    }
    using Cfg = StripCfg<T>;
    const int units = B * (S / kStripW);
    const int grid = units < sm_count_g() ? units : sm_count_g();
    note_launch();
    if (scale) {
        static bool ok = set_smem(gfinal_fwd_kernel<T, true>, Cfg::kSmem);
        (void)ok;
        gfinal_fwd_kernel<T, true><<<grid, kThreadsG, Cfg::kSmem, s>>>(in, scale, shift, w, bias, out, out_u8, B, S,
                                                                       act_slope);
    } else {
        static bool ok = set_smem(gfinal_fwd_kernel<T, false>, Cfg::kSmem);
        (void)ok;
        gfinal_fwd_kernel<T, false><<<grid, kThreadsG, Cfg::kSmem, s>>>(in, scale, shift, w, bias, out, out_u8, B, S,
                                                                       act_slope);
    }
}

template <typename T>
int final_conv_bwd_stencil(const float* dout, const float* out, const T* y, const float* scale, const float* shift,
                           const float* w, T* dbn, float* dW, float* dbias, float* part_w, float* part_bn, int B, int S,
                           int C, float act_slope, cudaStream_t s) {
    if (C != kFC || S % kStripW != 0) {
        note_unsupported("final_conv_bwd (stencil)", C);
        return 0;
    }
    using Cfg = StripCfg<T>;
    const int units = B * (S / kStripW);
    const int grid = units < sm_count_g() ? units : sm_count_g();
    static bool ok = set_smem(gfinal_bwd_kernel<T>, Cfg::kSmem);
    (void)ok;
    note_launch();
    gfinal_bwd_kernel<T><<<grid, kThreadsG, Cfg::kSmem, s>>>(y, scale, shift, dout, out, w, dbn, part_w, part_bn, B, S,
                                                             act_slope);
    vec_finalize(part_w, grid, 9 * kFC + 1, dW, 9 * kFC, dbias, s);
    return grid;
}

template void final_conv_tanh_stencil<float>(const float*, const float*, const float*, const float*, const float*, float*,
                                             uint8_t*, int, int, int, float, cudaStream_t);
template void final_conv_tanh_stencil<bf16>(const bf16*, const float*, const float*, const float*, const float*, float*,
                                            uint8_t*, int, int, int, float, cudaStream_t);
template int final_conv_bwd_stencil<float>(const float*, const float*, const float*, const float*, const float*,
                                           const float*, float*, float*, float*, float*, float*, int, int, int, float,
                                           cudaStream_t);
template int final_conv_bwd_stencil<bf16>(const float*, const float*, const bf16*, const float*, const float*,
                                          const float*, bf16*, float*, float*, float*, float*, int, int, int, float,
                                          cudaStream_t);

template <>
void final_conv_tanh<float>(const float* in, const float* scale, const float* shift, const float* w, const float* bias,
                            float* out, uint8_t* out_u8, int B, int S, int C, float act_slope, cudaStream_t s) {
    final_conv_tanh_stencil<float>(in, scale, shift, w, bias, out, out_u8, B, S, C, act_slope, s);
}
template <>
void final_conv_tanh<bf16>(const bf16* in, const float* scale, const float* shift, const float* w, const float* bias,
                           float* out, uint8_t* out_u8, int B, int S, int C, float act_slope, cudaStream_t s) {
    // the mma.sync kernels are specialised to ReLU; the LeakyReLU generator of the ablation runs the streaming stencil
    if (C == 2 * kFC && (S == 64 || S == 128) && !(scale && act_slope != 0.f))   // 2x-width variant: mma.sync kernel only
        return gfinal_fwd_mma(in, scale, shift, w, bias, out, out_u8, B, S, s, C);
    if (use_stencil_bf16() || C != kFC || (S != 64 && S != 128) || (scale && act_slope != 0.f))
        return final_conv_tanh_stencil<bf16>(in, scale, shift, w, bias, out, out_u8, B, S, C, act_slope, s);
    gfinal_fwd_mma(in, scale, shift, w, bias, out, out_u8, B, S, s);
}

template <>
int final_conv_bwd<float>(const float* dout, const float* out, const float* y, const float* scale, const float* shift,
                          const float* w, float* dbn, float* dW, float* dbias, float* part_w, float* part_bn, int B,
                          int S, int C, float act_slope, cudaStream_t s) {
    return final_conv_bwd_stencil<float>(dout, out, y, scale, shift, w, dbn, dW, dbias, part_w, part_bn, B, S, C, act_slope,
                                         s);
}
template <>
int final_conv_bwd<bf16>(const float* dout, const float* out, const bf16* y, const float* scale, const float* shift,
                         const float* w, bf16* dbn, float* dW, float* dbias, float* part_w, float* part_bn, int B, int S,
                         int C, float act_slope, cudaStream_t s) {
    if (C == 2 * kFC && (S == 64 || S == 128) && act_slope == 0.f) {
        // 2x-width variant: one launch per 32-channel half of the 64-channel level; the halves share d(pre-tanh) and are
        // independent otherwise (the bias gradient is taken from the first)
        int grid = 0;
        for (int c0 = 0; c0 < C; c0 += kFC) {
            grid = gfinal_bwd_mma(dout, out, y + c0, scale + c0, shift + c0, w + c0 * 9, dbn ? dbn + c0 : nullptr, part_w,
                                  part_bn + c0, B, S, dbn ? 0 : 1, nullptr, nullptr, nullptr, nullptr, nullptr, s, C);
            vec_finalize(part_w, grid, 9 * kFC + 1, dW + c0 * 9, 9 * kFC, c0 == 0 ? dbias : nullptr, s);
        }
        return grid;
    }
    if (use_stencil_bf16() || C != kFC || (S != 64 && S != 128) || act_slope != 0.f)
        return final_conv_bwd_stencil<bf16>(dout, out, y, scale, shift, w, dbn, dW, dbias, part_w, part_bn, B, S, C,
                                            act_slope, s);
    const int grid = gfinal_bwd_mma(dout, out, y, scale, shift, w, dbn, part_w, part_bn, B, S, dbn ? 0 : 1, nullptr,
                                    nullptr, nullptr, nullptr, nullptr, s);
    vec_finalize(part_w, grid, 9 * kFC + 1, dW, 9 * kFC, dbias, s);
    return grid;
}

bool final_conv_bwd_two_pass(int S, int C, float act_slope) {
    return (S == 64 || S == 128) && act_slope == 0.f && ((C == kFC && !use_stencil_bf16()) || C == 2 * kFC);
}
void final_conv_bwd_apply(const float* dout, const float* out, const bf16* y, const float* scale, const float* shift,
                          const float* w, const float* mean, const float* rstd, const float* k1, const float* k2,
                          const float* k3, bf16* dy, int B, int S, int C, cudaStream_t s) {
    for (int c0 = 0; c0 < C; c0 += kFC)
        gfinal_bwd_mma(dout, out, y + c0, scale + c0, shift + c0, w + c0 * 9, dy + c0, nullptr, nullptr, B, S, 2, mean + c0,
                       rstd + c0, k1 + c0, k2 + c0, k3 + c0, s, C);
}

}  // namespace sg
