// sg_gfinal_mma.cu — bf16 implementation of the Generator's tail, Conv3x3 (32 -> 1) + bias + tanh (gen…:153-163), and
// its backward, on warp-level tensor-core MMAs.
//
// The streaming-stencil kernels of sg_gfinal.cu touch every activation byte once but spend ~1000 CUDA-core
// instructions per pixel (288 FMAs + the BatchNorm/ReLU recomputed for three neighbours), which makes them
// issue-bound at ~30 % of HBM bandwidth. Here the contraction over the 32 channels runs on mma.sync instead:
//   forward   P[pixel][tap] = sum_c relu(bn(y[pixel][c])) * w[tap][c]      (one m16n8k16 GEMM row block per 16 pixels,
//             N = 9 taps padded to 16), staged in a shared-memory ring of image rows; the output pixel is then the sum
//             of nine shifted P entries + bias, tanh — 9 adds per pixel instead of 288 FMAs.
//   backward  d[pixel][c] = sum_tap dpre[pixel - shift(tap)] * w[tap][c]   (K = taps), masked by relu', stored as bf16,
//             with the BatchNorm-backward sums (sum d, sum d*y) accumulated in registers, and
//             dW[tap][c] = sum_pixel dpre[pixel - shift(tap)] * a[pixel][c] (K = pixels; the activation fragment is
//             turned into the B operand with movmatrix.trans), accumulated in registers for the whole launch.
// A thread loads channels [8t, 8t+8) of pixels gid and gid+8 of a 16-pixel group (one 16-byte piece each: the K order
// of the MMAs is permuted to match, which costs nothing), so the BatchNorm+ReLU is applied exactly once per element
// in the load layout. Loads go through a per-thread cp.async ring (the thread that issues a piece is the one that
// consumes it: no barrier, 4 groups = 4 KB per warp in flight) — the kernels are HBM-bound.
#include "sg_elem.cuh"
#include "sg_kernels.cuh"
#include "sg_mma.cuh"

#include <cstdio>

namespace sg {
namespace {

constexpr int kC = 32;          // channels of the last generator level (gen…:139,149)
constexpr int kThr = 256;       // 8 warps
constexpr int kChunkPx = 512;   // pixels per chunk: 32 groups of 16 = 4 per warp
constexpr int kDepth = 4;       // cp.async ring depth (groups per thread)
constexpr int kGroupsPerWarp = kChunkPx / 16 / (kThr / 32);
constexpr int kStageBytes = kThr * kDepth * 32;   // per 32-channel group: two 16-byte pieces per thread and ring slot
// 2x-width variant (BASELINE configs[4]): the last level has 64 channels. The forward kernel loops over CG = 2 groups of
// 32 channels per pixel group (the tap sums of both halves go into the same accumulators); the backward kernels handle one
// 32-channel half per launch inside the LD = 64 channel tensor (the halves are independent but for the bias gradient).

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ uint32_t movmatrix_t(uint32_t a) {
    uint32_t d;
    asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;" : "=r"(d) : "r"(a));
    return d;
}
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

// y * scale + shift of 8 channels held as one 16-byte piece, BEFORE the ReLU: the ReLU is folded into the bf16 conversion
// of the value that feeds the tensor cores (pack2_relu: cvt.rn.relu.bf16x2, the same result as max(x, 0) then rounding),
// and its derivative is taken on the sign of this pre-activation (max(t, 0) > 0 <=> t > 0) — 16 FMNMX fewer per 16-pixel
// group in kernels that ncu shows issue-bound (60-67 % of the issue slots busy)
template <bool kAffine>
__device__ __forceinline__ void bn_pre8(const uint4& u, const float (&sc)[8], const float (&sh)[8], float (&a)[8]) {
    unpack8(u, a);
    if (kAffine) {
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] = fmaf(a[j], sc[j], sh[j]);
    }
}
// {relu(lo), relu(hi)} as packed bf16
__device__ __forceinline__ uint32_t pack2_relu(float lo, float hi) {
    uint32_t d;
    asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
    return d;
}

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
template <int S, int CG = 1>
struct FwdCfg {
    static constexpr int kPitch = S + 4;               // == 4 (mod 16): conflict-free accumulator stores; >= S + 2
    static constexpr int kCR = kChunkPx / S;           // image rows per chunk
    static constexpr int kRing = 2 * kCR + 2;          // rows being written (next chunk) + rows being read
    static constexpr int kChunks = S / kCR;
    static constexpr int kPBytes = kRing * 9 * kPitch * 4;
    static constexpr int kSmem = kPBytes + CG * kStageBytes;
};

template <int S, bool kAffine, int CG = 1>
__global__ void __launch_bounds__(kThr, 2)
gfinal_fwd_mma_kernel(const bf16* __restrict__ in, const float* __restrict__ scale, const float* __restrict__ shift,
                      const float* __restrict__ w, const float* __restrict__ bias, float* __restrict__ out,
                      uint8_t* __restrict__ out_u8, int B) {
    using Cfg = FwdCfg<S, CG>;
    constexpr int LD = kC * CG;  // channels per pixel of the input tensor
    constexpr int pitch = Cfg::kPitch, CR = Cfg::kCR, RING = Cfg::kRing, CHUNKS = Cfg::kChunks;
    constexpr int lgS = S == 64 ? 6 : 7;
    extern __shared__ __align__(16) uint8_t smem_raw[];
    float* P = reinterpret_cast<float*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, gid = lane >> 2, t = lane & 3;
    // piece p of channel group cg in ring slot d: + ((d * CG + cg) * 2 + p) * kThr * 16
    const uint32_t my_stage = smem_addr(smem_raw + Cfg::kPBytes) + tid * 16;

    // B fragments: n-block nb holds taps nb*8 + gid; k-step ks covers channels 8t+4ks .. 8t+4ks+3 of every lane quad
    uint32_t bw[CG][2][2][2];
    float sc[CG][8], sh[CG][8];
#pragma unroll
    for (int cg = 0; cg < CG; ++cg) {
#pragma unroll
        for (int nb = 0; nb < 2; ++nb) {
            const int tap = nb * 8 + gid;
#pragma unroll
            for (int ks = 0; ks < 2; ++ks)
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    const int ch = cg * kC + 8 * t + 4 * ks + 2 * r;
                    bw[cg][nb][ks][r] = tap < 9 ? pack2_bf16(w[ch * 9 + tap], w[(ch + 1) * 9 + tap]) : 0u;
                }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            sc[cg][j] = kAffine ? scale[cg * kC + 8 * t + j] : 1.f;
            sh[cg][j] = kAffine ? shift[cg * kC + 8 * t + j] : 0.f;
        }
    }
    const float b0 = bias[0];
    for (int i = tid; i < RING * 9 * pitch; i += kThr) P[i] = 0.f;  // border columns 0 and S+1 stay zero forever
    __syncthreads();

    // this warp's work items of the launch: (image, chunk, g) in order; item -> first pixel of the 16-pixel group
    const int n_img = (B - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
    const long total_items = static_cast<long>(n_img) * CHUNKS * kGroupsPerWarp;
    auto item_src = [&](long it) -> const bf16* {
        const int li = static_cast<int>(it / (CHUNKS * kGroupsPerWarp));
        const int r = static_cast<int>(it - static_cast<long>(li) * (CHUNKS * kGroupsPerWarp));
        const int c = r / kGroupsPerWarp, g = r - c * kGroupsPerWarp;
        const int n = blockIdx.x + li * gridDim.x;
        const int pix = c * kChunkPx + (warp * kGroupsPerWarp + g) * 16 + gid;
        return in + (static_cast<size_t>(n) * S * S + pix) * LD + t * 8;
    };
    auto issue = [&](long it) {
        if (it < total_items) {
            const bf16* src = item_src(it);
#pragma unroll
            for (int cg = 0; cg < CG; ++cg) {
                const uint32_t dst = my_stage + (static_cast<uint32_t>(it % kDepth) * CG + cg) * 2 * kThr * 16;
                cp_async16(dst, src + cg * kC);
                cp_async16(dst + kThr * 16, src + 8 * LD + cg * kC);
            }
        }
        cp_async_commit();
    };
    for (int d = 0; d < kDepth - 1; ++d) issue(d);

    auto emit = [&](int n, int yo, int x) {
        float acc = b0;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
            const int yy = yo + ky - 1;
            if (yy < 0 || yy >= S) continue;
            const float* pr = P + ((yy % RING) * 9 + ky * 3) * pitch + x;  // column x-1 of tap (ky, 0)
            acc += pr[0] + pr[pitch + 1] + pr[2 * pitch + 2];
        }
        const float v = tanhf(acc);
        const size_t o = (static_cast<size_t>(n) * S + yo) * S + x;
        out[o] = v;
        if (out_u8) {
            float q = (v + 1.f) * 127.5f;
            q = fminf(fmaxf(q, 0.f), 255.f);
            out_u8[o] = static_cast<uint8_t>(q);  // numpy astype(uint8) truncates (utils/inference.py:129)
        }
    };

    long it = 0;
    for (int n = blockIdx.x; n < B; n += gridDim.x) {
        for (int c = 0; c < CHUNKS; ++c) {
#pragma unroll 1
            for (int g = 0; g < kGroupsPerWarp; ++g, ++it) {
                issue(it + kDepth - 1);
                cp_async_wait<kDepth - 1>();
                float d0[4] = {0.f, 0.f, 0.f, 0.f}, d1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int cg = 0; cg < CG; ++cg) {
                    const uint8_t* st = smem_raw + Cfg::kPBytes + tid * 16 + ((it % kDepth) * CG + cg) * 2 * kThr * 16;
                    const uint4 u0 = *reinterpret_cast<const uint4*>(st);
                    const uint4 u1 = *reinterpret_cast<const uint4*>(st + kThr * 16);
                    float a0[8], a1[8];
                    bn_pre8<kAffine>(u0, sc[cg], sh[cg], a0);
                    bn_pre8<kAffine>(u1, sc[cg], sh[cg], a1);
                    uint32_t p0[4], p1[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        p0[j] = kAffine ? pack2_relu(a0[2 * j], a0[2 * j + 1]) : (&u0.x)[j];
                        p1[j] = kAffine ? pack2_relu(a1[2 * j], a1[2 * j + 1]) : (&u1.x)[j];
                    }
                    const uint32_t A0[4] = {p0[0], p1[0], p0[1], p1[1]};
                    const uint32_t A1[4] = {p0[2], p1[2], p0[3], p1[3]};
                    mma_bf16(d0, A0, bw[cg][0][0][0], bw[cg][0][0][1]);
                    mma_bf16(d1, A0, bw[cg][1][0][0], bw[cg][1][0][1]);
                    mma_bf16(d0, A1, bw[cg][0][1][0], bw[cg][0][1][1]);
                    mma_bf16(d1, A1, bw[cg][1][1][0], bw[cg][1][1][1]);
                }
                const int pix = c * kChunkPx + (warp * kGroupsPerWarp + g) * 16;
                const int yy = pix >> lgS, x = (pix & (S - 1)) + gid;
                float* pr = P + (yy % RING) * 9 * pitch + 1 + x;
                pr[(2 * t) * pitch] = d0[0];
                pr[(2 * t + 1) * pitch] = d0[1];
                pr[(2 * t) * pitch + 8] = d0[2];
                pr[(2 * t + 1) * pitch + 8] = d0[3];
                if (t == 0) {
                    pr[8 * pitch] = d1[0];
                    pr[8 * pitch + 8] = d1[2];
                }
            }
            __syncthreads();
            // rows c*CR-1 .. c*CR+CR-2 are complete (their lower neighbour row is in the ring)
#pragma unroll
            for (int i = 0; i < kChunkPx / kThr; ++i) {
                const int idx = tid + kThr * i;
                const int yo = c * CR - 1 + (idx >> lgS);
                if (yo >= 0) emit(n, yo, idx & (S - 1));
            }
            if (c == CHUNKS - 1)
                for (int x = tid; x < S; x += kThr) emit(n, S - 1, x);
        }
        __syncthreads();  // the next image's first chunk overwrites ring rows the emission above still reads
    }
    cp_async_wait<0>();
}

// ------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------
template <int S>
struct BwdCfg {
    static constexpr int kPitch = S + 4;
    static constexpr int kDpBytes = (S + 2) * kPitch * 4;     // d(pre-tanh) of one image with a zero border
    static constexpr int kRedFloats = 9 * kC + 2 * kC + 1;    // per warp: dW, (sum d, sum d*y), dbias
    static constexpr int kRedBytes = (kThr / 32) * kRedFloats * 4;
    static constexpr int kSmem = kDpBytes + kStageBytes + kRedBytes;
};

// MODE 0: one pass — masked d(bn output) written to dbn, plus dW / dbias / (sum d, sum d*y) partial rows.
// MODE 1: reductions only (nothing written per pixel).  MODE 2: recompute d and write the BatchNorm-backward result
//         dy = k1 * (d - k2 - xhat * k3) directly (bn_bwd_apply folded in): with batch statistics the two passes
//         1 + 2 move 3 activation-sized tensors through HBM where pass 0 + bn_bwd_apply move 5.
struct BnBwdCoef {
    const float *mean, *rstd, *k1, *k2, *k3;
};
template <int S, int MODE, int LD = kC>
__global__ void __launch_bounds__(kThr, 2)
gfinal_bwd_mma_kernel(const bf16* __restrict__ y, const float* __restrict__ scale, const float* __restrict__ shift,
                      const float* __restrict__ dout, const float* __restrict__ outimg, const float* __restrict__ w,
                      bf16* __restrict__ dbn, float* __restrict__ part_w, float* __restrict__ part_bn, int B,
                      const BnBwdCoef coef) {
    constexpr bool kSums = MODE != 2, kWrite = MODE != 1, kApply = MODE == 2;
    using Cfg = BwdCfg<S>;
    constexpr int pitch = Cfg::kPitch;
    constexpr int lgS = S == 64 ? 6 : 7;
    constexpr int CHUNKS = S * S / kChunkPx;
    extern __shared__ __align__(16) uint8_t smem_raw[];
    float* dp = reinterpret_cast<float*>(smem_raw);
    float* red = reinterpret_cast<float*>(smem_raw + Cfg::kDpBytes + kStageBytes);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, gid = lane >> 2, t = lane & 3;
    const uint32_t my_stage = smem_addr(smem_raw + Cfg::kDpBytes) + tid * 16;

    // dpre[yo + 1 - ky][xo + 1 - kx] relative to the dp element of (row yo, col xo) at index yo*pitch + xo
    auto tap_off = [&](int tap) { return (2 - tap / 3) * pitch + (2 - tap % 3); };
    const int offT0 = tap_off(2 * t), offT1 = tap_off(2 * t + 1), off8 = tap_off(8), offW = tap_off(gid);
    // data-gradient B fragments [k = tap][n = channel]: n-block j, column gid <-> channel 8*(gid/2) + 2j + (gid&1), so
    // that accumulator columns (2t, 2t+1) of block j are channels 8t+2j, 8t+2j+1 — register j of the loaded piece
    uint32_t bd[4][2];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int ch = 8 * (gid >> 1) + 2 * j + (gid & 1);
        bd[j][0] = pack2_bf16(w[ch * 9 + 2 * t], w[ch * 9 + 2 * t + 1]);
        bd[j][1] = t == 0 ? pack2_bf16(w[ch * 9 + 8], 0.f) : 0u;
    }
    float sc[8], sh[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        sc[j] = scale[8 * t + j];
        sh[j] = shift[8 * t + j];
    }
    // MODE 2: dy = k1*d + cA*y + cB per channel
    float k1v[8], cA[8], cB[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        k1v[j] = cA[j] = cB[j] = 0.f;
        if (kApply) {
            const int ch = 8 * t + j;
            const float k1 = coef.k1[ch], g = k1 * coef.k3[ch] * coef.rstd[ch];
            k1v[j] = k1;
            cA[j] = -g;
            cB[j] = g * coef.mean[ch] - k1 * coef.k2[ch];
        }
    }
    float accw[4][4], s0[8], s1[8], dsum = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int e = 0; e < 4; ++e) accw[j][e] = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) s0[j] = s1[j] = 0.f;
    for (int i = tid; i < (S + 2) * pitch; i += kThr) dp[i] = 0.f;

    constexpr int kItemsPerImage = CHUNKS * kGroupsPerWarp;
    const int n_img = (B - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
    const long total_items = static_cast<long>(n_img) * kItemsPerImage;
    auto issue = [&](long it) {
        if (it < total_items) {
            const int li = static_cast<int>(it / kItemsPerImage);
            const int r = static_cast<int>(it - static_cast<long>(li) * kItemsPerImage);
            const int c = r / kGroupsPerWarp, g = r - c * kGroupsPerWarp;
            const int n = blockIdx.x + li * gridDim.x;
            const int pix = c * kChunkPx + (warp * kGroupsPerWarp + g) * 16 + gid;
            const bf16* src = y + (static_cast<size_t>(n) * S * S + pix) * LD + t * 8;
            const uint32_t dst = my_stage + static_cast<uint32_t>(it % kDepth) * 2 * kThr * 16;
            cp_async16(dst, src);
            cp_async16(dst + kThr * 16, src + 8 * LD);
        }
        cp_async_commit();
    };
    for (int d = 0; d < kDepth - 1; ++d) issue(d);

    long it = 0;
    for (int n = blockIdx.x; n < B; n += gridDim.x) {
        __syncthreads();  // every warp is done reading the previous image's dp
        {
            const float* dsrc = dout + static_cast<size_t>(n) * S * S;
            const float* osrc = outimg + static_cast<size_t>(n) * S * S;
            for (int i = tid * 4; i < S * S; i += kThr * 4) {
                const float4 dd = __ldg(reinterpret_cast<const float4*>(dsrc + i));
                const float4 oo = __ldg(reinterpret_cast<const float4*>(osrc + i));
                float* q = dp + ((i >> lgS) + 1) * pitch + (i & (S - 1)) + 1;
                const float v0 = dd.x * (1.f - oo.x * oo.x), v1 = dd.y * (1.f - oo.y * oo.y);
                const float v2 = dd.z * (1.f - oo.z * oo.z), v3 = dd.w * (1.f - oo.w * oo.w);
                q[0] = v0;
                q[1] = v1;
                q[2] = v2;
                q[3] = v3;
                dsum += (v0 + v1) + (v2 + v3);
            }
        }
        __syncthreads();
#pragma unroll 1
        for (int r = 0; r < kItemsPerImage; ++r, ++it) {
            issue(it + kDepth - 1);
            cp_async_wait<kDepth - 1>();
            const uint8_t* st = smem_raw + Cfg::kDpBytes + tid * 16 + (it % kDepth) * 2 * kThr * 16;
            const uint4 u0 = *reinterpret_cast<const uint4*>(st);
            const uint4 u1 = *reinterpret_cast<const uint4*>(st + kThr * 16);
            const int c = r / kGroupsPerWarp, g = r - c * kGroupsPerWarp;
            const int pix0 = c * kChunkPx + (warp * kGroupsPerWarp + g) * 16;
            const float* dpp = dp + (pix0 >> lgS) * pitch + (pix0 & (S - 1));
            float a0[8], a1[8];
            bn_pre8<true>(u0, sc, sh, a0);   // pre-activations: > 0 <=> the ReLU passed the element
            bn_pre8<true>(u1, sc, sh, a1);
            // ---- data gradient: [16 pixels][taps] x [taps][32 channels]
            uint32_t A[4];
            A[0] = pack2_bf16(dpp[gid + offT0], dpp[gid + offT1]);
            A[1] = pack2_bf16(dpp[gid + 8 + offT0], dpp[gid + 8 + offT1]);
            A[2] = t == 0 ? pack2_bf16(dpp[gid + off8], 0.f) : 0u;
            A[3] = t == 0 ? pack2_bf16(dpp[gid + 8 + off8], 0.f) : 0u;
            float y0[8], y1[8];
            unpack8(u0, y0);
            unpack8(u1, y1);
            uint32_t o0[4], o1[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float cd[4] = {0.f, 0.f, 0.f, 0.f};
                mma_bf16(cd, A, bd[j][0], bd[j][1]);
                float e0 = a0[2 * j] > 0.f ? cd[0] : 0.f, e1 = a0[2 * j + 1] > 0.f ? cd[1] : 0.f;
                float f0 = a1[2 * j] > 0.f ? cd[2] : 0.f, f1 = a1[2 * j + 1] > 0.f ? cd[3] : 0.f;
                if (kSums) {
                    s0[2 * j] += e0 + f0;
                    s0[2 * j + 1] += e1 + f1;
                    s1[2 * j] = fmaf(e0, y0[2 * j], fmaf(f0, y1[2 * j], s1[2 * j]));
                    s1[2 * j + 1] = fmaf(e1, y0[2 * j + 1], fmaf(f1, y1[2 * j + 1], s1[2 * j + 1]));
                }
                if (kApply) {
                    e0 = fmaf(k1v[2 * j], e0, fmaf(cA[2 * j], y0[2 * j], cB[2 * j]));
                    e1 = fmaf(k1v[2 * j + 1], e1, fmaf(cA[2 * j + 1], y0[2 * j + 1], cB[2 * j + 1]));
                    f0 = fmaf(k1v[2 * j], f0, fmaf(cA[2 * j], y1[2 * j], cB[2 * j]));
                    f1 = fmaf(k1v[2 * j + 1], f1, fmaf(cA[2 * j + 1], y1[2 * j + 1], cB[2 * j + 1]));
                }
                o0[j] = pack2_bf16(e0, e1);
                o1[j] = pack2_bf16(f0, f1);
            }
            if (kWrite) {
                bf16* dst = dbn + (static_cast<size_t>(n) * S * S + pix0 + gid) * LD + t * 8;
                *reinterpret_cast<uint4*>(dst) = make_uint4(o0[0], o0[1], o0[2], o0[3]);
                *reinterpret_cast<uint4*>(dst + 8 * LD) = make_uint4(o1[0], o1[1], o1[2], o1[3]);
            }
            if (!kSums) continue;
            // ---- weight gradient: [taps][16 pixels] x [16 pixels][32 channels]
            uint32_t AW[4];
            AW[0] = pack2_bf16(dpp[2 * t + offW], dpp[2 * t + 1 + offW]);
            AW[2] = pack2_bf16(dpp[2 * t + 8 + offW], dpp[2 * t + 9 + offW]);
            AW[1] = gid == 0 ? pack2_bf16(dpp[2 * t + off8], dpp[2 * t + 1 + off8]) : 0u;
            AW[3] = gid == 0 ? pack2_bf16(dpp[2 * t + 8 + off8], dpp[2 * t + 9 + off8]) : 0u;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint32_t b0 = movmatrix_t(pack2_relu(a0[2 * j], a0[2 * j + 1]));
                const uint32_t b1 = movmatrix_t(pack2_relu(a1[2 * j], a1[2 * j + 1]));
                mma_bf16(accw[j], AW, b0, b1);
            }
        }
    }
    cp_async_wait<0>();
    if (!kSums) return;
    // ---- reduction: lanes of equal t hold partial (sum d, sum d*y) of channels 8t..8t+7
#pragma unroll
    for (int o = 4; o < 32; o <<= 1) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            s0[j] += __shfl_xor_sync(0xffffffffu, s0[j], o);
            s1[j] += __shfl_xor_sync(0xffffffffu, s1[j], o);
        }
    }
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) dsum += __shfl_xor_sync(0xffffffffu, dsum, o);
    float* rw = red + warp * Cfg::kRedFloats;
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int ch = 8 * t + 2 * j + e;
            rw[ch * 9 + gid] = accw[j][e];
            if (gid == 0) rw[ch * 9 + 8] = accw[j][2 + e];
        }
    if (lane < 4) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            rw[9 * kC + 8 * t + j] = s0[j];
            rw[9 * kC + kC + 8 * t + j] = s1[j];
        }
    }
    if (lane == 0) rw[9 * kC + 2 * kC] = dsum;
    __syncthreads();
    for (int i = tid; i < Cfg::kRedFloats; i += kThr) {
        float s = 0.f;
#pragma unroll
        for (int wv = 0; wv < kThr / 32; ++wv) s += red[wv * Cfg::kRedFloats + i];
        if (i < 9 * kC)
            part_w[static_cast<size_t>(blockIdx.x) * (9 * kC + 1) + i] = s;
        else if (i < 9 * kC + 2 * kC) {  // row layout [2][LD]: this launch's 32 channels inside the level's LD
            const int e = i - 9 * kC;
            part_bn[static_cast<size_t>(blockIdx.x) * 2 * LD + (e / kC) * LD + (e % kC)] = s;
        }
        else
            part_w[static_cast<size_t>(blockIdx.x) * (9 * kC + 1) + 9 * kC] = s;
    }
}

int sm_count_gm() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    }
    return n;
}

template <typename K>
int grid_for(K kernel, int smem, int B) {
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kThr, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
    const int slots = sm_count_gm() * per_sm;
    return B < slots ? B : slots;
}

}  // namespace

void gfinal_fwd_mma(const bf16* in, const float* scale, const float* shift, const float* w, const float* bias, float* out,
                    uint8_t* out_u8, int B, int S, cudaStream_t s, int C) {
    note_launch();
#define SG_GF_LAUNCH(SZ, AFF, CGN)                                                                                      \
    do {                                                                                                                \
        static int grid_cache = 0, grid_b = -1;                                                                         \
        if (grid_b != B) {                                                                                              \
            cudaFuncSetAttribute(gfinal_fwd_mma_kernel<SZ, AFF, CGN>, cudaFuncAttributeMaxDynamicSharedMemorySize,      \
                                 FwdCfg<SZ, CGN>::kSmem);                                                               \
            grid_cache = grid_for(gfinal_fwd_mma_kernel<SZ, AFF, CGN>, FwdCfg<SZ, CGN>::kSmem, B);                      \
            grid_b = B;                                                                                                 \
        }                                                                                                               \
        gfinal_fwd_mma_kernel<SZ, AFF, CGN><<<grid_cache, kThr, FwdCfg<SZ, CGN>::kSmem, s>>>(in, scale, shift, w, bias,  \
                                                                                           out, out_u8, B);             \
    } while (0)
    if (C == 2 * kC) {
        if (S == 64) {
            if (scale) SG_GF_LAUNCH(64, true, 2); else SG_GF_LAUNCH(64, false, 2);
        } else {
            if (scale) SG_GF_LAUNCH(128, true, 2); else SG_GF_LAUNCH(128, false, 2);
        }
    } else if (S == 64) {
        if (scale) SG_GF_LAUNCH(64, true, 1); else SG_GF_LAUNCH(64, false, 1);
    } else {
        if (scale) SG_GF_LAUNCH(128, true, 1); else SG_GF_LAUNCH(128, false, 1);
    }
#undef SG_GF_LAUNCH
}

// mode 0 / 1 / 2: see gfinal_bwd_mma_kernel. Returns the number of partial rows written to part_w ([chunks][9*32+1])
// and part_bn ([chunks][2][ld]) (modes 0 and 1). One launch handles 32 channels of a level with `ld` channels per pixel
// (32, or 64 for the 2x-width variant: the caller pre-offsets every per-channel pointer by the half it wants).
int gfinal_bwd_mma(const float* dout, const float* out, const bf16* y, const float* scale, const float* shift,
                   const float* w, bf16* dbn, float* part_w, float* part_bn, int B, int S, int mode, const float* mean,
                   const float* rstd, const float* k1, const float* k2, const float* k3, cudaStream_t s, int ld) {
    note_launch();
    const BnBwdCoef coef{mean, rstd, k1, k2, k3};
    int grid = 0;
#define SG_GB_LAUNCH(SZ, MD, LDN)                                                                                          \
    do {                                                                                                                   \
        static int per_b = -1, g_cache = 0;                                                                                \
        if (per_b != B) {                                                                                                  \
            g_cache = grid_for(gfinal_bwd_mma_kernel<SZ, MD, LDN>, BwdCfg<SZ>::kSmem, B);                                  \
            if (g_cache > kMaxChunks) g_cache = kMaxChunks;                                                                \
            per_b = B;                                                                                                     \
        }                                                                                                                  \
        grid = g_cache;                                                                                                    \
        gfinal_bwd_mma_kernel<SZ, MD, LDN><<<grid, kThr, BwdCfg<SZ>::kSmem, s>>>(y, scale, shift, dout, out, w, dbn, part_w, \
                                                                                part_bn, B, coef);                         \
    } while (0)
#define SG_GB_MODES(SZ, LDN)                                                                   \
    do {                                                                                       \
        if (mode == 0) SG_GB_LAUNCH(SZ, 0, LDN); else if (mode == 1) SG_GB_LAUNCH(SZ, 1, LDN); \
        else SG_GB_LAUNCH(SZ, 2, LDN);                                                         \
    } while (0)
    if (ld == 2 * kC) {
        if (S == 64) SG_GB_MODES(64, 64); else SG_GB_MODES(128, 64);
    } else {
        if (S == 64) SG_GB_MODES(64, 32); else SG_GB_MODES(128, 32);
    }
#undef SG_GB_MODES
#undef SG_GB_LAUNCH
    return grid;
}

}  // namespace sg
