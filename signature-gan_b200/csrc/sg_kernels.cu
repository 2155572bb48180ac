// sg_kernels.cu — CUDA-core kernels of the siggan hot path (see sg_kernels.cuh).
#include "sg_kernels.cuh"
#include "sg_elem.cuh"

#include <cmath>
#include <cstdio>
#include <type_traits>

namespace sg {

unsigned long long g_launches = 0;
static thread_local char k_err[256] = "";
const char* kernels_last_error() { return k_err; }
static thread_local bool k_unsupported = false;  // a launcher refused its arguments since the last kernels_check
void note_unsupported(const char* what, int C) {
    snprintf(k_err, sizeof(k_err), "%s: %d channels unsupported in this precision mode (no kernel launched)", what, C);
    k_unsupported = true;
}
int kernels_check(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        snprintf(k_err, sizeof(k_err), "%s: %s", what, cudaGetErrorString(e));
        k_unsupported = false;
        return -1;
    }
    if (k_unsupported) {  // k_err holds the launcher's message
        k_unsupported = false;
        return -1;
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------
// packing
// ------------------------------------------------------------------------------------------------
__global__ void pack_w16_kernel(const float* __restrict__ src, bf16* __restrict__ dstAB, bf16* __restrict__ dstBA,
                                int A, int B) {
    const long total = static_cast<long>(A) * B * 16;
    for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long>(gridDim.x) * blockDim.x) {
        const int t = static_cast<int>(i & 15);
        const long ab = i >> 4;
        const int b = static_cast<int>(ab % B);
        const int a = static_cast<int>(ab / B);
        const bf16 v = __float2bfloat16(src[i]);
        if (dstAB) dstAB[(static_cast<long>(a) * 16 + t) * B + b] = v;
        if (dstBA) dstBA[(static_cast<long>(b) * 16 + t) * A + a] = v;
    }
}
void pack_w16(const float* src, bf16* dstAB, bf16* dstBA, int A, int B, cudaStream_t s) {
    note_launch();
    pack_w16_kernel<<<blocks_for(static_cast<long>(A) * B * 16, 256), 256, 0, s>>>(src, dstAB, dstBA, A, B);
}

__global__ void pack_fc_kernel(const float* __restrict__ W, const float* __restrict__ bias, bf16* __restrict__ Wp,
                               float* __restrict__ biasp, int C0, int latent, int Kp) {
    const int F = C0 * 16;
    const long total = static_cast<long>(F) * Kp;
    for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long>(gridDim.x) * blockDim.x) {
        const int k = static_cast<int>(i % Kp);
        const int j = static_cast<int>(i / Kp);
        const int f = (j % C0) * 16 + j / C0;
        Wp[i] = __float2bfloat16(k < latent ? W[static_cast<long>(f) * latent + k] : 0.f);
        if (k == 0) biasp[j] = bias[f];
    }
}
void pack_fc(const float* W, const float* bias, bf16* Wp, float* biasp, int C0, int latent, int Kp, cudaStream_t s) {
    note_launch();
    pack_fc_kernel<<<blocks_for(static_cast<long>(C0) * 16 * Kp, 256), 256, 0, s>>>(W, bias, Wp, biasp, C0, latent, Kp);
}

__global__ void pack_classifier_kernel(const float* __restrict__ w, float* __restrict__ wp, int C) {
    const int F = C * 16;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < F; j += gridDim.x * blockDim.x)
        wp[j] = w[(j % C) * 16 + j / C];
}
void pack_classifier(const float* w, float* wp, int C, cudaStream_t s) {
    note_launch();
    pack_classifier_kernel<<<blocks_for(C * 16, 256), 256, 0, s>>>(w, wp, C);
}

// One launch for all weight packs of a network (table in the parameter bank). w16 segments are transposed through
// shared memory in 32 x 32 x 16 tiles so that both packed layouts are written in 64-byte runs; the fc / classifier
// permutations are small elementwise sections (4096 elements per block).
__global__ void __launch_bounds__(256) pack_plan_kernel(const __grid_constant__ PackPlan plan) {
    // w16 tile: 16 a x 32 b x 16 taps, rows of (b, tap) padded by one 4-byte word against bank conflicts of the
    // transposed reads
    constexpr int kRow = 32 * 16 + 2;
    __shared__ bf16 tile[16 * kRow];
    int sidx = 0;
    while (sidx + 1 < plan.nseg && static_cast<int>(blockIdx.x) >= plan.seg[sidx + 1].tile0) ++sidx;
    const PackSeg& sg_ = plan.seg[sidx];
    const int t_local = blockIdx.x - sg_.tile0;
    const int tid = threadIdx.x;
    if (sg_.kind == 0) {
        const int A = sg_.A, B = sg_.B;
        const int bt = B / 32;
        const int a0 = (t_local / bt) * 16, b0 = (t_local % bt) * 32;
        // load: for each a, the 32 b's x 16 taps are 512 contiguous floats
        for (int e = tid; e < 16 * 128; e += 256) {
            const int a = e >> 7, r4 = (e & 127) * 4;
            const float4 v = __ldg(reinterpret_cast<const float4*>(sg_.src + (static_cast<long>(a0 + a) * B + b0) * 16 + r4));
            __nv_bfloat162* d = reinterpret_cast<__nv_bfloat162*>(&tile[a * kRow + r4]);
            d[0] = __floats2bfloat162_rn(v.x, v.y);
            d[1] = __floats2bfloat162_rn(v.z, v.w);
        }
        __syncthreads();
        if (sg_.ab) {  // [a][t][b]: 32 b's contiguous, written as bf16 pairs
            for (int e = tid; e < 16 * 16 * 16; e += 256) {
                const int b2 = (e & 15) * 2, t = (e >> 4) & 15, a = e >> 8;
                __nv_bfloat162 v;
                v.x = tile[a * kRow + b2 * 16 + t];
                v.y = tile[a * kRow + (b2 + 1) * 16 + t];
                *reinterpret_cast<__nv_bfloat162*>(&sg_.ab[(static_cast<long>(a0 + a) * 16 + t) * B + b0 + b2]) = v;
            }
        }
        if (sg_.ba) {  // [b][t][a]: 16 a's contiguous
            for (int e = tid; e < 32 * 16 * 8; e += 256) {
                const int a2 = (e & 7) * 2, t = (e >> 3) & 15, b = e >> 7;
                __nv_bfloat162 v;
                v.x = tile[a2 * kRow + b * 16 + t];
                v.y = tile[(a2 + 1) * kRow + b * 16 + t];
                *reinterpret_cast<__nv_bfloat162*>(&sg_.ba[(static_cast<long>(b0 + b) * 16 + t) * A + a0 + a2]) = v;
            }
        }
    } else if (sg_.kind == 1) {  // generator fc: rows permuted NCHW feature -> NHWC column, K padded
        const int C0 = sg_.A, latent = sg_.B, Kp = sg_.Kp;
        const long total = static_cast<long>(C0) * 16 * Kp;
        for (long i = static_cast<long>(t_local) * 4096 + tid; i < total && i < static_cast<long>(t_local + 1) * 4096; i += 256) {
            const int k = static_cast<int>(i % Kp);
            const int j = static_cast<int>(i / Kp);
            const int f = (j % C0) * 16 + j / C0;
            sg_.ab[i] = __float2bfloat16(k < latent ? sg_.src[static_cast<long>(f) * latent + k] : 0.f);
            if (k == 0) sg_.fdst[j] = sg_.src2[f];
        }
    } else {  // classifier weight: NCHW -> NHWC order, fp32
        const int C = sg_.A, F = C * 16;
        for (int j = t_local * 4096 + tid; j < F && j < (t_local + 1) * 4096; j += 256)
            sg_.fdst[j] = sg_.src[(j % C) * 16 + j / C];
    }
}
int pack_plan_add_w16(PackPlan& p, const float* src, bf16* ab, bf16* ba, int A, int B) {
    if (p.nseg >= 8 || A % 32 || B % 32) return -1;
    PackSeg& s = p.seg[p.nseg++];
    s = PackSeg{};
    s.src = src; s.ab = ab; s.ba = ba; s.A = A; s.B = B; s.kind = 0; s.tile0 = p.total_tiles;
    p.total_tiles += (A / 16) * (B / 32);
    return 0;
}
int pack_plan_add_fc(PackPlan& p, const float* W, const float* bias, bf16* Wp, float* biasp, int C0, int latent, int Kp) {
    if (p.nseg >= 8) return -1;
    PackSeg& s = p.seg[p.nseg++];
    s = PackSeg{};
    s.src = W; s.src2 = bias; s.ab = Wp; s.fdst = biasp; s.A = C0; s.B = latent; s.Kp = Kp; s.kind = 1; s.tile0 = p.total_tiles;
    p.total_tiles += static_cast<int>((static_cast<long>(C0) * 16 * Kp + 4095) / 4096);
    return 0;
}
int pack_plan_add_classifier(PackPlan& p, const float* w, float* wp, int C) {
    if (p.nseg >= 8) return -1;
    PackSeg& s = p.seg[p.nseg++];
    s = PackSeg{};
    s.src = w; s.fdst = wp; s.A = C; s.kind = 2; s.tile0 = p.total_tiles;
    p.total_tiles += (C * 16 + 4095) / 4096;
    return 0;
}
void pack_plan_launch(const PackPlan& p, cudaStream_t s) {
    if (p.total_tiles <= 0) return;
    note_launch();
    pack_plan_kernel<<<p.total_tiles, 256, 0, s>>>(p);
}

template <typename T>
__global__ void cast_pad_z_kernel(const float* __restrict__ z, T* __restrict__ zp, int B, int latent, int Kp) {
    const long total = static_cast<long>(B) * Kp;
    for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long>(gridDim.x) * blockDim.x) {
        const int k = static_cast<int>(i % Kp);
        const long b = i / Kp;
        zp[i] = from_f<T>(k < latent ? z[b * latent + k] : 0.f);
    }
}
template <typename T>
void cast_pad_z(const float* z, T* zp, int B, int latent, int Kp, cudaStream_t s) {
    note_launch();
    cast_pad_z_kernel<T><<<blocks_for(static_cast<long>(B) * Kp, 256), 256, 0, s>>>(z, zp, B, latent, Kp);
}
template void cast_pad_z<float>(const float*, float*, int, int, int, cudaStream_t);
template void cast_pad_z<bf16>(const float*, bf16*, int, int, int, cudaStream_t);

// ------------------------------------------------------------------------------------------------
// two-stage deterministic column reductions
// ------------------------------------------------------------------------------------------------
template <typename T, int MODE>
__global__ void __launch_bounds__(256)
col_reduce_kernel(const T* __restrict__ a, const T* __restrict__ y, const float* __restrict__ mean,
                  const float* __restrict__ rstd, const float* __restrict__ rowscale, long rows, int C,
                  int cols_per_block, long rows_per_chunk, float* __restrict__ partial) {
    __shared__ float sm[256 * 8];
    const int tpr = cols_per_block >> 3;
    const int rpi = 256 / tpr;
    const int tcol = threadIdx.x % tpr, trow = threadIdx.x / tpr;
    const int c0 = blockIdx.x * cols_per_block + tcol * 8;
    float s0[8], s1[8], mu[8], rs[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        s0[j] = 0.f;
        s1[j] = 0.f;
        mu[j] = (MODE == 1) ? mean[c0 + j] : 0.f;
        rs[j] = (MODE == 1) ? rstd[c0 + j] : 0.f;
    }
    const long r_begin = blockIdx.y * rows_per_chunk;
    long r_end = r_begin + rows_per_chunk;
    if (r_end > rows) r_end = rows;
    // 2-4 rows per iteration with all their loads issued before the first use (one 16-byte load in flight per thread
    // left these reductions at 3-4.7 TB/s)
    using Raw = typename RawOf<T>::type;
    auto accum = [&](const Raw& ra, const Raw& ry, float w) {
        float v[8];
        unpack8(ra, v);
        if (MODE == 0) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                s0[j] += v[j];
                s1[j] = fmaf(v[j], v[j], s1[j]);
            }
        } else if (MODE == 1) {
            float yv[8];
            unpack8(ry, yv);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                s0[j] += v[j];
                s1[j] = fmaf(v[j], (yv[j] - mu[j]) * rs[j], s1[j]);
            }
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) s0[j] = fmaf(v[j], w, s0[j]);
        }
    };
    constexpr int U = MODE == 1 ? 2 : 4;  // rows in flight (mode 1 loads two tensors per row; 4 there cost occupancy)
    long r = r_begin + trow;
    for (; r + static_cast<long>(U - 1) * rpi < r_end; r += static_cast<long>(U) * rpi) {
        Raw ra[U], ry[U];
        float w[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            ra[u] = ldraw8(a + (r + static_cast<long>(u) * rpi) * C + c0);
            if (MODE == 1) ry[u] = ldraw8(y + (r + static_cast<long>(u) * rpi) * C + c0);
            w[u] = MODE == 2 ? rowscale[r + static_cast<long>(u) * rpi] : 0.f;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) accum(ra[u], MODE == 1 ? ry[u] : ra[u], w[u]);
    }
    for (; r < r_end; r += rpi) {
        const Raw ra = ldraw8(a + r * C + c0);
        const Raw ry = MODE == 1 ? ldraw8(y + r * C + c0) : ra;
        accum(ra, ry, MODE == 2 ? rowscale[r] : 0.f);
    }
    float* out = partial + static_cast<long>(blockIdx.y) * 2 * C + c0;
#pragma unroll
    for (int which = 0; which < 2; ++which) {
        if (which == 1 && MODE == 2) break;
        __syncthreads();
#pragma unroll
        for (int j = 0; j < 8; ++j) sm[threadIdx.x * 8 + j] = which == 0 ? s0[j] : s1[j];
        __syncthreads();
        if (trow == 0) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float acc = 0.f;
                for (int t = 0; t < rpi; ++t) acc += sm[(t * tpr + tcol) * 8 + j];
                out[which * C + j] = acc;
            }
        }
    }
}

template <typename T>
int col_reduce(int mode, const T* a, const T* y, const float* mean, const float* rstd, const float* rowscale,
               long rows, int C, float* partial, cudaStream_t s) {
    const int cpb = C > 2048 ? 2048 : C;
    const int col_blocks = C / cpb;
    const int rpi = 256 / (cpb / 8);
    long chunks = (rows + static_cast<long>(rpi) * 8 - 1) / (static_cast<long>(rpi) * 8);
    const long cap = kMaxChunks / col_blocks > 0 ? kMaxChunks / col_blocks : 1;
    if (chunks > cap) chunks = cap;
    if (chunks < 1) chunks = 1;
    long rpc = (rows + chunks - 1) / chunks;
    rpc = (rpc + rpi - 1) / rpi * rpi;
    chunks = (rows + rpc - 1) / rpc;
    dim3 grid(col_blocks, static_cast<unsigned>(chunks));
    note_launch();
    if (mode == 0)
        col_reduce_kernel<T, 0><<<grid, 256, 0, s>>>(a, y, mean, rstd, rowscale, rows, C, cpb, rpc, partial);
    else if (mode == 1)
        col_reduce_kernel<T, 1><<<grid, 256, 0, s>>>(a, y, mean, rstd, rowscale, rows, C, cpb, rpc, partial);
    else
        col_reduce_kernel<T, 2><<<grid, 256, 0, s>>>(a, y, mean, rstd, rowscale, rows, C, cpb, rpc, partial);
    return static_cast<int>(chunks);
}
template int col_reduce<float>(int, const float*, const float*, const float*, const float*, const float*, long, int,
                               float*, cudaStream_t);
template int col_reduce<bf16>(int, const bf16*, const bf16*, const float*, const float*, const float*, long, int,
                              float*, cudaStream_t);

__device__ __forceinline__ int perm_index(int j, int perm_c0) {
    return perm_c0 > 0 ? (j % perm_c0) * 16 + j / perm_c0 : j;
}

// The finalize kernels fold `chunks` (up to 592) partial rows per output element. One WARP per element: lanes stride
// over the chunks (32 independent loads in flight) and the lane sums are folded in a fixed butterfly order, so the
// result is deterministic. (One thread per element walking the rows serially took 20-65 us per launch for the
// 32-channel layers — ~20 such launches per training step.)
constexpr int kFinThreads = 256;
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ void chunk_sums(const float* __restrict__ partial, int chunks, long stride, long off0,
                                           long off1, bool two, double& s0, double& s1) {
    const int lane = threadIdx.x & 31;
    double a = 0.0, b = 0.0;
    for (int k = lane; k < chunks; k += 32) {
        a += partial[k * stride + off0];
        if (two) b += partial[k * stride + off1];
    }
    s0 = warp_sum(a);
    s1 = two ? warp_sum(b) : 0.0;
}
inline int fin_blocks(int elems) { return (elems + kFinThreads / 32 - 1) / (kFinThreads / 32); }

__global__ void bn_finalize_kernel(const float* __restrict__ partial, int chunks, long rows, int C,
                                   const float* __restrict__ gamma, const float* __restrict__ beta,
                                   float* __restrict__ running_mean, float* __restrict__ running_var, float momentum,
                                   float eps, int batch_stats, int perm_c0, float* __restrict__ mean,
                                   float* __restrict__ rstd, float* __restrict__ scale, float* __restrict__ shift) {
    const int j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;  // one warp per channel
    if (j >= C) return;
    const int p = perm_index(j, perm_c0);
    float mu, var;
    double s0 = 0.0, s1 = 0.0;
    if (batch_stats) chunk_sums(partial, chunks, 2L * C, j, C + j, true, s0, s1);
    if ((threadIdx.x & 31) != 0) return;
    if (batch_stats) {
        const double m = s0 / rows;
        double v = s1 / rows - m * m;
        if (v < 0.0) v = 0.0;
        mu = static_cast<float>(m);
        var = static_cast<float>(v);
        const double unbiased = rows > 1 ? v * rows / (rows - 1) : v;
        running_mean[p] = (1.f - momentum) * running_mean[p] + momentum * mu;
        running_var[p] = (1.f - momentum) * running_var[p] + momentum * static_cast<float>(unbiased);
    } else {
        mu = running_mean[p];
        var = running_var[p];
    }
    const float r = 1.0f / sqrtf(var + eps);
    mean[j] = mu;
    rstd[j] = r;
    const float sc = gamma[p] * r;
    scale[j] = sc;
    shift[j] = beta[p] - mu * sc;
}
// Eval mode: the coefficients of ALL BatchNorm layers of the Generator from their running statistics in one launch
// (one thread per channel; the per-layer launches were 5-6 tiny kernels in front of every eval forward).
__global__ void bn_eval_all_kernel(const BnEvalPlan plan, float eps) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= plan.total) return;
    int l = 0;
    while (l + 1 < plan.n && i >= plan.layer[l + 1].first) ++l;
    const BnEvalLayer& L = plan.layer[l];
    const int j = i - L.first;
    const int p = perm_index(j, L.perm_c0);
    const float mu = L.running_mean[p];
    const float r = 1.0f / sqrtf(L.running_var[p] + eps);
    L.mean[j] = mu;
    L.rstd[j] = r;
    const float sc = L.gamma[p] * r;
    L.scale[j] = sc;
    L.shift[j] = L.beta[p] - mu * sc;
}
void bn_eval_all(const BnEvalPlan& plan, float eps, cudaStream_t s) {
    if (plan.total <= 0) return;
    note_launch();
    bn_eval_all_kernel<<<(plan.total + 255) / 256, 256, 0, s>>>(plan, eps);
}

void bn_finalize(const float* partial, int chunks, long rows, int C, const float* gamma, const float* beta,
                 float* running_mean, float* running_var, float momentum, float eps, int batch_stats, int perm_c0,
                 float* mean, float* rstd, float* scale, float* shift, cudaStream_t s) {
    note_launch();
    bn_finalize_kernel<<<fin_blocks(C), kFinThreads, 0, s>>>(partial, chunks, rows, C, gamma, beta, running_mean,
                                                      running_var, momentum, eps, batch_stats, perm_c0, mean, rstd,
                                                      scale, shift);
}

// Per-channel vectors are read as two float4 per 8 channels: with scalar loads these kernels issued 16-40 LDG per
// 16-byte activation load and sat on the load/store queue (ncu: stall_lg at the first vector load).
__device__ __forceinline__ void ldvec8(const float* __restrict__ p, float (&v)[8]) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p));
    const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
    v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}

// kHoist: the grid stride (in elements) is a multiple of C, so a thread meets the same 8 channels in every iteration and
// loads their coefficients once — with the per-iteration vector loads these element-wise kernels moved several times more
// bytes through the L1 than through HBM (ncu: l1tex 60-89 % busy at 3.5-5.5 TB/s of DRAM traffic).
template <typename T, bool kHoist>
__global__ void bn_apply_relu_kernel(const T* __restrict__ y, const float* __restrict__ scale,
                                     const float* __restrict__ shift, T* __restrict__ a, long n8, int C,
                                     float act_slope) {
    const long i0 = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x;
    float sc[8], sh[8];
    if (kHoist) {
        const int c0 = static_cast<int>((i0 * 8) & (C - 1));
        ldvec8(scale + c0, sc);
        ldvec8(shift + c0, sh);
    }
    for (long i = i0; i < n8; i += static_cast<long>(gridDim.x) * blockDim.x) {
        float v[8];
        load8(y + i * 8, v);
        if (!kHoist) {
            const int c0 = static_cast<int>((i * 8) & (C - 1));
            ldvec8(scale + c0, sc);
            ldvec8(shift + c0, sh);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float q = fmaf(v[j], sc[j], sh[j]);
            v[j] = fmaxf(q, 0.f) + act_slope * fminf(q, 0.f);  // slope 0: ReLU (+0 for negative inputs)
        }
        store8(a + i * 8, v);
    }
}
template <typename T>
void bn_apply_relu(const T* y, const float* scale, const float* shift, T* a, long rows, int C, float act_slope,
                   cudaStream_t s) {
    const long n8 = rows * C / 8;
    const int blocks = blocks_for(n8, 256);
    note_launch();
    if ((static_cast<long>(blocks) * 256 * 8) % C == 0)
        bn_apply_relu_kernel<T, true><<<blocks, 256, 0, s>>>(y, scale, shift, a, n8, C, act_slope);
    else
        bn_apply_relu_kernel<T, false><<<blocks, 256, 0, s>>>(y, scale, shift, a, n8, C, act_slope);
}
template void bn_apply_relu<float>(const float*, const float*, const float*, float*, long, int, float, cudaStream_t);
template void bn_apply_relu<bf16>(const bf16*, const float*, const float*, bf16*, long, int, float, cudaStream_t);

__global__ void bn_bwd_finalize_kernel(const float* __restrict__ partial, int chunks, long rows, int C,
                                       const float* __restrict__ gamma, const float* __restrict__ rstd,
                                       const float* __restrict__ raw_mean, int batch_stats, int perm_c0,
                                       float* __restrict__ dgamma,
                                       float* __restrict__ dbeta, float* __restrict__ k1, float* __restrict__ k2,
                                       float* __restrict__ k3) {
    const int j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;  // one warp per channel
    if (j >= C) return;
    const int p = perm_index(j, perm_c0);
    double s0, s1;
    chunk_sums(partial, chunks, 2L * C, j, C + j, true, s0, s1);
    if ((threadIdx.x & 31) != 0) return;
    // raw mode: the second partial is sum d*y; sum d*xhat = rstd * (sum d*y - mean * sum d)
    if (raw_mean) s1 = static_cast<double>(rstd[j]) * (s1 - static_cast<double>(raw_mean[j]) * s0);
    dbeta[p] = static_cast<float>(s0);
    dgamma[p] = static_cast<float>(s1);
    k1[j] = gamma[p] * rstd[j];
    k2[j] = batch_stats ? static_cast<float>(s0 / rows) : 0.f;
    k3[j] = batch_stats ? static_cast<float>(s1 / rows) : 0.f;
}
void bn_bwd_finalize(const float* partial, int chunks, long rows, int C, const float* gamma, const float* rstd,
                     const float* raw_mean, int batch_stats, int perm_c0, float* dgamma, float* dbeta, float* k1,
                     float* k2, float* k3, cudaStream_t s) {
    note_launch();
    bn_bwd_finalize_kernel<<<fin_blocks(C), kFinThreads, 0, s>>>(partial, chunks, rows, C, gamma, rstd, raw_mean, batch_stats,
                                                          perm_c0, dgamma, dbeta, k1, k2, k3);
}

template <typename T, bool kHoist>
__global__ void bn_bwd_apply_kernel(const T* __restrict__ d, const T* __restrict__ y, const float* __restrict__ mean,
                                    const float* __restrict__ rstd, const float* __restrict__ k1,
                                    const float* __restrict__ k2, const float* __restrict__ k3, T* __restrict__ dy,
                                    long n8, int C) {
    const long i0 = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x;
    float mu[8], rs[8], a1[8], a2[8], a3[8];
    auto coefficients = [&](long i) {
        const int c0 = static_cast<int>((i * 8) & (C - 1));
        ldvec8(mean + c0, mu);
        ldvec8(rstd + c0, rs);
        ldvec8(k1 + c0, a1);
        ldvec8(k2 + c0, a2);
        ldvec8(k3 + c0, a3);
    };
    if (kHoist) coefficients(i0);  // see bn_apply_relu_kernel
    for (long i = i0; i < n8; i += static_cast<long>(gridDim.x) * blockDim.x) {
        float dv[8], yv[8];
        load8(d + i * 8, dv);
        load8(y + i * 8, yv);
        if (!kHoist) coefficients(i);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float xhat = (yv[j] - mu[j]) * rs[j];
            dv[j] = a1[j] * (dv[j] - a2[j] - xhat * a3[j]);
        }
        store8(dy + i * 8, dv);
    }
}
template <typename T>
void bn_bwd_apply(const T* d, const T* y, const float* mean, const float* rstd, const float* k1, const float* k2,
                  const float* k3, T* dy, long rows, int C, cudaStream_t s) {
    const long n8 = rows * C / 8;
    const int blocks = blocks_for(n8, 256);
    note_launch();
    if ((static_cast<long>(blocks) * 256 * 8) % C == 0)
        bn_bwd_apply_kernel<T, true><<<blocks, 256, 0, s>>>(d, y, mean, rstd, k1, k2, k3, dy, n8, C);
    else
        bn_bwd_apply_kernel<T, false><<<blocks, 256, 0, s>>>(d, y, mean, rstd, k1, k2, k3, dy, n8, C);
}
template void bn_bwd_apply<float>(const float*, const float*, const float*, const float*, const float*, const float*,
                                  const float*, float*, long, int, cudaStream_t);
template void bn_bwd_apply<bf16>(const bf16*, const bf16*, const float*, const float*, const float*, const float*,
                                 const float*, bf16*, long, int, cudaStream_t);

__global__ void col_finalize_kernel(const float* __restrict__ partial, int chunks, int C, int perm_c0,
                                    float* __restrict__ out0, float* __restrict__ out1) {
    const int j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;  // one warp per column
    if (j >= C) return;
    const int p = perm_index(j, perm_c0);
    double s0, s1;
    chunk_sums(partial, chunks, 2L * C, j, C + j, out1 != nullptr, s0, s1);
    if ((threadIdx.x & 31) != 0) return;
    if (out0) out0[p] = static_cast<float>(s0);
    if (out1) out1[p] = static_cast<float>(s1);
}
void col_finalize(const float* partial, int chunks, int C, int perm_c0, float* out0, float* out1, cudaStream_t s) {
    note_launch();
    col_finalize_kernel<<<fin_blocks(C), kFinThreads, 0, s>>>(partial, chunks, C, perm_c0, out0, out1);
}

// ------------------------------------------------------------------------------------------------
// partial[chunks][n] -> out_a[0..na) and out_b[0..n-na) (deterministic, double accumulation)
// ------------------------------------------------------------------------------------------------
__global__ void vec_finalize_kernel(const float* __restrict__ partial, int chunks, int n, float* __restrict__ out_a,
                                    int na, float* __restrict__ out_b) {
    const int j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;  // one warp per element
    if (j >= n) return;
    double s, unused;
    chunk_sums(partial, chunks, n, j, 0, false, s, unused);
    if ((threadIdx.x & 31) != 0) return;
    if (j < na)
        out_a[j] = static_cast<float>(s);
    else if (out_b)
        out_b[j - na] = static_cast<float>(s);
}

void vec_finalize(const float* partial, int chunks, int n, float* out_a, int na, float* out_b, cudaStream_t s) {
    note_launch();
    vec_finalize_kernel<<<fin_blocks(n), kFinThreads, 0, s>>>(partial, chunks, n, out_a, na, out_b);
}

// ------------------------------------------------------------------------------------------------
// Discriminator head: Conv 4x4 s2 p1 from the 1-channel image, forward / wgrad / dgrad
// ------------------------------------------------------------------------------------------------
// Forward: a thread computes 4 consecutive output pixels x 8 channels from a 4x10 window of the image held in
// registers; the 16x8 filter slice is read from shared memory once per 4 pixels (1 LDS.128 per 16 FMAs).
template <typename T>
__global__ void __launch_bounds__(256)
d_conv0_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
               const float* __restrict__ mask, float slope, T* __restrict__ a, int B, int S, int C) {
    extern __shared__ float wsm[];  // [16][C] then bias[C]
    for (int i = threadIdx.x; i < C * 16; i += blockDim.x) wsm[(i % 16) * C + i / 16] = w[i];
    for (int i = threadIdx.x; i < C; i += blockDim.x) wsm[16 * C + i] = bias[i];
    __syncthreads();
    const int O = S / 2, g = C / 8, Q = O / 4;  // Q pixel-quads per output row
    const int lg = ilog2(g), lq = ilog2(Q), lo = ilog2(O);
    const long total = static_cast<long>(B) * O * Q * g;
    for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long>(gridDim.x) * blockDim.x) {
        const int c0 = (static_cast<int>(i) & (g - 1)) * 8;
        const long quad = i >> lg;
        const int ox0 = (static_cast<int>(quad) & (Q - 1)) * 4;
        const int oy = static_cast<int>(quad >> lq) & (O - 1);
        const long n = quad >> (lq + lo);
        float xv[4][10];
#pragma unroll
        for (int ky = 0; ky < 4; ++ky) {
            const int iy = 2 * oy - 1 + ky;
            const bool yok = iy >= 0 && iy < S;
            const float* row = x + (n * S + (yok ? iy : 0)) * S;
#pragma unroll
            for (int k = 0; k < 10; ++k) {
                const int ix = 2 * ox0 - 1 + k;
                xv[ky][k] = (yok && ix >= 0 && ix < S) ? __ldg(row + ix) : 0.f;
            }
        }
        float acc[4][8];
#pragma unroll
        for (int p = 0; p < 4; ++p)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[p][j] = wsm[16 * C + c0 + j];
#pragma unroll
        for (int ky = 0; ky < 4; ++ky)
#pragma unroll
            for (int kx = 0; kx < 4; ++kx) {
                const float4 w0 = *reinterpret_cast<const float4*>(wsm + (ky * 4 + kx) * C + c0);
                const float4 w1 = *reinterpret_cast<const float4*>(wsm + (ky * 4 + kx) * C + c0 + 4);
                const float ww[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
                for (int p = 0; p < 4; ++p) {
                    const float xvv = xv[ky][2 * p + kx];
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[p][j] = fmaf(xvv, ww[j], acc[p][j]);
                }
            }
        float mk[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) mk[j] = mask ? mask[n * C + c0 + j] : 1.f;
        const long pix0 = (n * O + oy) * O + ox0;
#pragma unroll
        for (int p = 0; p < 4; ++p) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float v = acc[p][j] > 0.f ? acc[p][j] : acc[p][j] * slope;
                acc[p][j] = v * mk[j];
            }
            store8(a + (pix0 + p) * C + c0, acc[p]);
        }
    }
}
template <typename T>
void d_conv0(const float* x, const float* w, const float* bias, const float* mask, float slope, T* a, int B, int S,
             int C, cudaStream_t s) {
    if constexpr (std::is_same<T, bf16>::value) {
        if (C == 64 || C == 128) {  // 64 channels per launch inside the C-channel NHWC tensor
            for (int c0 = 0; c0 < C; c0 += 64)
                dconv0_fwd_mma(x, w + c0 * 16, bias + c0, mask ? mask + c0 : nullptr, slope, a + c0, B, S, s, C);
            return;
        }
    }
    const long total = static_cast<long>(B) * (S / 2) * (S / 8) * (C / 8);
    note_launch();
    d_conv0_kernel<T><<<blocks_for(total, 256, 148 * 16), 256, 17 * C * sizeof(float), s>>>(x, w, bias, mask, slope, a,
                                                                                           B, S, C);
}
template void d_conv0<float>(const float*, const float*, const float*, const float*, float, float*, int, int, int,
                             cudaStream_t);
template void d_conv0<bf16>(const float*, const float*, const float*, const float*, float, bf16*, int, int, int,
                            cudaStream_t);

// Weight gradient: 16 lanes per output pixel = 8 channel slices x 2 filter-row halves; each lane keeps its
// 8 channels x (2 ky x 4 kx) accumulators (+ the bias gradient) in registers while striding over pixels.
// partial[block][C*16 + C] with C == 64.
template <typename T>
__global__ void __launch_bounds__(256)
d_conv0_wgrad_kernel(const float* __restrict__ x, const T* __restrict__ dy, float* __restrict__ partial, int B, int S) {
    constexpr int C = 64;
    __shared__ float sm[8][C * 17];
    const int sub = threadIdx.x & 15;
    const int c0 = (sub & 7) * 8, half = sub >> 3;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int O = S / 2;
    float acc[2][4][8], bacc[8];
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[r][k][j] = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) bacc[j] = 0.f;
    const long total = static_cast<long>(B) * O * O;  // multiple of 2: both pixel slots of a warp stay in the loop
    const long stride = static_cast<long>(gridDim.x) * blockDim.x / 16;
    const int lo = ilog2(O);
    for (long pix = (blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x) / 16; pix < total; pix += stride) {
        const int ox = static_cast<int>(pix) & (O - 1);
        const int oy = static_cast<int>(pix >> lo) & (O - 1);
        const long n = pix >> (2 * lo);
        const typename RawOf<T>::type draw = ldraw8(dy + pix * C + c0);
        float xv[2][4];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int iy = 2 * oy - 1 + half * 2 + r;
            const int cy = min(max(iy, 0), S - 1);
            const float* row = x + (n * S + cy) * S;
#pragma unroll
            for (int kx = 0; kx < 4; ++kx) {
                const int ix = 2 * ox - 1 + kx;
                const int cx = min(max(ix, 0), S - 1);
                const float v = __ldg(row + cx);
                xv[r][kx] = (iy == cy && ix == cx) ? v : 0.f;
            }
        }
        float d[8];
        unpack8(draw, d);
        if (half == 0) {
#pragma unroll
            for (int j = 0; j < 8; ++j) bacc[j] += d[j];
        }
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int kx = 0; kx < 4; ++kx)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[r][kx][j] = fmaf(d[j], xv[r][kx], acc[r][kx][j]);
    }
    // the two pixel slots of a warp (lane and lane^16) hold the same outputs
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[r][k][j] += __shfl_xor_sync(0xffffffffu, acc[r][k][j], 16);
#pragma unroll
    for (int j = 0; j < 8; ++j) bacc[j] += __shfl_xor_sync(0xffffffffu, bacc[j], 16);
    if (lane < 16) {
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int k = 0; k < 4; ++k)
#pragma unroll
                for (int j = 0; j < 8; ++j) sm[warp][(c0 + j) * 16 + (half * 2 + r) * 4 + k] = acc[r][k][j];
        if (half == 0) {
#pragma unroll
            for (int j = 0; j < 8; ++j) sm[warp][C * 16 + c0 + j] = bacc[j];
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < C * 17; i += blockDim.x) {
        float s = 0.f;
#pragma unroll
        for (int wv = 0; wv < 8; ++wv) s += sm[wv][i];
        partial[static_cast<long>(blockIdx.x) * (C * 17) + i] = s;
    }
}
template <typename T>
void d_conv0_wgrad(const float* x, const T* dy, float* dW, float* partial, int B, int S, int C, cudaStream_t s) {
    if constexpr (std::is_same<T, bf16>::value) {
        if (C == 128) {  // 2x-width variant: one launch per 64-channel half; dW (C*16) is followed by dbias (C)
            for (int c0 = 0; c0 < C; c0 += 64) {
                const int chunks = dconv0_wgrad_mma(x, dy + c0, partial, B, S, s, C);
                vec_finalize(partial, chunks, 64 * 17, dW + c0 * 16, 64 * 16, dW + C * 16 + c0, s);
            }
            return;
        }
    }
    if (C != 64) {
        note_unsupported("d_conv0_wgrad", C);
        return;
    }
    const long total = static_cast<long>(B) * (S / 2) * (S / 2);
    int chunks = blocks_for(total * 16, 256, kMaxChunks);
    if constexpr (std::is_same<T, bf16>::value) {
        chunks = dconv0_wgrad_mma(x, dy, partial, B, S, s);
    } else {
        note_launch();
        d_conv0_wgrad_kernel<T><<<chunks, 256, 0, s>>>(x, dy, partial, B, S);
    }
    const int n = C * 17;
    // dW (C*16 floats) is immediately followed by dbias (C floats) in the flat gradient buffer
    vec_finalize(partial, chunks, n, dW, n, nullptr, s);
}
template void d_conv0_wgrad<float>(const float*, const float*, float*, float*, int, int, int, cudaStream_t);
template void d_conv0_wgrad<bf16>(const float*, const bf16*, float*, float*, int, int, int, cudaStream_t);

// Image gradient: dx[n,iy,ix] = sum over the 2x2 output pixels that see (iy,ix) and all channels of
// dy[n,oy,ox,c] * w[c][ky][kx]. blockIdx.y selects the (iy&1, ix&1) parity class so every thread of a block uses
// the same 4 filter taps; 4 lanes per pixel hold 16 channels x 4 taps of weights in registers.
template <typename T>
__global__ void __launch_bounds__(256)
d_conv0_dgrad_kernel(const T* __restrict__ dy, const float* __restrict__ w, float* __restrict__ dx, int B, int S) {
    constexpr int C = 64;
    const int py = blockIdx.y >> 1, px = blockIdx.y & 1;
    const int q = threadIdx.x & 3;
    const int O = S / 2;
    float wr[2][2][16];
    int kyv[2], kxv[2];
#pragma unroll
    for (int t = 0; t < 2; ++t) {
        kyv[t] = ((py + 1) & 1) + 2 * t;
        kxv[t] = ((px + 1) & 1) + 2 * t;
    }
#pragma unroll
    for (int ty = 0; ty < 2; ++ty)
#pragma unroll
        for (int tx = 0; tx < 2; ++tx)
#pragma unroll
            for (int j = 0; j < 16; ++j) wr[ty][tx][j] = w[(q * 16 + j) * 16 + kyv[ty] * 4 + kxv[tx]];
    const long total = static_cast<long>(B) * O * O;  // pixels of this parity class
    const long stride = static_cast<long>(gridDim.x) * blockDim.x / 4;
    const int lo = ilog2(O);
    for (long i = (blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x) / 4; i < total; i += stride) {
        const int xh = static_cast<int>(i) & (O - 1);
        const int yh = static_cast<int>(i >> lo) & (O - 1);
        const long n = i >> (2 * lo);
        const int iy = 2 * yh + py, ix = 2 * xh + px;
        float acc = 0.f;
        typename RawOf<T>::type raw[2][2][2];
        bool ok[2][2];
#pragma unroll
        for (int ty = 0; ty < 2; ++ty) {
            const int oy = (iy + 1 - kyv[ty]) >> 1;   // arithmetic shift: -1 for the out-of-range row above the image
            const int cy = min(max(oy, 0), O - 1);
#pragma unroll
            for (int tx = 0; tx < 2; ++tx) {
                const int ox = (ix + 1 - kxv[tx]) >> 1;
                const int cx = min(max(ox, 0), O - 1);
                ok[ty][tx] = (oy == cy) && (ox == cx);
                const T* p = dy + ((n * O + cy) * O + cx) * C + q * 16;
                raw[ty][tx][0] = ldraw8(p);
                raw[ty][tx][1] = ldraw8(p + 8);
            }
        }
#pragma unroll
        for (int ty = 0; ty < 2; ++ty)
#pragma unroll
            for (int tx = 0; tx < 2; ++tx) {
                float v[8], part = 0.f;
                unpack8(raw[ty][tx][0], v);
#pragma unroll
                for (int j = 0; j < 8; ++j) part = fmaf(v[j], wr[ty][tx][j], part);
                unpack8(raw[ty][tx][1], v);
#pragma unroll
                for (int j = 0; j < 8; ++j) part = fmaf(v[j], wr[ty][tx][8 + j], part);
                acc += ok[ty][tx] ? part : 0.f;
            }
        acc += __shfl_xor_sync(0xffffffffu, acc, 1);
        acc += __shfl_xor_sync(0xffffffffu, acc, 2);
        if (q == 0) dx[(n * S + iy) * S + ix] = acc;
    }
}
template <typename T>
void d_conv0_dgrad(const T* dy, const float* w, float* dx, int B, int S, int C, cudaStream_t s) {
    if constexpr (std::is_same<T, bf16>::value) {
        if (C == 64 || C == 128) {
            for (int c0 = 0; c0 < C; c0 += 64) dconv0_dgrad_mma(dy + c0, w + c0 * 16, dx, B, S, s, C, c0 > 0 ? 1 : 0);
            return;
        }
    }
    if (C != 64) {
        note_unsupported("d_conv0_dgrad", C);
        return;
    }
    const long threads = static_cast<long>(B) * (S / 2) * (S / 2) * 4;
    dim3 grid(blocks_for(threads, 256, 148 * 4), 4);
    note_launch();
    d_conv0_dgrad_kernel<T><<<grid, 256, 0, s>>>(dy, w, dx, B, S);
}
template void d_conv0_dgrad<float>(const float*, const float*, float*, int, int, int, cudaStream_t);
template void d_conv0_dgrad<bf16>(const bf16*, const float*, float*, int, int, int, cudaStream_t);

// ------------------------------------------------------------------------------------------------
// classifier + sigmoid, and backward pieces
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void classifier_sigmoid_kernel(const T* __restrict__ a, const float* __restrict__ wp,
                                          const float* __restrict__ bias, float* __restrict__ prob, int B, int F) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= B) return;
    const T* row = a + static_cast<long>(warp) * F;
    float acc = 0.f;
#pragma unroll 4
    for (int j = lane * 8; j < F; j += 256) {
        float v[8], w[8];
        load8(row + j, v);
        ldvec8(wp + j, w);
#pragma unroll
        for (int k = 0; k < 8; ++k) acc = fmaf(v[k], w[k], acc);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) {
        const float logit = acc + bias[0];
        prob[warp] = 1.f / (1.f + expf(-logit));
    }
}
template <typename T>
void classifier_sigmoid(const T* a, const float* wp, const float* bias, float* prob, int B, int F, cudaStream_t s) {
    note_launch();
    classifier_sigmoid_kernel<T><<<(B * 32 + 255) / 256, 256, 0, s>>>(a, wp, bias, prob, B, F);
}
template void classifier_sigmoid<float>(const float*, const float*, const float*, float*, int, int, cudaStream_t);
template void classifier_sigmoid<bf16>(const bf16*, const float*, const float*, float*, int, int, cudaStream_t);

template <typename T>
__global__ void features_nchw_kernel(const T* __restrict__ a, float* __restrict__ feat, int B, int C) {
    const int F = C * 16;
    const long total = static_cast<long>(B) * F;
    for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long>(gridDim.x) * blockDim.x) {
        const int f = static_cast<int>(i % F);
        const long b = i / F;
        const int c = f / 16, hw = f % 16;
        feat[i] = to_f(a[b * F + hw * C + c]);
    }
}
template <typename T>
void features_nchw(const T* a, float* feat, int B, int C, cudaStream_t s) {
    note_launch();
    features_nchw_kernel<T><<<blocks_for(static_cast<long>(B) * C * 16, 256), 256, 0, s>>>(a, feat, B, C);
}
template void features_nchw<float>(const float*, float*, int, int, cudaStream_t);
template void features_nchw<bf16>(const bf16*, float*, int, int, cudaStream_t);

__global__ void sigmoid_bwd_kernel(const float* __restrict__ prob, const float* __restrict__ dprob,
                                   float* __restrict__ dlogit, int B) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < B) {
        const float p = prob[i];
        dlogit[i] = dprob[i] * p * (1.f - p);
    }
}
void sigmoid_bwd(const float* prob, const float* dprob, float* dlogit, int B, cudaStream_t s) {
    note_launch();
    sigmoid_bwd_kernel<<<(B + 255) / 256, 256, 0, s>>>(prob, dprob, dlogit, B);
}

template <typename T>
__global__ void classifier_bwd_dy_kernel(const float* __restrict__ dlogit, const float* __restrict__ wp,
                                         const float* __restrict__ mask, const T* __restrict__ a, float slope,
                                         T* __restrict__ dy, int B, int C) {
    const int F = C * 16;
    const long n8 = static_cast<long>(B) * F / 8;
    for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < n8;
         i += static_cast<long>(gridDim.x) * blockDim.x) {
        const int j0 = static_cast<int>((i * 8) & (F - 1));
        const long b = (i * 8) >> ilog2(F);
        const int c0 = j0 & (C - 1);
        const float dl = dlogit[b];
        float av[8], w[8], mk[8];
        load8(a + i * 8, av);
        ldvec8(wp + j0, w);
        if (mask) ldvec8(mask + b * C + c0, mk);  // (8 scalar mask loads per 16-byte store made this issue-bound)
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            float v = dl * w[k] * (av[k] > 0.f ? 1.f : slope);
            if (mask) v *= mk[k];
            av[k] = v;
        }
        store8(dy + i * 8, av);
    }
}
template <typename T>
void classifier_bwd_dy(const float* dlogit, const float* wp, const float* mask, const T* a, float slope, T* dy, int B,
                       int C, cudaStream_t s) {
    const long n8 = static_cast<long>(B) * C * 16 / 8;
    note_launch();
    classifier_bwd_dy_kernel<T><<<blocks_for(n8, 256), 256, 0, s>>>(dlogit, wp, mask, a, slope, dy, B, C);
}
template void classifier_bwd_dy<float>(const float*, const float*, const float*, const float*, float, float*, int, int,
                                       cudaStream_t);
template void classifier_bwd_dy<bf16>(const float*, const float*, const float*, const bf16*, float, bf16*, int, int,
                                      cudaStream_t);

// single-block deterministic reductions -----------------------------------------------------------
template <int N>
__device__ __forceinline__ void block_sum(float (&v)[N], float* sm /*[32*N]*/) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
    for (int k = 0; k < N; ++k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
    }
    __syncthreads();
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < N; ++k) sm[warp * N + k] = v[k];
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int k = 0; k < N; ++k) {
            float t = lane < nw ? sm[lane * N + k] : 0.f;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
            v[k] = t;
        }
    }
}

__global__ void sum_vector_kernel(const float* __restrict__ v, int n, float* __restrict__ out) {
    __shared__ float sm[32];
    float acc[1] = {0.f};
    for (int i = threadIdx.x; i < n; i += blockDim.x) acc[0] += v[i];
    block_sum<1>(acc, sm);
    if (threadIdx.x == 0) out[0] = acc[0];
}
void sum_vector(const float* v, int n, float* out, cudaStream_t s) { sum_vector_kernel<<<1, 1024, 0, s>>>(v, n, out); }

__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
__global__ void dropout_masks_kernel(uint64_t seed, uint64_t offset, long n, float p, float* __restrict__ out) {
    const float keep = 1.f / (1.f - p);
    for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<long>(gridDim.x) * blockDim.x) {
        const uint64_t h = splitmix64(splitmix64(seed) ^ (offset + static_cast<uint64_t>(i)));
        const float u = static_cast<float>(h >> 40) * (1.0f / 16777216.0f);
        out[i] = u >= p ? keep : 0.f;
    }
}
void dropout_masks(uint64_t seed, uint64_t offset, long n, float p, float* out, cudaStream_t s) {
    note_launch();
    dropout_masks_kernel<<<blocks_for(n, 256), 256, 0, s>>>(seed, offset, n, p, out);
}
__global__ void dropout_masks_dev_kernel(uint64_t seed, const unsigned long long* __restrict__ offset_ptr, long n, float p,
                                         float* __restrict__ out) {
    const float keep = 1.f / (1.f - p);
    const uint64_t offset = *offset_ptr;
    for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<long>(gridDim.x) * blockDim.x) {
        const uint64_t h = splitmix64(splitmix64(seed) ^ (offset + static_cast<uint64_t>(i)));
        const float u = static_cast<float>(h >> 40) * (1.0f / 16777216.0f);
        out[i] = u >= p ? keep : 0.f;
    }
}
void dropout_masks_dev(uint64_t seed, const unsigned long long* offset_ptr, long n, float p, float* out, cudaStream_t s) {
    note_launch();
    dropout_masks_dev_kernel<<<blocks_for(n, 256), 256, 0, s>>>(seed, offset_ptr, n, p, out);
}

// ------------------------------------------------------------------------------------------------
// BCE on probabilities
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float bce_term(float p, float y) {
    const float lp = fmaxf(logf(p), -100.f);
    const float l1p = fmaxf(logf(1.f - p), -100.f);
    return -(y * lp + (1.f - y) * l1p);
}
__device__ __forceinline__ float bce_dp(float p, float y, float inv_n) {
    return (p - y) / fmaxf(p * (1.f - p), 1e-12f) * inv_n;
}

__global__ void bce_forward_kernel(const float* __restrict__ prob, const float* __restrict__ target, int n,
                                   float* __restrict__ loss) {
    __shared__ float sm[32];
    float acc[1] = {0.f};
    for (int i = threadIdx.x; i < n; i += blockDim.x) acc[0] += bce_term(prob[i], target[i]);
    block_sum<1>(acc, sm);
    if (threadIdx.x == 0) loss[0] = acc[0] / n;
}
void bce_forward(const float* prob, const float* target, int n, float* loss, cudaStream_t s) {
    note_launch();
    bce_forward_kernel<<<1, 1024, 0, s>>>(prob, target, n, loss);
}
__global__ void bce_backward_kernel(const float* __restrict__ prob, const float* __restrict__ target, int n,
                                    const float* __restrict__ grad_loss, float* __restrict__ dprob) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dprob[i] = grad_loss[0] * bce_dp(prob[i], target[i], 1.f / n);
}
void bce_backward(const float* prob, const float* target, int n, const float* grad_loss, float* dprob,
                  cudaStream_t s) {
    note_launch();
    bce_backward_kernel<<<(n + 255) / 256, 256, 0, s>>>(prob, target, n, grad_loss, dprob);
}

__global__ void d_loss_metrics_kernel(const float* __restrict__ prob, int B, float smoothing,
                                      float* __restrict__ metrics, float* __restrict__ dlogit) {
    __shared__ float sm[32 * 6];
    float acc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};  // loss_r, loss_f, acc_r, acc_f, mean_r, mean_f
    const float inv = 1.f / B;
    for (int i = threadIdx.x; i < B; i += blockDim.x) {
        const float pr = prob[i], pf = prob[B + i];
        acc[0] += bce_term(pr, smoothing);
        acc[1] += bce_term(pf, 0.f);
        acc[2] += pr > 0.5f ? 1.f : 0.f;
        acc[3] += pf < 0.5f ? 1.f : 0.f;
        acc[4] += pr;
        acc[5] += pf;
        dlogit[i] = bce_dp(pr, smoothing, inv) * pr * (1.f - pr);
        dlogit[B + i] = bce_dp(pf, 0.f, inv) * pf * (1.f - pf);
    }
    block_sum<6>(acc, sm);
    if (threadIdx.x == 0) {
        const float lr = acc[0] * inv, lf = acc[1] * inv;
        metrics[0] = lr + lf;
        metrics[1] = lr;
        metrics[2] = lf;
        metrics[3] = acc[2] * inv;
        metrics[4] = acc[3] * inv;
        metrics[5] = acc[4] * inv;
        metrics[6] = acc[5] * inv;
    }
}
void d_loss_metrics(const float* prob, int B, float smoothing, float* metrics, float* dlogit, cudaStream_t s) {
    note_launch();
    d_loss_metrics_kernel<<<1, 1024, 0, s>>>(prob, B, smoothing, metrics, dlogit);
}
__global__ void g_loss_metrics_kernel(const float* __restrict__ prob, int B, float* __restrict__ metrics,
                                      float* __restrict__ dlogit) {
    __shared__ float sm[32 * 2];
    float acc[2] = {0.f, 0.f};
    const float inv = 1.f / B;
    for (int i = threadIdx.x; i < B; i += blockDim.x) {
        const float p = prob[i];
        acc[0] += bce_term(p, 1.f);
        acc[1] += p;
        dlogit[i] = bce_dp(p, 1.f, inv) * p * (1.f - p);
    }
    block_sum<2>(acc, sm);
    if (threadIdx.x == 0) {
        metrics[7] = acc[0] * inv;
        metrics[8] = acc[1] * inv;
    }
}
void g_loss_metrics(const float* prob, int B, float* metrics, float* dlogit, cudaStream_t s) {
    note_launch();
    g_loss_metrics_kernel<<<1, 1024, 0, s>>>(prob, B, metrics, dlogit);
}

// ------------------------------------------------------------------------------------------------
// Adam over a flat buffer (torch.optim.Adam single-tensor formulas, vanilla…:110-120)
// ------------------------------------------------------------------------------------------------
// One Adam update (torch.optim.Adam, wd = 0, amsgrad = False; vanilla…:110-120); returns the new parameter value.
struct AdamCoef {
    float b1, b2, eps, grad_scale, step_size, inv_bc2_sqrt;
};
__device__ __forceinline__ float adam_update(float pi, float graw, float& mref, float& vref, const AdamCoef& k) {
    // every rounding spelled out (no compiler-chosen FMA contraction): the three Adam kernels of this file must agree bit
    // for bit, whichever of them updates a parameter
    const float gi = __fmul_rn(graw, k.grad_scale);
    const float mi = __fmaf_rn(__fsub_rn(gi, mref), 1.f - k.b1, mref);                       // lerp form used by torch
    const float vi = __fmaf_rn(__fmul_rn(1.f - k.b2, gi), gi, __fmul_rn(vref, k.b2));
    mref = mi;
    vref = vi;
    const float denom = __fmaf_rn(sqrtf(vi), k.inv_bc2_sqrt, k.eps);
    return __fsub_rn(pi, __fmul_rn(k.step_size, __fdiv_rn(mi, denom)));
}
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, long n, float b1, float b2, float eps, float step_size,
                            float inv_bc2_sqrt) {
    const AdamCoef k{b1, b2, eps, 1.f, step_size, inv_bc2_sqrt};
    for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<long>(gridDim.x) * blockDim.x) {
        float mi = m[i], vi = v[i];
        p[i] = adam_update(p[i], g[i], mi, vi, k);
        m[i] = mi;
        v[i] = vi;
    }
}
void adam_step(float* p, const float* g, float* m, float* v, long n, float lr, float b1, float b2, float eps, long step,
               cudaStream_t s) {
    const double bc1 = 1.0 - std::pow(static_cast<double>(b1), static_cast<double>(step));
    const double bc2 = 1.0 - std::pow(static_cast<double>(b2), static_cast<double>(step));
    note_launch();
    adam_kernel<<<blocks_for(n, 256), 256, 0, s>>>(p, g, m, v, n, b1, b2, eps, static_cast<float>(lr / bc1),
                                                   static_cast<float>(1.0 / std::sqrt(bc2)));
}

// Device-resident step counters (the fused training step is captured in CUDA graphs, so nothing that changes from step
// to step may be a kernel argument): step_prep advances the Adam step of one network, derives its bias-correction
// scalars with the same double-precision formulas adam_step() uses on the host, and advances the dropout counter.
__global__ void step_prep_kernel(StepCounters* __restrict__ c, int which, float lr, float b1, float b2,
                                 unsigned long long dropout_advance) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    long long step;
    if (which == 0) step = ++c->g_step; else step = ++c->d_step;
    const double bc1 = 1.0 - pow(static_cast<double>(b1), static_cast<double>(step));
    const double bc2 = 1.0 - pow(static_cast<double>(b2), static_cast<double>(step));
    float* o = which == 0 ? c->adam_g : c->adam_d;
    o[0] = static_cast<float>(static_cast<double>(lr) / bc1);
    o[1] = static_cast<float>(1.0 / sqrt(bc2));
    c->dropout_offset += dropout_advance;
}
void step_prep(StepCounters* c, int which, float lr, float b1, float b2, unsigned long long dropout_advance, cudaStream_t s) {
    note_launch();
    step_prep_kernel<<<1, 32, 0, s>>>(c, which, lr, b1, b2, dropout_advance);
}
__global__ void step_set_kernel(StepCounters* __restrict__ c, int field, long long value) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    if (field == 0) c->g_step = value;
    else if (field == 1) c->d_step = value;
    else c->dropout_offset = static_cast<unsigned long long>(value);
}
void step_set(StepCounters* c, int field, long long value, cudaStream_t s) {
    note_launch();
    step_set_kernel<<<1, 32, 0, s>>>(c, field, value);
}
// grad_scale: 1 / world_size when the bucket holds the SUM over data-parallel ranks
__global__ void adam_dev_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                float* __restrict__ v, long n, float b1, float b2, float eps,
                                const float* __restrict__ scal, float grad_scale) {
    const AdamCoef k{b1, b2, eps, grad_scale, scal[0], scal[1]};
    for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<long>(gridDim.x) * blockDim.x) {
        float mi = m[i], vi = v[i];
        p[i] = adam_update(p[i], g[i], mi, vi, k);
        m[i] = mi;
        v[i] = vi;
    }
}

// ---- Adam + weight packs in one launch -------------------------------------------------------------
constexpr int kMaxPlain = 16;
struct AdamPackArgs {
    PackPlan plan;
    float* p;
    const float* g;
    float* m;
    float* v;
    float b1, b2, eps, grad_scale;
    const float* scal;
    int nplain, plain_block0;          // blocks [plain_block0, gridDim.x) update the ranges no pack segment covers
    long plain_off[kMaxPlain];
    long plain_cum[kMaxPlain + 1];     // prefix sums of the range lengths
};
__global__ void __launch_bounds__(256) adam_pack_kernel(const __grid_constant__ AdamPackArgs a) {
    constexpr int kRow = 32 * 16 + 2;
    __shared__ bf16 tile[16 * kRow];
    const AdamCoef k{a.b1, a.b2, a.eps, a.grad_scale, a.scal[0], a.scal[1]};
    const int tid = threadIdx.x;
    auto upd = [&](long i) {  // in place; returns the new value
        float mi = a.m[i], vi = a.v[i];
        const float pn = adam_update(a.p[i], a.g[i], mi, vi, k);
        a.p[i] = pn;
        a.m[i] = mi;
        a.v[i] = vi;
        return pn;
    };
    if (static_cast<int>(blockIdx.x) >= a.plain_block0) {
        const long total = a.plain_cum[a.nplain];
        const long i0 = (static_cast<long>(blockIdx.x) - a.plain_block0) * 4096;
        int r = 0;
        for (long i = i0 + tid; i < total && i < i0 + 4096; i += 256) {
            while (i >= a.plain_cum[r + 1]) ++r;
            upd(a.plain_off[r] + (i - a.plain_cum[r]));
        }
        return;
    }
    int sidx = 0;
    while (sidx + 1 < a.plan.nseg && static_cast<int>(blockIdx.x) >= a.plan.seg[sidx + 1].tile0) ++sidx;
    const PackSeg& sg_ = a.plan.seg[sidx];
    const int t_local = blockIdx.x - sg_.tile0;
    const long base = sg_.src - a.p;  // element offset of the segment's source tensor in the flat buffers
    if (sg_.kind == 0) {
        const int A = sg_.A, B = sg_.B;
        const int bt = B / 32;
        const int a0 = (t_local / bt) * 16, b0 = (t_local % bt) * 32;
        for (int e = tid; e < 16 * 128; e += 256) {
            const int ar = e >> 7, r4 = (e & 127) * 4;
            const long o = base + (static_cast<long>(a0 + ar) * B + b0) * 16 + r4;
            const float4 pv = *reinterpret_cast<const float4*>(a.p + o);
            const float4 gv = __ldg(reinterpret_cast<const float4*>(a.g + o));
            float4 mv = *reinterpret_cast<const float4*>(a.m + o);
            float4 vv = *reinterpret_cast<const float4*>(a.v + o);
            float4 pn;
            pn.x = adam_update(pv.x, gv.x, mv.x, vv.x, k);
            pn.y = adam_update(pv.y, gv.y, mv.y, vv.y, k);
            pn.z = adam_update(pv.z, gv.z, mv.z, vv.z, k);
            pn.w = adam_update(pv.w, gv.w, mv.w, vv.w, k);
            *reinterpret_cast<float4*>(a.p + o) = pn;
            *reinterpret_cast<float4*>(a.m + o) = mv;
            *reinterpret_cast<float4*>(a.v + o) = vv;
            __nv_bfloat162* d = reinterpret_cast<__nv_bfloat162*>(&tile[ar * kRow + r4]);
            d[0] = __floats2bfloat162_rn(pn.x, pn.y);
            d[1] = __floats2bfloat162_rn(pn.z, pn.w);
        }
        __syncthreads();
        if (sg_.ab) {
            for (int e = tid; e < 16 * 16 * 16; e += 256) {
                const int b2 = (e & 15) * 2, t = (e >> 4) & 15, ar = e >> 8;
                __nv_bfloat162 w;
                w.x = tile[ar * kRow + b2 * 16 + t];
                w.y = tile[ar * kRow + (b2 + 1) * 16 + t];
                *reinterpret_cast<__nv_bfloat162*>(&sg_.ab[(static_cast<long>(a0 + ar) * 16 + t) * B + b0 + b2]) = w;
            }
        }
        if (sg_.ba) {
            for (int e = tid; e < 32 * 16 * 8; e += 256) {
                const int a2 = (e & 7) * 2, t = (e >> 3) & 15, b = e >> 7;
                __nv_bfloat162 w;
                w.x = tile[a2 * kRow + b * 16 + t];
                w.y = tile[(a2 + 1) * kRow + b * 16 + t];
                *reinterpret_cast<__nv_bfloat162*>(&sg_.ba[(static_cast<long>(b0 + b) * 16 + t) * A + a0 + a2]) = w;
            }
        }
    } else if (sg_.kind == 1) {  // generator fc: every (feature, latent) pair and every bias appears once in packed order
        const int C0 = sg_.A, latent = sg_.B, Kp = sg_.Kp;
        const long total = static_cast<long>(C0) * 16 * Kp;
        const long bbase = sg_.src2 - a.p;
        for (long i = static_cast<long>(t_local) * 4096 + tid; i < total && i < static_cast<long>(t_local + 1) * 4096; i += 256) {
            const int kk = static_cast<int>(i % Kp);
            const int j = static_cast<int>(i / Kp);
            const int f = (j % C0) * 16 + j / C0;
            sg_.ab[i] = __float2bfloat16(kk < latent ? upd(base + static_cast<long>(f) * latent + kk) : 0.f);
            if (kk == 0) sg_.fdst[j] = upd(bbase + f);
        }
    } else {  // classifier weight: NCHW -> NHWC order, fp32
        const int C = sg_.A, F = C * 16;
        for (int j = t_local * 4096 + tid; j < F && j < (t_local + 1) * 4096; j += 256)
            sg_.fdst[j] = upd(base + (j % C) * 16 + j / C);
    }
}
int adam_pack_step_dev(const PackPlan& plan, float* p, const float* g, float* m, float* v, long n, float b1, float b2,
                       float eps, const float* scal, float grad_scale, cudaStream_t s) {
    AdamPackArgs a;
    a.plan = plan;
    a.p = p; a.g = g; a.m = m; a.v = v;
    a.b1 = b1; a.b2 = b2; a.eps = eps; a.grad_scale = grad_scale; a.scal = scal;
    // covered intervals of the flat buffer, sorted; the complement becomes the plain ranges
    long lo[2 * 8], hi[2 * 8];
    int nc = 0;
    for (int i = 0; i < plan.nseg; ++i) {
        const PackSeg& sg_ = plan.seg[i];
        const long off = sg_.src - p;
        long len = 0;
        if (sg_.kind == 0) len = static_cast<long>(sg_.A) * sg_.B * 16;
        else if (sg_.kind == 1) len = static_cast<long>(sg_.A) * 16 * sg_.B;
        else len = static_cast<long>(sg_.A) * 16;
        if (off < 0 || off + len > n || (off & 3)) return -1;
        lo[nc] = off; hi[nc] = off + len; ++nc;
        if (sg_.kind == 1) {
            const long boff = sg_.src2 - p;
            if (boff < 0 || boff + sg_.A * 16 > n) return -1;
            lo[nc] = boff; hi[nc] = boff + sg_.A * 16; ++nc;
        }
    }
    for (int i = 1; i < nc; ++i)
        for (int j = i; j > 0 && lo[j] < lo[j - 1]; --j) {
            const long tl = lo[j], th = hi[j];
            lo[j] = lo[j - 1]; hi[j] = hi[j - 1];
            lo[j - 1] = tl; hi[j - 1] = th;
        }
    a.nplain = 0;
    a.plain_cum[0] = 0;
    long cur = 0;
    for (int i = 0; i <= nc; ++i) {
        const long end = i < nc ? lo[i] : n;
        if (end < cur) return -1;  // overlapping segments
        if (end > cur) {
            if (a.nplain >= kMaxPlain) return -1;
            a.plain_off[a.nplain] = cur;
            a.plain_cum[a.nplain + 1] = a.plain_cum[a.nplain] + (end - cur);
            ++a.nplain;
        }
        if (i < nc) cur = hi[i];
    }
    a.plain_block0 = plan.total_tiles;
    const int plain_blocks = static_cast<int>((a.plain_cum[a.nplain] + 4095) / 4096);
    note_launch();
    adam_pack_kernel<<<plan.total_tiles + plain_blocks, 256, 0, s>>>(a);
    return 0;
}
void adam_step_dev(float* p, const float* g, float* m, float* v, long n, float b1, float b2, float eps, const float* scal,
                   float grad_scale, cudaStream_t s) {
    note_launch();
    adam_dev_kernel<<<blocks_for(n, 256), 256, 0, s>>>(p, g, m, v, n, b1, b2, eps, scal, grad_scale);
}

// ------------------------------------------------------------------------------------------------
// direct convolutions (validation mode)
// ------------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ float epi_apply(float v, int n, long img, long oidx, const Epi& e) {
    if (e.bias) v += e.bias[n];
    if (e.scale) v = fmaf(v, e.scale[n], e.shift[n]);
    if (e.act == 1)
        v = fmaxf(v, 0.f);
    else if (e.act == 2)
        v = v > 0.f ? v : v * e.slope;
    if (e.mask) v *= e.mask[img * e.ldmask + n];
    if (e.gate) v *= to_f(static_cast<const T*>(e.gate)[oidx]) > 0.f ? 1.f : e.slope;
    return v;
}

template <typename T>
__global__ void conv_s2_direct_kernel(const T* __restrict__ in, const float* __restrict__ w, long so, long si, Epi e,
                                      T* __restrict__ out, int B, int inH, int inW, int Cin, int Cout) {
    const int OH = inH / 2, OW = inW / 2;
    const long total = static_cast<long>(B) * OH * OW * Cout;
    for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long>(gridDim.x) * blockDim.x) {
        const int co = static_cast<int>(i % Cout);
        const long pix = i / Cout;
        const int ox = static_cast<int>(pix % OW);
        const int oy = static_cast<int>((pix / OW) % OH);
        const long n = pix / (static_cast<long>(OW) * OH);
        float acc = 0.f;
        for (int ky = 0; ky < 4; ++ky) {
            const int iy = 2 * oy - 1 + ky;
            if (iy < 0 || iy >= inH) continue;
            for (int kx = 0; kx < 4; ++kx) {
                const int ix = 2 * ox - 1 + kx;
                if (ix < 0 || ix >= inW) continue;
                const T* p = in + ((n * inH + iy) * inW + ix) * Cin;
                const float* ww = w + co * so + ky * 4 + kx;
                for (int ci = 0; ci < Cin; ++ci) acc = fmaf(to_f(p[ci]), ww[ci * si], acc);
            }
        }
        out[i] = from_f<T>(epi_apply<T>(acc, co, n, i, e));
    }
}
template <typename T>
void conv_s2_direct(const T* in, const float* w, long so, long si, const Epi& e, T* out, int B, int inH, int inW,
                    int Cin, int Cout, cudaStream_t s) {
    const long total = static_cast<long>(B) * (inH / 2) * (inW / 2) * Cout;
    note_launch();
    conv_s2_direct_kernel<T><<<blocks_for(total, 256, 148 * 64), 256, 0, s>>>(in, w, so, si, e, out, B, inH, inW, Cin,
                                                                             Cout);
}
template void conv_s2_direct<float>(const float*, const float*, long, long, const Epi&, float*, int, int, int, int, int,
                                    cudaStream_t);
template void conv_s2_direct<bf16>(const bf16*, const float*, long, long, const Epi&, bf16*, int, int, int, int, int,
                                   cudaStream_t);

template <typename T>
__global__ void convT_direct_kernel(const T* __restrict__ in, const float* __restrict__ w, long so, long si, Epi e,
                                    T* __restrict__ out, int B, int inH, int inW, int Cin, int Cout) {
    const int OH = inH * 2, OW = inW * 2;
    const long total = static_cast<long>(B) * OH * OW * Cout;
    for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long>(gridDim.x) * blockDim.x) {
        const int co = static_cast<int>(i % Cout);
        const long pix = i / Cout;
        const int ox = static_cast<int>(pix % OW);
        const int oy = static_cast<int>((pix / OW) % OH);
        const long n = pix / (static_cast<long>(OW) * OH);
        float acc = 0.f;
        for (int ty = 0; ty < 2; ++ty) {
            const int ky = ((oy + 1) & 1) + 2 * ty;
            const int t = oy + 1 - ky;
            if (t < 0 || (t >> 1) >= inH) continue;
            const int iy = t >> 1;
            for (int tx = 0; tx < 2; ++tx) {
                const int kx = ((ox + 1) & 1) + 2 * tx;
                const int u = ox + 1 - kx;
                if (u < 0 || (u >> 1) >= inW) continue;
                const int ix = u >> 1;
                const T* p = in + ((n * inH + iy) * inW + ix) * Cin;
                const float* ww = w + co * so + ky * 4 + kx;
                for (int ci = 0; ci < Cin; ++ci) acc = fmaf(to_f(p[ci]), ww[ci * si], acc);
            }
        }
        out[i] = from_f<T>(epi_apply<T>(acc, co, n, i, e));
    }
}
template <typename T>
void convT_direct(const T* in, const float* w, long so, long si, const Epi& e, T* out, int B, int inH, int inW, int Cin,
                  int Cout, cudaStream_t s) {
    const long total = static_cast<long>(B) * inH * 2 * inW * 2 * Cout;
    note_launch();
    convT_direct_kernel<T><<<blocks_for(total, 256, 148 * 64), 256, 0, s>>>(in, w, so, si, e, out, B, inH, inW, Cin,
                                                                           Cout);
}
template void convT_direct<float>(const float*, const float*, long, long, const Epi&, float*, int, int, int, int, int,
                                  cudaStream_t);
template void convT_direct<bf16>(const bf16*, const float*, long, long, const Epi&, bf16*, int, int, int, int, int,
                                 cudaStream_t);

// one warp per (m, n, tap): lanes stride over pixels, shuffle-reduce (fixed order => deterministic)
template <typename T>
__global__ void wgrad_direct_kernel(const T* __restrict__ coarse, const T* __restrict__ fine, float* __restrict__ dW,
                                    int B, int cH, int cW, int Mc, int Nf) {
    const long warp = (blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const long total = static_cast<long>(Mc) * Nf * 16;
    if (warp >= total) return;
    const int tap = static_cast<int>(warp & 15);
    const int n = static_cast<int>((warp >> 4) % Nf);
    const int m = static_cast<int>((warp >> 4) / Nf);
    const int ky = tap >> 2, kx = tap & 3;
    const int fH = 2 * cH, fW = 2 * cW;
    const long pixels = static_cast<long>(B) * cH * cW;
    float acc = 0.f;
    for (long pix = lane; pix < pixels; pix += 32) {
        const int x = static_cast<int>(pix % cW);
        const int y = static_cast<int>((pix / cW) % cH);
        const long b = pix / (static_cast<long>(cW) * cH);
        const int fy = 2 * y - 1 + ky, fx = 2 * x - 1 + kx;
        if (fy < 0 || fy >= fH || fx < 0 || fx >= fW) continue;
        acc = fmaf(to_f(coarse[pix * Mc + m]), to_f(fine[((b * fH + fy) * fW + fx) * Nf + n]), acc);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) dW[warp] = acc;
}
template <typename T>
void wgrad_direct(const T* coarse, const T* fine, float* dW, int B, int cH, int cW, int Mc, int Nf, cudaStream_t s) {
    const long total = static_cast<long>(Mc) * Nf * 16;
    note_launch();
    wgrad_direct_kernel<T><<<static_cast<unsigned>((total * 32 + 255) / 256), 256, 0, s>>>(coarse, fine, dW, B, cH, cW,
                                                                                           Mc, Nf);
}
template void wgrad_direct<float>(const float*, const float*, float*, int, int, int, int, int, cudaStream_t);
template void wgrad_direct<bf16>(const bf16*, const bf16*, float*, int, int, int, int, int, cudaStream_t);

template <typename T>
__global__ void fc_direct_kernel(const T* __restrict__ zp, int Kp, const float* __restrict__ W,
                                 const float* __restrict__ bias, T* __restrict__ y, int B, int C0, int latent) {
    const int F = C0 * 16;
    const long total = static_cast<long>(B) * F;
    for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long>(gridDim.x) * blockDim.x) {
        const int j = static_cast<int>(i % F);
        const long b = i / F;
        const int f = (j % C0) * 16 + j / C0;
        float acc = 0.f;
        for (int k = 0; k < latent; ++k) acc = fmaf(to_f(zp[b * Kp + k]), W[static_cast<long>(f) * latent + k], acc);
        y[i] = from_f<T>(acc + bias[f]);
    }
}
template <typename T>
void fc_direct(const T* zp, int Kp, const float* W, const float* bias, T* y, int B, int C0, int latent,
               cudaStream_t s) {
    note_launch();
    fc_direct_kernel<T><<<blocks_for(static_cast<long>(B) * C0 * 16, 256, 148 * 64), 256, 0, s>>>(zp, Kp, W, bias, y, B,
                                                                                                 C0, latent);
}
template void fc_direct<float>(const float*, int, const float*, const float*, float*, int, int, int, cudaStream_t);
template void fc_direct<bf16>(const bf16*, int, const float*, const float*, bf16*, int, int, int, cudaStream_t);

template <typename T>
__global__ void fc_wgrad_direct_kernel(const T* __restrict__ dy, const T* __restrict__ zp, int Kp,
                                       float* __restrict__ dW, int B, int C0, int latent) {
    const int F = C0 * 16;
    const long total = static_cast<long>(F) * latent;
    for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long>(gridDim.x) * blockDim.x) {
        const int j = static_cast<int>(i % F);  // consecutive threads -> consecutive columns of dy (coalesced)
        const int k = static_cast<int>(i / F);
        const int f = (j % C0) * 16 + j / C0;
        float acc = 0.f;
        for (int b = 0; b < B; ++b) acc = fmaf(to_f(dy[static_cast<long>(b) * F + j]), to_f(zp[static_cast<long>(b) * Kp + k]), acc);
        dW[static_cast<long>(f) * latent + k] = acc;
    }
}
template <typename T>
void fc_wgrad_direct(const T* dy, const T* zp, int Kp, float* dW, int B, int C0, int latent, cudaStream_t s) {
    note_launch();
    fc_wgrad_direct_kernel<T><<<blocks_for(static_cast<long>(C0) * 16 * latent, 256, 148 * 64), 256, 0, s>>>(
        dy, zp, Kp, dW, B, C0, latent);
}
template void fc_wgrad_direct<float>(const float*, const float*, int, float*, int, int, int, cudaStream_t);
template void fc_wgrad_direct<bf16>(const bf16*, const bf16*, int, float*, int, int, int, cudaStream_t);

template <typename T>
__global__ void fc_dz_direct_kernel(const T* __restrict__ dy, const float* __restrict__ W, float* __restrict__ dz,
                                    int B, int C0, int latent) {
    const int F = C0 * 16;
    const long total = static_cast<long>(B) * latent;
    for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long>(gridDim.x) * blockDim.x) {
        const int k = static_cast<int>(i % latent);
        const long b = i / latent;
        float acc = 0.f;
        for (int j = 0; j < F; ++j) {
            const int f = (j % C0) * 16 + j / C0;
            acc = fmaf(to_f(dy[b * F + j]), W[static_cast<long>(f) * latent + k], acc);
        }
        dz[i] = acc;
    }
}
template <typename T>
void fc_dz_direct(const T* dy, const float* W, float* dz, int B, int C0, int latent, cudaStream_t s) {
    note_launch();
    fc_dz_direct_kernel<T><<<blocks_for(static_cast<long>(B) * latent, 256), 256, 0, s>>>(dy, W, dz, B, C0, latent);
}
template void fc_dz_direct<float>(const float*, const float*, float*, int, int, int, cudaStream_t);
template void fc_dz_direct<bf16>(const bf16*, const float*, float*, int, int, int, cudaStream_t);

}  // namespace sg
