// sg_spectral.cu — spectral normalisation of the Discriminator's weights, the `use_spectral_norm=True` variant of
// the reference (src/discriminator_vanilla_gan.py:61-62 Conv2d, :201-202 Linear wrap their layer in
// torch.nn.utils.spectral_norm; ablation_vanilla_gan_signatures.py:367-371 builds that variant).
//
// What torch's SpectralNorm.compute_weight does on every forward of the wrapped layer, with W = weight_orig viewed
// as (rows = Cout, cols = Cin*kh*kw), u (rows) and v (cols) the persistent buffers weight_u / weight_v:
//     training mode, n_power_iterations (= 1) times, in place and without gradient:
//         v <- W^T u / max(||W^T u||, eps);   u <- W v / max(||W v||, eps)
//     sigma = u . (W v);   weight = W / sigma            (u, v are constants for the gradient)
// so   dL/dW = (G - <G, weight> u v^T) / sigma           with G = dL/dweight.
// It is weight preprocessing only: the convolutions themselves run on the effective weights through the same
// tcgen05 kernels. All of it is HBM-bound streaming over at most 8 MB per layer; sums are fp32 and deterministic
// (fixed reduction order, no atomics).
#include <cuda_runtime.h>

#include "sg_kernels.cuh"

namespace sg {
namespace {

constexpr int kSnThreads = 256;
constexpr int kSnMaxPartials = 512;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Sum over the block, same value returned to every thread; the order of additions is fixed by the launch shape.
__device__ __forceinline__ float block_sum(float v, float* red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();  // red may still be read by the previous call
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float t = (lane < nwarp) ? red[lane] : 0.f;
    return warp_sum(t);
}

// t[c] = sum_r W[r][c] * u[r]; one thread per column, rows walked eight at a time (independent loads in flight).
__global__ void __launch_bounds__(kSnThreads) sn_wt_u_kernel(const float* __restrict__ W, const float* __restrict__ u,
                                                             float* __restrict__ t, int rows, int cols) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cols) return;
    float acc = 0.f;
    int r = 0;
    for (; r + 8 <= rows; r += 8) {
        float w[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) w[k] = __ldg(W + static_cast<size_t>(r + k) * cols + c);
#pragma unroll
        for (int k = 0; k < 8; ++k) acc = fmaf(w[k], __ldg(u + r + k), acc);
    }
    for (; r < rows; ++r) acc = fmaf(__ldg(W + static_cast<size_t>(r) * cols + c), __ldg(u + r), acc);
    t[c] = acc;
}

// x <- x / max(||x||_2, eps)  (torch.nn.functional.normalize), one block.
__global__ void __launch_bounds__(1024) sn_normalize_kernel(float* __restrict__ x, int n, float eps) {
    __shared__ float red[32];
    float ss = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) ss = fmaf(x[i], x[i], ss);
    const float d = fmaxf(sqrtf(block_sum(ss, red)), eps);
    for (int i = threadIdx.x; i < n; i += blockDim.x) x[i] = x[i] / d;
}

// s[r] = sum_c W[r][c] * v[c]; one block per row.
__global__ void __launch_bounds__(kSnThreads) sn_w_v_kernel(const float* __restrict__ W, const float* __restrict__ v,
                                                            float* __restrict__ s, int cols) {
    __shared__ float red[32];
    const float* row = W + static_cast<size_t>(blockIdx.x) * cols;
    float acc = 0.f;
    if ((cols & 3) == 0) {
        const float4* r4 = reinterpret_cast<const float4*>(row);
        const float4* v4 = reinterpret_cast<const float4*>(v);
        for (int i = threadIdx.x; i < cols / 4; i += blockDim.x) {
            const float4 a = __ldg(r4 + i), b = __ldg(v4 + i);
            acc = fmaf(a.x, b.x, acc);
            acc = fmaf(a.y, b.y, acc);
            acc = fmaf(a.z, b.z, acc);
            acc = fmaf(a.w, b.w, acc);
        }
    } else {
        for (int i = threadIdx.x; i < cols; i += blockDim.x) acc = fmaf(__ldg(row + i), __ldg(v + i), acc);
    }
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) s[blockIdx.x] = acc;
}

// normalise != 0: u <- s / max(||s||, eps). Then sigma = u . s. One block.
__global__ void __launch_bounds__(1024) sn_finish_u_kernel(const float* __restrict__ s, float* __restrict__ u, int rows,
                                                           float eps, int normalise, float* __restrict__ sigma) {
    __shared__ float red[32];
    if (normalise) {
        float ss = 0.f;
        for (int i = threadIdx.x; i < rows; i += blockDim.x) ss = fmaf(s[i], s[i], ss);
        const float d = fmaxf(sqrtf(block_sum(ss, red)), eps);
        for (int i = threadIdx.x; i < rows; i += blockDim.x) u[i] = s[i] / d;
    }
    float dot = 0.f;  // every thread re-reads only the u[i] it wrote itself
    for (int i = threadIdx.x; i < rows; i += blockDim.x) dot = fmaf(u[i], s[i], dot);
    dot = block_sum(dot, red);
    if (threadIdx.x == 0) sigma[0] = dot;
}

__global__ void __launch_bounds__(kSnThreads) sn_scale_kernel(const float* __restrict__ W, const float* __restrict__ sigma,
                                                              float* __restrict__ out, long long n) {
    const float sg = __ldg(sigma);
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
        out[i] = __ldg(W + i) / sg;
}

// partial[b] = sum over block b's grid-stride slice of g * w
__global__ void __launch_bounds__(kSnThreads) sn_inner_kernel(const float* __restrict__ g, const float* __restrict__ w,
                                                              long long n, float* __restrict__ partial) {
    __shared__ float red[32];
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    float acc = 0.f;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
        acc = fmaf(__ldg(g + i), __ldg(w + i), acc);
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) partial[blockIdx.x] = acc;
}

// g[r][c] <- (g[r][c] - inner * u[r] * v[c]) / sigma, inner = sum of the partials (same order in every block).
__global__ void __launch_bounds__(kSnThreads) sn_grad_kernel(float* __restrict__ g, const float* __restrict__ u,
                                                             const float* __restrict__ v, const float* __restrict__ sigma,
                                                             const float* __restrict__ partial, int npartial, int rows,
                                                             int cols) {
    __shared__ float red[32];
    float acc = 0.f;
    for (int i = threadIdx.x; i < npartial; i += blockDim.x) acc += partial[i];
    const float inner = block_sum(acc, red);
    const float sg = __ldg(sigma);
    const long long n = static_cast<long long>(rows) * cols;
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int r = static_cast<int>(i / cols), c = static_cast<int>(i - static_cast<long long>(r) * cols);
        g[i] = (g[i] - inner * __ldg(u + r) * __ldg(v + c)) / sg;
    }
}

int stream_blocks(long long n) {
    long long b = (n + kSnThreads * 4 - 1) / (kSnThreads * 4);
    if (b < 1) b = 1;
    if (b > kSnMaxPartials) b = kSnMaxPartials;  // 148 SMs x 3-4 resident blocks
    return static_cast<int>(b);
}

}  // namespace

void spectral_norm_weight(const float* w_orig, float* u, float* v, int rows, int cols, int power_iterations, float eps,
                          float* w_out, float* sigma, float* scratch, cudaStream_t s) {
    for (int it = 0; it < power_iterations; ++it) {
        sn_wt_u_kernel<<<(cols + kSnThreads - 1) / kSnThreads, kSnThreads, 0, s>>>(w_orig, u, v, rows, cols);
        sn_normalize_kernel<<<1, 1024, 0, s>>>(v, cols, eps);
        sn_w_v_kernel<<<rows, kSnThreads, 0, s>>>(w_orig, v, scratch, cols);
        sn_finish_u_kernel<<<1, 1024, 0, s>>>(scratch, u, rows, eps, 1, sigma);
        note_launch(4);
    }
    if (power_iterations <= 0) {
        sn_w_v_kernel<<<rows, kSnThreads, 0, s>>>(w_orig, v, scratch, cols);
        sn_finish_u_kernel<<<1, 1024, 0, s>>>(scratch, u, rows, eps, 0, sigma);
        note_launch(2);
    }
    const long long n = static_cast<long long>(rows) * cols;
    sn_scale_kernel<<<stream_blocks(n), kSnThreads, 0, s>>>(w_orig, sigma, w_out, n);
    note_launch();
}

void spectral_norm_backward(const float* w_eff, const float* u, const float* v, const float* sigma, int rows, int cols,
                            float* grad, float* scratch, cudaStream_t s) {
    const long long n = static_cast<long long>(rows) * cols;
    const int blocks = stream_blocks(n);
    sn_inner_kernel<<<blocks, kSnThreads, 0, s>>>(grad, w_eff, n, scratch);
    sn_grad_kernel<<<blocks, kSnThreads, 0, s>>>(grad, u, v, sigma, scratch, blocks, rows, cols);
    note_launch(2);
}

}  // namespace sg
