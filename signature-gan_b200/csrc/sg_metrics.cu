// sg_metrics.cu — per-image ink statistics of a batch of generated / real images in one pass over HBM
// (SURVEY.md §8f-4). Reference: src/utils/metrics.py:118-174 `calculate_stroke_density` / `calculate_foreground_ratio`
// (called by evaluate_vanilla_gan_signatures.py:306-332):
//     if images.min() < 0: images = (images + 1) / 2          # data-dependent rescale, decided on the WHOLE batch
//     stroke = (images < threshold).float();  per-image mean over the pixels
// The rescale decision needs the global minimum, so one pass produces, per image, the minimum and BOTH counts
// (#(x < t) and #((x + 1) / 2 < t), the latter with the same two float32 roundings as torch); the caller picks the column
// after reducing the B minima. Counts are integers: bit-exact. One warp-strided CTA per image at a time, float4 loads,
// grid = a multiple of the SM count; 4 bytes read per pixel, 12 bytes written per image.
#include <cuda_runtime.h>

#include <cfloat>

#include "sg_kernels.cuh"

namespace sg {
namespace {

__global__ void __launch_bounds__(256) ink_stats_kernel(const float* __restrict__ images, int n_images, int pixels,
                                                        float threshold, int* __restrict__ count_raw,
                                                        int* __restrict__ count_rescaled, float* __restrict__ minimum) {
    __shared__ int s_raw[8], s_res[8];
    __shared__ float s_min[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int img = blockIdx.x; img < n_images; img += gridDim.x) {
        const float4* p = reinterpret_cast<const float4*>(images + static_cast<long long>(img) * pixels);
        int raw = 0, res = 0;
        float mn = FLT_MAX;
        for (int i = threadIdx.x; i < pixels / 4; i += blockDim.x) {
            const float4 v = __ldg(p + i);
            const float x[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                raw += x[j] < threshold;
                res += __fdiv_rn(__fadd_rn(x[j], 1.0f), 2.0f) < threshold;
                mn = fminf(mn, x[j]);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            raw += __shfl_xor_sync(0xffffffffu, raw, o);
            res += __shfl_xor_sync(0xffffffffu, res, o);
            mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        }
        if (lane == 0) {
            s_raw[warp] = raw;
            s_res[warp] = res;
            s_min[warp] = mn;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int a = 0, b = 0;
            float m = FLT_MAX;
            for (int w = 0; w < 8; ++w) {
                a += s_raw[w];
                b += s_res[w];
                m = fminf(m, s_min[w]);
            }
            count_raw[img] = a;
            count_rescaled[img] = b;
            minimum[img] = m;
        }
        __syncthreads();
    }
}

}  // namespace

void ink_stats(const float* images, int n_images, int pixels, float threshold, int* count_raw, int* count_rescaled,
               float* minimum, cudaStream_t s) {
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (sms <= 0) sms = 148;
    }
    int grid = sms * 8;
    if (grid > n_images) grid = n_images;
    note_launch();
    ink_stats_kernel<<<grid, 256, 0, s>>>(images, n_images, pixels, threshold, count_raw, count_rescaled, minimum);
}

}  // namespace sg
