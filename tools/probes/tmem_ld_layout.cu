// Probe: which (lane, column) of TMEM does each register of tcgen05.ld.16x256b.x4 return?
// Writes lane*100 + column with 32x32b stores (4 warps = 128 lanes x 32 columns), reads back with 16x256b.x4.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void probe(int* out) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)), "r"(32) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = slot;
    const uint32_t wbase = base + ((uint32_t)(warp * 32) << 16);
    for (int c = 0; c < 32; c += 4) {
        uint32_t v0 = (warp * 32 + lane) * 100 + c, v1 = v0 + 1, v2 = v0 + 2, v3 = v0 + 3;
        asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(wbase + c), "r"(v0), "r"(v1), "r"(v2), "r"(v3) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int half = 0; half < 2; ++half) {
        uint32_t r[16];
        asm volatile(
            "tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
              "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
            : "r"(wbase + ((uint32_t)(half * 16) << 16))
            : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int i = 0; i < 16; ++i) out[((warp * 2 + half) * 32 + lane) * 16 + i] = (int)r[i];
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(32) : "memory");
}
int main() {
    int* d;
    cudaMalloc(&d, 4 * 2 * 32 * 16 * 4);
    probe<<<1, 128>>>(d);
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    static int h[4 * 2 * 32 * 16];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int w = 0; w < 4; ++w)
        for (int half = 0; half < 2; ++half)
            for (int t = 0; t < 32; ++t)
                for (int i = 0; i < 16; ++i) {
                    const int n = i >> 2, j = i & 3;
                    const int row = w * 32 + half * 16 + (t >> 2) + (j >> 1) * 8, col = 8 * n + 2 * (t & 3) + (j & 1);
                    const int got = h[((w * 2 + half) * 32 + t) * 16 + i];
                    if (got != row * 100 + col) {
                        if (bad < 20) printf("w%d half%d t%d r%d: got lane %d col %d, expected lane %d col %d\n", w, half, t, i, got / 100, got % 100, row, col);
                        ++bad;
                    }
                }
    printf("mismatches vs mma-C-fragment hypothesis: %d\n", bad);
    return 0;
}
