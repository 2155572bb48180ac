// sg_mma.cuh — warp-level mma.sync / ldmatrix helpers for the thin (HBM-bound) layers, where a 128-row UMMA tile
// would be mostly padding and the accumulators are wanted in registers for a fused epilogue.
#pragma once
#include <cuda_bf16.h>
#include <cstdint>

namespace sg {

__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t (&r)[4]) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t (&r)[4]) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(addr));
}
// D (16x8, fp32) += A (16x16, bf16, row) * B (16x8, bf16, col). Fragment layout (gid = lane >> 2, t4 = lane & 3):
//   a0: (row gid, k 2*t4..+1)  a1: (row gid+8, same k)  a2: (row gid, k+8)  a3: (row gid+8, k+8)
//   b0: (k 2*t4..+1, n gid)    b1: (k+8, n gid)          d0,d1: (row gid, n 2*t4..+1)  d2,d3: (row gid+8, ...)
__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
        "{%0, %1, %2, %3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack2_bf16(float lo, float hi) {
    __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&h);
}

}  // namespace sg
