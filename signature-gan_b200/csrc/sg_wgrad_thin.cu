// sg_wgrad_thin.cu — weight gradient of the thin 4x4/s2 (transposed) convolutions (fine tensor with 32 channels,
// coarse tensor with 32 or 64: the Generator's last two upsample blocks, gen…:131-149).
//
//   dW[m][n][ky][kx] = sum over coarse pixels (img, y, x) of coarse[img,y,x][m] * fine[img, 2y-1+ky, 2x-1+kx][n]
//
// Why not the tcgen05 kernel (wgrad_umma_kernel): with 32-wide operands a 128 x 64 UMMA tile is 7/8 padding, and —
// what actually bounds it — every filter tap needs its own TMA-staged copy of the fine tensor (the contraction runs
// over pixels, and a swizzled MN-major operand cannot be shifted by one pixel), so each 64-byte pixel row crosses the
// TMA unit 16 times (measured: 2.5 ms at B = 4096, TMA row rate bound). Here every pixel row is staged ONCE:
//   * one CTA per SM walks (image, band of 128 coarse pixels) units; an elected thread streams the band of the coarse
//     tensor and the 2R+2 fine rows it touches through a 3-slot shared-memory ring with two TMA tensor loads per unit
//     (hardware swizzle; rows above/below the image are zero-filled by TMA = the convolution's padding);
//   * the 8 warps own 2 filter taps each (same ky, adjacent kx) and run mma.sync.m16n8k16 (bf16 in, fp32 accumulate)
//     over the band; both operands are pixel-major in memory, i.e. transposed with respect to the MMA fragment
//     layouts, which is exactly what ldmatrix.trans undoes — and since each lane supplies its own row address,
//     the stride-2, per-tap shifted gather of fine pixels costs nothing (left/right padding = a zero row);
//   * accumulators live in registers for the whole launch; each CTA writes one partial [16][M][N] block and the
//     existing deterministic reduction folds the <= 148 partials into the flat gradient bucket.
// The kernel is HBM-bound by design: 64 KB + 256 KB per image for the 32 -> 32 block at 64x64.
#include "sg_conv_umma.cuh"
#include "sg_kernels.cuh"
#include "sg_mma.cuh"
#include "sg_umma.cuh"

#include <cstdio>
#include <cstring>

namespace sg {

void wgrad_reduce(const float* partial, float* dW, int S, int M, int N, int accumulate, cudaStream_t stream,
                  const float* bias_partial = nullptr, float* dbias = nullptr, int SB = 0);

namespace {

constexpr int kBandPix = 128;  // coarse pixels per unit
constexpr int kWtSlots = 3;
constexpr int kWtThreads = 256;

struct WgradThinArgs {
    CUtensorMap cmap;  // coarse as [pixels][MC], box {MC, 128}
    CUtensorMap fmap;  // fine as [N][2cH][2cW][32], box {32, 2cW, 2R+2, 1}
    int cH, cW, nimg;
    float* partial;    // [gridDim.x][16][MC][32]
};

template <int MC>
struct WtCfg {
    static constexpr int kCoarseBytes = kBandPix * MC * 2;
    // fine rows per unit: 2R+2 rows of 2cW pixels with R = 128/cW  ->  (512 + 4cW) pixels of 64 bytes, cW <= 64
    static constexpr int kFineBytesMax = (512 + 4 * 64) * 64;
    static constexpr int kSlotBytes = kCoarseBytes + kFineBytesMax;
    static constexpr int kSmem = kWtSlots * kSlotBytes + 1024 /*zero rows*/ + 2 * kWtSlots * 8 + 1024 /*align*/;
};

template <int MC>
__device__ __forceinline__ void wt_issue(const WgradThinArgs& args, uint8_t* smem, uint64_t* full, uint64_t* empty,
                                         uint32_t q, int bands, uint32_t fine_bytes) {
    using Cfg = WtCfg<MC>;
    const int u = blockIdx.x + static_cast<int>(q) * gridDim.x;
    const int n = u / bands, band = u - n * bands;
    const int R = kBandPix / args.cW;
    const int slot = q % kWtSlots;
    mbar_wait(&empty[slot], ((q / kWtSlots) & 1) ^ 1);
    mbar_arrive_expect_tx(&full[slot], Cfg::kCoarseBytes + fine_bytes);
    uint8_t* dst = smem + slot * Cfg::kSlotBytes;
    tma_load_2d(dst, &args.cmap, &full[slot], 0, (n * args.cH + band * R) * args.cW);
    tma_load_4d(dst + Cfg::kCoarseBytes, &args.fmap, &full[slot], 0, 0, 2 * band * R - 1, n);
}

template <int MC>
__global__ void __launch_bounds__(kWtThreads, 1) wgrad_thin_kernel(const __grid_constant__ WgradThinArgs args) {
    using Cfg = WtCfg<MC>;
    constexpr int NF = 32;
    constexpr int MB = MC / 16;             // m-blocks of 16 coarse channels
    constexpr int CROW = MC * 2;            // bytes per coarse pixel row (64: SW64, 128: SW128)
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* zero = smem + kWtSlots * Cfg::kSlotBytes;
    uint64_t* full = reinterpret_cast<uint64_t*>(zero + 1024);
    uint64_t* empty = full + kWtSlots;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < 256; i += kWtThreads) reinterpret_cast<uint32_t*>(zero)[i] = 0;
    if (tid == 0) {
        tma_prefetch_desc(&args.cmap);
        tma_prefetch_desc(&args.fmap);
        for (int s = 0; s < kWtSlots; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], kWtThreads / 32);
        }
        mbar_fence_init();
    }
    __syncthreads();

    const int cW = args.cW, fW = 2 * cW;
    const int R = kBandPix / cW;                        // coarse rows per band
    const int bands = args.cH / R, units = args.nimg * bands;
    const uint32_t fine_bytes = static_cast<uint32_t>((2 * R + 2) * fW) * 64;
    const uint32_t total_q = static_cast<uint32_t>((units - static_cast<int>(blockIdx.x) + gridDim.x - 1) / gridDim.x);
    if (tid == 0)
        for (uint32_t q = 0; q < kWtSlots - 1 && q < total_q; ++q) wt_issue<MC>(args, smem, full, empty, q, bands, fine_bytes);

    // this warp's two taps: ky = warp >> 1, kx = 2 * (warp & 1) + {0, 1}
    const int ky = warp >> 1, kx0 = 2 * (warp & 1);
    // ldmatrix lane roles: matrix mi = lane >> 3, row ri = lane & 7
    const int mi = lane >> 3, ri = lane & 7;
    // A (coarse): matrices (m-chunk cm, k-half kh) = (mi & 1, mi >> 1): pixel kh*8 + ri of the k16 step, 16-byte chunk cm
    const int a_pix = (mi >> 1) * 8 + ri;
    uint32_t a_lane[MB];  // byte offset inside the coarse tile for step 0 (swizzle is step-invariant)
#pragma unroll
    for (int mb = 0; mb < MB; ++mb) {
        const int chunk = mb * 2 + (mi & 1);
        const int sw = MC == 32 ? ((a_pix >> 1) & 3) : (a_pix & 7);
        a_lane[mb] = a_pix * CROW + ((chunk ^ sw) << 4);
    }
    // B (fine): matrices (k-half kh, n-chunk cn) = (mi & 1, mi >> 1) (+2 chunks for the second ldmatrix)
    const int b_k = (mi & 1) * 8 + ri;   // coarse pixel within the k16 step
    float acc[2][MB][4][4];
#pragma unroll
    for (int t = 0; t < 2; ++t)
#pragma unroll
        for (int mb = 0; mb < MB; ++mb)
#pragma unroll
            for (int nb = 0; nb < 4; ++nb)
#pragma unroll
                for (int e = 0; e < 4; ++e) acc[t][mb][nb][e] = 0.f;

    const uint32_t zero_u32 = smem_u32(zero);
    for (uint32_t it = 0; it < total_q; ++it) {
        const int slot = it % kWtSlots;
        if (tid == 0 && it + kWtSlots - 1 < total_q)
            wt_issue<MC>(args, smem, full, empty, it + kWtSlots - 1, bands, fine_bytes);
        __syncwarp();
        mbar_wait(&full[slot], (it / kWtSlots) & 1);
        const uint32_t cbase = smem_u32(smem + slot * Cfg::kSlotBytes);
        const uint32_t fbase = cbase + Cfg::kCoarseBytes;
#pragma unroll 1
        for (int ks = 0; ks < kBandPix / 16; ++ks) {
            const int p0 = ks * 16;            // first coarse pixel of the step inside the band
            const int yl = p0 / cW, x0 = p0 - yl * cW;
            uint32_t a[MB][4];
#pragma unroll
            for (int mb = 0; mb < MB; ++mb) ldsm_x4_t(cbase + p0 * CROW + a_lane[mb], a[mb]);
            const int frow = 2 * yl + ky;      // fine row inside the staged rows (row 0 = 2*y0 - 1)
#pragma unroll
            for (int t = 0; t < 2; ++t) {
                const int fx = 2 * (x0 + b_k) - 1 + kx0 + t;
                const int prow = frow * fW + fx;   // pixel row inside the fine tile
                const bool oob = fx < 0 || fx >= fW;
                uint32_t b[2][4];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int chunk = h * 2 + (mi >> 1);
                    const uint32_t addr = fbase + prow * 64 + ((chunk ^ ((prow >> 1) & 3)) << 4);
                    ldsm_x4_t(oob ? zero_u32 : addr, b[h]);
                }
#pragma unroll
                for (int mb = 0; mb < MB; ++mb)
#pragma unroll
                    for (int nb = 0; nb < 4; ++nb)
                        mma_bf16(acc[t][mb][nb], a[mb], b[nb >> 1][(nb & 1) * 2], b[nb >> 1][(nb & 1) * 2 + 1]);
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[slot]);
    }
    // ---- partial[cta][tap][m][n]: c0,c1 = (m = gid, n = 2*t4, +1), c2,c3 = (m = gid + 8, ...)
    const int gid = lane >> 2, t4 = lane & 3;
    float* pbase = args.partial + static_cast<size_t>(blockIdx.x) * 16 * MC * NF;
#pragma unroll
    for (int t = 0; t < 2; ++t) {
        const int tap = ky * 4 + kx0 + t;
#pragma unroll
        for (int mb = 0; mb < MB; ++mb)
#pragma unroll
            for (int nb = 0; nb < 4; ++nb) {
                float* p = pbase + (static_cast<size_t>(tap) * MC + mb * 16 + gid) * NF + nb * 8 + 2 * t4;
                *reinterpret_cast<float2*>(p) = make_float2(acc[t][mb][nb][0], acc[t][mb][nb][1]);
                *reinterpret_cast<float2*>(p + 8 * NF) = make_float2(acc[t][mb][nb][2], acc[t][mb][nb][3]);
            }
    }
}

int sm_count_w() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    }
    return n;
}

template <int MC>
int launch_wt(const WgradThinArgs& a, int grid, cudaStream_t stream) {
    using Cfg = WtCfg<MC>;
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(wgrad_thin_kernel<MC>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmem) !=
            cudaSuccess)
            return -1;
        attr_set = true;
    }
    note_launch();
    wgrad_thin_kernel<MC><<<grid, kWtThreads, Cfg::kSmem, stream>>>(a);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

}  // namespace

bool wgrad_thin_supported(int cH, int cW, int Mc, int Nf) {
    return Nf == 32 && (Mc == 32 || Mc == 64) && cW >= 16 && cW <= 64 && (cW & (cW - 1)) == 0 && cH % (kBandPix / cW) == 0;
}
int wgrad_thin_ctas(int nimg, int cH, int cW) {
    const int units = nimg * (cH / (kBandPix / cW));
    return units < sm_count_w() ? units : sm_count_w();
}

// Same contract as launch_wgrad (sg_conv_umma.cuh); `partial` must hold wgrad_thin_ctas()*16*Mc*Nf floats.
int launch_wgrad_thin(const __nv_bfloat16* coarse, const __nv_bfloat16* fine, int nimg, int cH, int cW, int Mc, int Nf,
                      float* partial, float* dW, int accumulate, cudaStream_t stream) {
    WgradThinArgs a;
    memset(&a, 0, sizeof(a));
    a.cH = cH;
    a.cW = cW;
    a.nimg = nimg;
    a.partial = partial;
    const int R = kBandPix / cW;
    if (make_map_2d(&a.cmap, coarse, Mc, static_cast<uint64_t>(nimg) * cH * cW, Mc, Mc, kBandPix)) return -1;
    if (make_map_nhwc(&a.fmap, fine, nimg, 2 * cH, 2 * cW, Nf, 1, 0, 0, 32, 2 * cW, 2 * R + 2, 1)) return -1;
    const int grid = wgrad_thin_ctas(nimg, cH, cW);
    const int rc = Mc == 32 ? launch_wt<32>(a, grid, stream) : launch_wt<64>(a, grid, stream);
    if (rc) return rc;
    wgrad_reduce(partial, dW, grid, Mc, Nf, accumulate, stream);
    return 0;
}

}  // namespace sg
