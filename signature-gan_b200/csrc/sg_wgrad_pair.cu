// sg_wgrad_pair.cu — weight gradient of the Generator's last ConvTranspose2d block (gen…:46-54: 32 -> 32 channels,
// 32 x 32 -> 64 x 64) on tcgen05, replacing the mma.sync kernel of sg_wgrad_thin.cu for that shape (0.50 ms per step at
// B = 4096: one HMMA per 12-16 clk per sub-partition plus its ldmatrix feed, 2.5 TB/s).
//
//   dW[m][n][ky][kx] = sum over coarse pixels (img, iy, ix) of coarse[img,iy,ix][m] * fine[img, 2iy-1+ky, 2ix-1+kx][n]
//
// The contraction runs over pixels, so both operands are MN-major straight from NHWC. What made tcgen05 unattractive
// (32-wide operands, 16 shifted copies of the fine tensor) is removed by the pixel-pair formulation of
// sg_convs2_thin.cu, transposed:
//   * K index = (coarse row iy, fine pixel PAIR j) of a tile of 4 coarse rows; a fine pair row is 64 values [px][n] =
//     one 128-byte swizzle row, loaded once per vertical tap ky (row-parity plane, rows iy-1 / iy / iy / iy+1).
//     The boxes of ky = 1, 2 (and of ky = 0, 3) sit side by side as the two 64-wide atoms of ONE M = 128 operand:
//     M = [ky][px][n].
//   * the horizontal taps come from the COARSE side, which is 4x smaller: three copies of the coarse tile shifted by
//     s = 0, -1, +1 pixels (TMA boxes starting at x = s; the column outside the image is zero-filled = padding) stacked
//     along N = [s][m] = 96. Pair j / pixel half px meets coarse pixel ix = j + s: (s 0, px 0) -> kx 1, (s 0, px 1) ->
//     kx 2, (s -1, px 0) -> kx 3, (s +1, px 1) -> kx 0; the other two (s, px) combinations are unused columns.
//   * 16 tcgen05.mma (128 x 96 x 16) per 128 coarse pixels into two accumulators that live in TMEM for the whole
//     launch; every CTA writes one [16][32][32] partial, folded by the existing deterministic reduction.
#include "sg_conv_umma.cuh"
#include "sg_kernels.cuh"
#include "sg_umma.cuh"

#include <cstdio>
#include <cstring>

namespace sg {

int make_map_tiled(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box, int swizzle_bytes);
void wgrad_reduce(const float* partial, float* dW, int S, int M, int N, int accumulate, cudaStream_t stream,
                  const float* bias_partial = nullptr, float* dbias = nullptr, int SB = 0);

namespace {

constexpr int kWpThreads = 64 + 32 * 4;  // warp 0 TMA producer, warp 1 MMA issuer, warps 2..5 epilogue (once, at the end)
constexpr int kWpFineBox = 128 * 128;    // 128 pair rows (4 rows x 32 pairs, or 8 rows x 16 pairs) x 128 bytes
constexpr int kWpStages = 2;
// MC = channels of the coarse tensor: 32 (the last block: 64-byte pixel rows, 64-byte swizzle, N = 3 x 32 = 96) or 64 (the
// block before it: 128-byte pixel rows, 128-byte swizzle, N = 3 x 64 = 192). PPR = pixel pairs (= coarse pixels) per
// 32- or 16-wide row block of a tile; a tile is 128 / PPR coarse rows.
template <int MC>
struct WpCfg {
    static constexpr int kCoarseBox = 128 * MC * 2;  // 128 coarse pixels x MC channels bf16
    static constexpr int kStageBytes = 4 * kWpFineBox + 3 * kCoarseBox;
    static constexpr int kSmemBytes = kWpStages * kStageBytes + 1024 + 256;
    static constexpr int kN = 3 * MC;                 // [s][m]
    static constexpr int kGrpCols = MC == 32 ? 128 : 256;  // TMEM column distance between the two accumulators
    static constexpr int kTmemCols = 2 * kGrpCols;
};

struct WgradPairArgs {
    CUtensorMap fmap[2];  // fine tensor, row-parity planes: [64 (px,n)][32 pairs][cH rows][N]
    CUtensorMap cmap;     // coarse tensor NHWC [32][32][cH][N], box {32, 32, 4, 1}
    int cH, nimg, total_tiles;
    int halves;      // PPR-pixel column blocks per coarse row: 1, or 2 (64-wide coarse rows of the 128 x 128 model)
    float* partial;  // [gridDim.x][16][MC][32]
};

template <int MC, int PPR>
__global__ void __launch_bounds__(kWpThreads, 1) wgrad_pair_kernel(const __grid_constant__ WgradPairArgs args) {
    using Cfg = WpCfg<MC>;
    constexpr int kWpStageBytes = Cfg::kStageBytes, kWpCoarseBox = Cfg::kCoarseBox, kWpTmemCols = Cfg::kTmemCols;
    constexpr int kRows = 128 / PPR;  // coarse rows per tile
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kWpStages * kWpStageBytes);
    uint64_t* empty_bar = full_bar + kWpStages;
    uint64_t* accum_bar = empty_bar + kWpStages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

    const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
    const int tpi = args.cH / kRows;
    const int total_tiles = args.total_tiles;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&args.fmap[0]);
        tma_prefetch_desc(&args.fmap[1]);
        tma_prefetch_desc(&args.cmap);
        for (int s = 0; s < kWpStages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(accum_bar, 1);
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, kWpTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ---------------- TMA producer ----------------
        const bool issuer = elect_one();
        int s = 0;
        uint32_t ph = 0;
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
            // tile = (image, 4 coarse rows, 32-column block): pair j of the block meets coarse pixel x0 + j + s, so the
            // shifted coarse boxes of an inner block edge read the neighbouring block's real pixels, and only the image
            // border is zero-filled
            const int tb = t / args.halves, x0 = (t - tb * args.halves) * PPR;
            const int n0 = tb / tpi, y0 = (tb - n0 * tpi) * kRows;
            mbar_wait(&empty_bar[s], ph ^ 1);
            if (issuer) {
                mbar_arrive_expect_tx(&full_bar[s], kWpStageBytes);
                uint8_t* sf = smem + s * kWpStageBytes;
                uint8_t* sc = sf + 4 * kWpFineBox;
                // atoms [ky 1 | ky 2] and [ky 0 | ky 3]; fine row 2 iy - 1 + ky: ky 0 -> odd plane row iy - 1, 1 -> even
                // plane row iy, 2 -> odd plane row iy, 3 -> even plane row iy + 1
                tma_load_4d(sf + 0 * kWpFineBox, &args.fmap[0], &full_bar[s], 0, x0, y0, n0);      // ky 1
                tma_load_4d(sf + 1 * kWpFineBox, &args.fmap[1], &full_bar[s], 0, x0, y0, n0);      // ky 2
                tma_load_4d(sf + 2 * kWpFineBox, &args.fmap[1], &full_bar[s], 0, x0, y0 - 1, n0);  // ky 0
                tma_load_4d(sf + 3 * kWpFineBox, &args.fmap[0], &full_bar[s], 0, x0, y0 + 1, n0);  // ky 3
                tma_load_4d(sc + 0 * kWpCoarseBox, &args.cmap, &full_bar[s], 0, x0, y0, n0);       // s = 0
                tma_load_4d(sc + 1 * kWpCoarseBox, &args.cmap, &full_bar[s], 0, x0 - 1, y0, n0);   // s = -1
                tma_load_4d(sc + 2 * kWpCoarseBox, &args.cmap, &full_bar[s], 0, x0 + 1, y0, n0);   // s = +1
            }
            if (++s == kWpStages) {
                s = 0;
                ph ^= 1;
            }
        }
    } else if (warp == 1) {
        // ---------------- MMA issuer: both operands MN-major ----------------
        constexpr uint32_t idesc = make_idesc_bf16(128, Cfg::kN, 1, 1);
        const bool issuer = elect_one();
        int s = 0;
        uint32_t ph = 0;
        bool first = true;
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
            mbar_wait(&full_bar[s], ph);
            tc_fence_after();
            const uint32_t f_addr = smem_u32(smem + s * kWpStageBytes);
            const uint32_t c_addr = f_addr + 4 * kWpFineBox;
            if (issuer) {
#pragma unroll
                for (int grp = 0; grp < 2; ++grp) {
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        // A: 16 pair rows x 128 B per K step; two 64-wide atoms one box apart. B: 16 pixel rows x 64 B per
                        // K step (64-byte swizzle: 8-row groups of 512 B); three 32-wide blocks one coarse box apart.
                        const uint64_t da = make_smem_desc(f_addr + grp * 2 * kWpFineBox + k * 2048, kWpFineBox, 1024, kLayoutSW128);
                        // MC = 32: 16 pixel rows x 64 B per K step (64-byte swizzle: 8-row groups of 512 B), three 32-wide
                        // blocks one coarse box apart; MC = 64: 16 rows x 128 B, three 64-wide atoms one box apart
                        const uint64_t db = MC == 32 ? make_smem_desc(c_addr + k * 1024, kWpCoarseBox, 512, kLayoutSW64)
                                                     : make_smem_desc(c_addr + k * 2048, kWpCoarseBox, 1024, kLayoutSW128);
                        umma_bf16_ss(tmem_base + grp * Cfg::kGrpCols, da, db, idesc, !first || k != 0);
                    }
                }
                umma_commit(&empty_bar[s]);
            }
            first = false;
            if (++s == kWpStages) {
                s = 0;
                ph ^= 1;
            }
        }
        if (issuer) umma_commit(accum_bar);
    } else {
        // ---------------- Epilogue (once): TMEM -> this CTA's partial [16][MC m][32 n] ----------------
        const int q = warp & 3;  // TMEM lane quarter: rows [ky half][px][n] -> q = kyh * 2 + px, lane = n
        const int kyh = q >> 1, px = q & 1;
        const bool any = static_cast<int>(blockIdx.x) < total_tiles;
        if (any) {
            mbar_wait(accum_bar, 0);
            tc_fence_after();
        }
        float* dst = args.partial + static_cast<size_t>(blockIdx.x) * 16 * MC * 32;
#pragma unroll 1
        for (int grp = 0; grp < 2; ++grp) {
            const int ky = grp == 0 ? (kyh == 0 ? 1 : 2) : (kyh == 0 ? 0 : 3);
#pragma unroll 1
            for (int si = 0; si < 3; ++si) {
                // (s, px) -> kx: (0,0) 1, (0,1) 2, (-1,0) 3, (+1,1) 0; the other combinations are unused products
                int kx;
                if (si == 0) kx = px == 0 ? 1 : 2;
                else if (si == 1) kx = px == 0 ? 3 : -1;
                else kx = px == 1 ? 0 : -1;
#pragma unroll 1
                for (int mb = 0; mb < MC / 32; ++mb) {  // 32 columns (coarse channels m) at a time
                    uint32_t v[32];
                    if (any) {
                        tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + grp * Cfg::kGrpCols + si * MC + mb * 32,
                                      v);
                        tmem_ld_wait();
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = 0;
                    }
                    if (kx >= 0) {
                        float* o = dst + (static_cast<size_t>(ky * 4 + kx) * MC + mb * 32) * 32 + lane;  // [tap][m][n = lane]
#pragma unroll
                        for (int m = 0; m < 32; ++m) o[m * 32] = __uint_as_float(v[m]);
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, kWpTmemCols);
}

int sm_count_wp() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    }
    return n;
}

}  // namespace

// Shapes: fine tensor [nimg][2 cH][2 cW][32] (Nf = 32), coarse tensor [nimg][cH][cW][Mc] with
//   Mc = 32, cW = 32 or 64  (the last block of the 64 x 64 / 128 x 128 generator), tiles of 4 rows x 32 pixels,
//   Mc = 64, cW = 16 or 32  (the block before it), tiles of 8 rows x 16 pixels / 4 rows x 32 pixels.
bool wgrad_pair_supported(int cH, int cW, int Mc, int Nf) {
    static const bool on = [] {
        const char* e = getenv("SIGGAN_WGRAD_PAIR");
        return !(e && e[0] == '0');
    }();
    if (!on || Nf != 32) return false;
    if (Mc == 32) return (cW == 32 || cW == 64) && cH >= 4 && cH % 4 == 0;
    if (Mc == 64) return (cW == 16 && cH % 8 == 0) || (cW == 32 && cH % 4 == 0);
    return false;
}
static int wp_ppr(int cW) { return cW == 16 ? 16 : 32; }
int wgrad_pair_ctas(int nimg, int cH, int cW) {
    const int ppr = wp_ppr(cW);
    const int tiles = nimg * (cH / (128 / ppr)) * (cW / ppr);
    return tiles < sm_count_wp() ? tiles : sm_count_wp();
}

template <int MC, int PPR>
static int launch_wp(const WgradPairArgs& a, int grid, cudaStream_t stream) {
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(wgrad_pair_kernel<MC, PPR>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 WpCfg<MC>::kSmemBytes) != cudaSuccess)
            return -1;
        attr_set = true;
    }
    wgrad_pair_kernel<MC, PPR><<<grid, kWpThreads, WpCfg<MC>::kSmemBytes, stream>>>(a);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

// Same contract as launch_wgrad; `partial` must hold wgrad_pair_ctas() * 16 * Mc * 32 floats.
int launch_wgrad_pair(const __nv_bfloat16* coarse, const __nv_bfloat16* fine, int nimg, int cH, int cW, int Mc,
                      float* partial, float* dW, int accumulate, cudaStream_t stream) {
    WgradPairArgs a;
    memset(&a, 0, sizeof(a));
    const int ppr = wp_ppr(cW), rows = 128 / ppr;
    a.cH = cH;
    a.nimg = nimg;
    a.halves = cW / ppr;
    a.total_tiles = nimg * (cH / rows) * a.halves;
    a.partial = partial;
    const int fW = 2 * cW, fH = 2 * cH;
    const uint64_t row_bytes = static_cast<uint64_t>(fW) * 32 * 2;
    for (int py = 0; py < 2; ++py) {
        const uint64_t dims[4] = {64, static_cast<uint64_t>(fW / 2), static_cast<uint64_t>(cH), static_cast<uint64_t>(nimg)};
        const uint64_t strides[3] = {128, 2 * row_bytes, static_cast<uint64_t>(fH) * row_bytes};
        const uint32_t box[4] = {64, static_cast<uint32_t>(ppr), static_cast<uint32_t>(rows), 1};
        if (make_map_tiled(&a.fmap[py], reinterpret_cast<const char*>(fine) + py * row_bytes, 4, dims, strides, box, 128))
            return -1;
    }
    {
        const uint64_t px_bytes = static_cast<uint64_t>(Mc) * 2;
        const uint64_t dims[4] = {static_cast<uint64_t>(Mc), static_cast<uint64_t>(cW), static_cast<uint64_t>(cH),
                                  static_cast<uint64_t>(nimg)};
        const uint64_t strides[3] = {px_bytes, static_cast<uint64_t>(cW) * px_bytes, static_cast<uint64_t>(cH) * cW * px_bytes};
        const uint32_t box[4] = {static_cast<uint32_t>(Mc), static_cast<uint32_t>(ppr), static_cast<uint32_t>(rows), 1};
        if (make_map_tiled(&a.cmap, coarse, 4, dims, strides, box, Mc == 32 ? 64 : 128)) return -1;
    }
    const int grid = wgrad_pair_ctas(nimg, cH, cW);
    note_launch();
    int rc;
    if (Mc == 32) rc = launch_wp<32, 32>(a, grid, stream);
    else if (ppr == 16) rc = launch_wp<64, 16>(a, grid, stream);
    else rc = launch_wp<64, 32>(a, grid, stream);
    if (rc) return -1;
    wgrad_reduce(partial, dW, grid, Mc, 32, accumulate, stream);
    return 0;
}

}  // namespace sg
