// sg_conv_umma.cu — tcgen05 implicit-GEMM kernels for the GAN's 4x4 stride-2 (transposed) convolutions.
//
// Layout: activations NHWC bf16, weights packed [Cout][ky*4+kx][Cin] bf16 (K contiguous).
// One CTA computes a 128 x BN output tile. Warp 0 = TMA producer, warp 1 = MMA issuer (+TMEM
// owner), warps 2..5 = epilogue (TMEM -> registers -> global). The im2col gather is done by TMA
// itself: for every filter tap the producer loads one 4-D box {BK channels, GW, BH, BN_img} of the
// (parity-split) input at a shifted coordinate; out-of-bounds rows/cols are zero-filled by TMA, which
// implements the padding. The 128 box rows land as a canonical K-major swizzled UMMA operand.
#include "sg_conv_umma.cuh"
#include "sg_umma.cuh"
#include "sg_kernels.cuh"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

namespace sg {

static thread_local char g_err[512] = "";
const char* umma_last_error() { return g_err; }
#define SG_FAIL(...)                                \
    do {                                            \
        snprintf(g_err, sizeof(g_err), __VA_ARGS__); \
        return -1;                                  \
    } while (0)

// ----------------------------------------------------------------------------
// Tensor maps
// ----------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

static CUtensorMapSwizzle swizzle_for(uint32_t inner_bytes) {
    return inner_bytes >= 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
}

int make_map_2d(CUtensorMap* m, const void* base, uint64_t inner, uint64_t outer, uint64_t row_stride_elems,
                uint32_t box_inner, uint32_t box_outer) {
    EncodeTiledFn enc = get_encode();
    if (!enc) SG_FAIL("cuTensorMapEncodeTiled entry point unavailable");
    cuuint64_t gdim[2] = {inner, outer};
    cuuint64_t gstr[1] = {row_stride_elems * 2};
    cuuint32_t box[2] = {box_inner, box_outer};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstr, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(box_inner * 2), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        SG_FAIL("cuTensorMapEncodeTiled(2d) failed: %d (inner=%llu outer=%llu box=%u,%u)", (int)r,
                (unsigned long long)inner, (unsigned long long)outer, box_inner, box_outer);
    return 0;
}

int make_map_nhwc(CUtensorMap* m, const void* base, int N, int H, int W, int C, int step, int yp, int xp,
                  uint32_t box_c, uint32_t box_w, uint32_t box_h, uint32_t box_n) {
    EncodeTiledFn enc = get_encode();
    if (!enc) SG_FAIL("cuTensorMapEncodeTiled entry point unavailable");
    const char* b = static_cast<const char*>(base) + (static_cast<size_t>(yp) * W + xp) * C * 2;
    cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)(W / step), (cuuint64_t)(H / step), (cuuint64_t)N};
    cuuint64_t gstr[3] = {(cuuint64_t)step * C * 2, (cuuint64_t)step * W * C * 2, (cuuint64_t)H * W * C * 2};
    cuuint32_t box[4] = {box_c, box_w, box_h, box_n};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<char*>(b), gdim, gstr, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(box_c * 2), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        SG_FAIL("cuTensorMapEncodeTiled(4d) failed: %d (N=%d H=%d W=%d C=%d step=%d box=%u,%u,%u,%u)", (int)r, N, H, W,
                C, step, box_c, box_w, box_h, box_n);
    return 0;
}

// Generic tiled map over bf16 elements: dims / box innermost first, strides_bytes[i] = stride of dimension i + 1.
int make_map_tiled(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box, int swizzle_bytes) {
    EncodeTiledFn enc = get_encode();
    if (!enc) SG_FAIL("cuTensorMapEncodeTiled entry point unavailable");
    cuuint64_t gdim[5], gstr[4];
    cuuint32_t bx[5], es[5];
    for (int i = 0; i < rank; ++i) {
        gdim[i] = dims[i];
        bx[i] = box[i];
        es[i] = 1;
        if (i + 1 < rank) gstr[i] = strides_bytes[i];
    }
    const CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                  : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                                        : CU_TENSOR_MAP_SWIZZLE_NONE;
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(base), gdim, gstr, bx, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) SG_FAIL("cuTensorMapEncodeTiled(rank %d) failed: %d", rank, (int)r);
    return 0;
}

// ----------------------------------------------------------------------------
// Forward / dgrad implicit GEMM — persistent, warp-specialised
//
// One CTA per SM loops over 128 x BN output tiles (static round-robin schedule). Three pipelines:
//   smem ring   : TMA producer (warp 0)  -> full/empty mbarriers  -> MMA issuer (warp 1)
//   TMEM ring   : two BN-column accumulators; MMA issuer -> tfull/tempty mbarriers -> epilogue warps
//   tile loop   : every role walks the same tile sequence, so no scheduler traffic is needed
// The epilogue of tile j therefore overlaps the main loop of tile j+1, and the ~2.7k-cycle CTA prologue
// (barrier init, TMEM allocation, descriptor prefetch) is paid once per launch instead of once per tile.
// Tile order: the tiles that read the same A rows (other output-channel tiles, other ConvT phases) are
// adjacent in the schedule so that they run at the same time on neighbouring SMs and share A through L2.
// ----------------------------------------------------------------------------
constexpr int kTileM = 128;
constexpr int kEpiWarps = 8;                      // two warps per TMEM lane quarter, each takes half of the columns
constexpr int kThreads = 64 + 32 * kEpiWarps;     // warp 0 producer, warp 1 MMA, warps 2..9 epilogue
constexpr int kEpiStageBytes = 32 * 128;          // per epilogue warp: 32 rows x 32 fp32 columns, swizzled

// kPair (transposed convolution, BN <= 128): one schedule unit = both horizontal output parities (px = 0, 1) of a
// (128-input-pixel tile, py) pair, accumulated side by side in TMEM, so that the epilogue of a row writes the two
// horizontally adjacent output pixels — a contiguous run — instead of every other pixel.
template <int BN, int BK, bool kPair>
struct ConvCfg {
    static constexpr int kABytes = kTileM * BK * 2;
    static constexpr int kBBytes = BN * BK * 2;
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kAccCols = kPair ? 2 * BN : BN;                  // columns of one accumulator buffer
    static constexpr int kTmemCols = 2 * kAccCols < 32 ? 32 : 2 * kAccCols;  // double-buffered accumulator
    // Thin tiles are epilogue/latency-bound rather than MMA-bound: run two persistent CTAs per SM.
    // Fat tiles (BN >= 128) run as clusters of two CTAs that work on M-adjacent tiles and share the weight tile: each
    // CTA loads half of B and TMA-multicasts it to both, which removes a quarter to a third of the shared-memory
    // fill traffic — the resource these kernels are bound by (L2 -> SM at ~64 B/clk against 128x256x64 MMA stages).
    static constexpr int kCluster = BN >= 128 ? 2 : 1;
    static constexpr int kCtasPerSm = BN <= 64 ? 2 : 1;   // (2 x kTmemCols <= 512 holds for BN <= 64, paired or not)
    static constexpr int kStagesFit = ((kCtasPerSm == 2 ? 76 : 192) * 1024) / kStageBytes;
    static constexpr int kStages = kStagesFit > 8 ? 8 : kStagesFit;
    static constexpr int kColsPerWarp = kAccCols >= 64 ? kAccCols / 2 : kAccCols;
    static constexpr int kActiveEpiWarps = kAccCols >= 64 ? 8 : 4;
    static constexpr int kSmemBytes = kStages * kStageBytes + kEpiWarps * kEpiStageBytes + 1024 /*align*/ +
                                      256 /*barriers*/;
    static constexpr uint32_t kLayout = (BK == 64) ? kLayoutSW128 : kLayoutSW64;
    static constexpr uint32_t kSBO = 8 * BK * 2;  // 8 rows of one swizzle atom
};

__device__ __forceinline__ float bf16_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t u) { return __uint_as_float(u & 0xFFFF0000u); }
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}

struct TileCoord {
    int tile_m, tile_n, phase;
};
__device__ __forceinline__ TileCoord decode_tile(int t, int n_tiles, int phases) {
    TileCoord c;
    c.tile_n = t % n_tiles;
    const int r = t / n_tiles;
    c.phase = r % phases;
    c.tile_m = r / phases;
    return c;
}

// 8 floats of a per-column vector (uniform across the warp, so these are broadcast loads that hit L1).
__device__ __forceinline__ void load_vec8(const float* p, bool vec_ok, float (&v)[8]) {
    if (vec_ok) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(p));
        const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
        v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = __ldg(p + j);
    }
}

// Epilogue staging buffer of one warp: 32 rows x 128 bytes (fp32 accumulators); 16-byte piece k of row r lives at piece
// k ^ (r & 7), which makes both the row-per-lane stores (a lane writes its row's 8 pieces) and the write-back reads
// (4 lanes per row, 2 pieces each) bank-conflict free.
__device__ __forceinline__ uint32_t epi_off(int row, int k) { return row * 128 + ((k ^ (row & 7)) << 4); }

// kStats (paired transposed convolution only; there BN == N_total): the epilogue also accumulates per-channel sum and
// sum of squares of the stored (bf16-rounded) outputs — the BatchNorm batch statistics of gen…:58 — and the CTA writes
// one partial row stats_partial[blockIdx.x][2][BN], so that no separate pass over the convolution output is needed.
template <int BN, int BK, bool kPair, bool kStats>
__global__ void __launch_bounds__(kThreads, (kStats && (BN > 64 || (kPair && BN == 64))) ? 1 : ConvCfg<BN, BK, kPair>::kCtasPerSm)  // statistics: +32..64 registers
conv_umma_kernel(const __grid_constant__ ConvGemmArgs args) {
    using Cfg = ConvCfg<BN, BK, kPair>;
    constexpr int STAGES = Cfg::kStages;
    // Epilogues with fused reductions (training-mode ConvT forward: BatchNorm statistics; BatchNorm-gated data gradient:
    // BatchNorm-backward sums) have no bias / affine / activation / dropout mask (launch_conv_gemm checks), which keeps
    // the thin variants inside the 96-register budget of two CTAs per SM
    constexpr bool kGateStats = kStats;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* epi_smem = smem + STAGES * Cfg::kStageBytes;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(epi_smem + kEpiWarps * kEpiStageBytes);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tfull_bar = empty_bar + STAGES;   // [2] accumulator ready for the epilogue
    uint64_t* tempty_bar = tfull_bar + 2;       // [2] accumulator drained, MMA may overwrite
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

    const int warp = uniform_warp_idx();
    const int lane = threadIdx.x & 31;
    const int mode = args.mode;
    const int cc_n = args.Cin / BK;
    const int taps = (mode == kConvS2) ? 16 : (mode == kConvT ? 4 : 1);
    // schedule units per (tile_m, tile_n): 4 parity phases, or 2 vertical parities when px is paired in the unit
    const int phases = (mode == kConvT) ? (kPair ? 2 : 4) : 1;
    constexpr int kSub = kPair ? 2 : 1;  // accumulations per unit
    const int num_k = taps * cc_n;
    const int R = args.GH * args.GW;
    const int n_tiles = args.N_total / BN;
    constexpr int CL = Cfg::kCluster;
    // schedule units are handed to clusters; CTA `rank` of a cluster takes M tile CL * unit.tile_m + rank (a tile
    // index past the end is harmless: TMA zero-fills, the epilogue skips rows >= M_total)
    const int m_tiles = ((args.M_total + kTileM - 1) / kTileM + CL - 1) / CL;
    const int total_tiles = m_tiles * n_tiles * phases;
    const int cta_rank = CL > 1 ? static_cast<int>(cluster_ctarank()) : 0;
    const int first_unit = blockIdx.x / CL, unit_step = gridDim.x / CL;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&args.amap[0]);
        tma_prefetch_desc(&args.bmap);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], CL);   // every CTA of the cluster must have consumed the stage (multicast B)
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tfull_bar[a], 1);
            mbar_init(&tempty_bar[a], Cfg::kActiveEpiWarps);
        }
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, Cfg::kTmemCols);
    tc_fence_before();
    if (CL > 1) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ---------------- TMA producer: whole warp in uniform control flow, one elected lane issues ----------------
        // A thin stage (128 x 128 x 64) is 256 MMA cycles, so the producer's instruction stream is on the critical path:
        // everything per-stage is tabulated on the host (args.tab, read with uniform constant loads) and the loop runs
        // converged so that coordinates, shared-memory and barrier addresses stay in uniform registers.
        const bool issuer = elect_one();
        // B tile: whole (single CTA) or this CTA's half, multicast to both CTAs of the cluster
        auto load_b = [&](uint8_t* sb, uint64_t* bar, int kc, int nc) {
            if (CL == 1)
                tma_load_2d(sb, &args.bmap, bar, kc, nc);
            else
                tma_load_2d_mc(sb + cta_rank * (Cfg::kBBytes / CL), &args.bmap, bar, kc, nc + cta_rank * (BN / CL),
                               static_cast<uint16_t>((1u << CL) - 1));
        };
        int s = 0;
        uint32_t ph = 0;  // stage index / phase bit of the running stage counter
        const int tpi = R >= kTileM ? R / kTileM : 1, bh = kTileM / args.GW, ipt = R >= kTileM ? 1 : kTileM / R;
        for (int t = first_unit; t < total_tiles; t += unit_step) {
            TileCoord tc = decode_tile(t, n_tiles, phases);
            tc.tile_m = tc.tile_m * CL + cta_rank;
            int n0 = 0, y0 = 0;
            if (mode != kPlain) {
                if (R >= kTileM) {
                    n0 = tc.tile_m / tpi;
                    y0 = (tc.tile_m - n0 * tpi) * bh;
                } else {
                    n0 = tc.tile_m * ipt;
                }
            }
            const int nb = tc.tile_n * BN;
#pragma unroll 1
            for (int sub = 0; sub < kSub; ++sub) {
                const int ph4 = mode == kConvT ? (kPair ? tc.phase * 2 + sub : tc.phase) : 0;
                const int tbase = ph4 * num_k;
#pragma unroll 1
                for (int it = 0; it < num_k; ++it) {
                    mbar_wait(&empty_bar[s], ph ^ 1);
                    if (issuer) {
                        uint8_t* sa = smem + s * Cfg::kStageBytes;
                        uint8_t* sb = sa + Cfg::kABytes;
                        mbar_arrive_expect_tx(&full_bar[s], Cfg::kStageBytes);
                        if (mode == kPlain) {
                            tma_load_2d(sa, &args.amap[0], &full_bar[s], it * BK, tc.tile_m * kTileM);
                            load_b(sb, &full_bar[s], it * BK, nb);
                        } else {
                            const int4 e = args.tab[tbase + it];
                            tma_load_4d(sa, &args.amap[(e.y >> 4) & 3], &full_bar[s], e.x, (e.y & 3) - 1,
                                        y0 + ((e.y >> 2) & 3) - 1, n0);
                            load_b(sb, &full_bar[s], e.z, nb);
                        }
                    }
                    if (++s == STAGES) {
                        s = 0;
                        ph ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ---------------- MMA issuer: whole warp in uniform control flow, one elected lane issues ----------------
        constexpr uint32_t idesc = make_idesc_bf16(kTileM, BN, 0, 0);
        const bool issuer = elect_one();
        int s = 0;
        uint32_t ph = 0;
        int j = 0;  // local tile counter
        for (int t = first_unit; t < total_tiles; t += unit_step, ++j) {
            const int acc = j & 1;
            mbar_wait(&tempty_bar[acc], ((j >> 1) & 1) ^ 1);
            tc_fence_after();
#pragma unroll 1
            for (int sub = 0; sub < kSub; ++sub) {
                const uint32_t tmem_d = tmem_base + acc * Cfg::kAccCols + sub * BN;
#pragma unroll 1
                for (int it = 0; it < num_k; ++it) {
                    mbar_wait(&full_bar[s], ph);
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(smem + s * Cfg::kStageBytes);
                    const uint32_t b_addr = a_addr + Cfg::kABytes;
                    if (issuer) {
#pragma unroll
                        for (int k = 0; k < BK / 16; ++k) {
                            const uint64_t da = make_smem_desc(a_addr + k * 32, 0, Cfg::kSBO, Cfg::kLayout);
                            const uint64_t db = make_smem_desc(b_addr + k * 32, 0, Cfg::kSBO, Cfg::kLayout);
                            umma_bf16_ss(tmem_d, da, db, idesc, (it | k) != 0);
                        }
                        if (CL == 1)
                            umma_commit(&empty_bar[s]);
                        else
                            umma_commit_mc(&empty_bar[s], static_cast<uint16_t>((1u << CL) - 1));
                    }
                    if (++s == STAGES) {
                        s = 0;
                        ph ^= 1;
                    }
                }
            }
            if (issuer) umma_commit(&tfull_bar[acc]);
        }
    } else if (warp - 2 < Cfg::kActiveEpiWarps) {
        // ---------------- Epilogue: TMEM -> registers -> (swizzled smem transpose) -> coalesced global ----------------
        const int q = warp & 3;                       // TMEM lane quarter this warp may access
        const int half = (warp - 2) >> 2;             // which half of the unit's accumulator columns
        uint8_t* stage = epi_smem + (warp - 2) * kEpiStageBytes;
        const uint32_t stage_u32 = smem_u32(stage);
        const bool ygate = args.gate && args.gate_scale;  // BatchNorm gate: sign of y * gate_scale + gate_shift
        const bool vec_ok = (((args.bias ? reinterpret_cast<uintptr_t>(args.bias) : 0) |
                              (args.scale ? reinterpret_cast<uintptr_t>(args.scale) : 0) |
                              (args.shift ? reinterpret_cast<uintptr_t>(args.shift) : 0) |
                              (ygate ? reinterpret_cast<uintptr_t>(args.gate_scale) | reinterpret_cast<uintptr_t>(args.gate_shift) : 0) |
                              (args.mask ? reinterpret_cast<uintptr_t>(args.mask) : 0)) & 15) == 0 &&
                            (args.ldmask % 4 == 0);
        const int lgR = 31 - __clz(R), lgW = 31 - __clz(args.GW);
        const int wr_row = lane >> 2, wr_k = lane & 3;  // write-back role: 4 lanes per row, 16 bytes each
        __nv_bfloat16* const outp = static_cast<__nv_bfloat16*>(args.out);
        constexpr int NCH = Cfg::kColsPerWarp / 32;  // 32-column chunks per warp
        float st_sum[kStats ? NCH : 1][8], st_sq[kStats ? NCH : 1][8];
#pragma unroll
        for (int ci = 0; ci < (kStats ? NCH : 1); ++ci)
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) st_sum[ci][jj] = st_sq[ci][jj] = 0.f;
        // GEMM rows of this warp in unit t and, per lane, the output row (pixel index) of its GEMM row (for paired
        // units: of its px = 0 pixel)
        auto unit_rows = [&](int t, TileCoord& tc, int& row0, int& img, int& orow) {
            tc = decode_tile(t, n_tiles, phases);
            tc.tile_m = tc.tile_m * CL + cta_rank;
            row0 = tc.tile_m * kTileM + q * 32;
            const int gm = row0 + lane;
            orow = gm;
            img = 0;
            if (mode != kPlain) {
                img = gm >> lgR;
                if (mode == kConvT) {
                    const int rem = gm & (R - 1);
                    const int yh = rem >> lgW, xh = rem & (args.GW - 1);
                    const int py = kPair ? tc.phase : (tc.phase >> 1), px = kPair ? 0 : (tc.phase & 1);
                    orow = ((img * 2 * args.GH + 2 * yh + py) * 2 * args.GW) + 2 * xh + px;
                }
            }
        };
        // ALL epilogue math runs in the write-back layout: lane (wr_row, wr_k) owns columns [8 wr_k, 8 wr_k + 8) of rows
        // 8 i + wr_row (i = 0..3) of a 32 x 32 chunk. The raw fp32 accumulators are transposed through the staging
        // buffer; bias / scale / shift / dropout mask of the lane's 8 channels are loaded ONCE per chunk, before the
        // accumulator load (in the row-per-lane layout every group of 8 columns needed its own loads, each behind a
        // branch: ~16 serialised L1/L2 round trips per chunk, 3-4x the time of the MMAs they overlap), and the gate
        // (saved activation at the output position) is loaded directly in this layout, one chunk ahead.
        auto gate_load = [&](const TileCoord& tcg, int row0g, int orowg, int ci, uint4 (&g)[4]) {
            const int c0 = half * Cfg::kColsPerWarp + ci * 32;
            const int sub = kPair ? c0 / BN : 0;
            const int n_base = tcg.tile_n * BN + (kPair ? c0 - sub * BN : c0);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int r = i * 8 + wr_row;
                const int o = __shfl_sync(0xffffffffu, orowg, r) + sub;
                g[i] = make_uint4(0, 0, 0, 0);
                if (row0g + r < args.M_total)
                    g[i] = __ldg(reinterpret_cast<const uint4*>(args.gate + static_cast<size_t>(o) * args.ldo + n_base +
                                                                 wr_k * 8));
            }
        };
        const bool one_img = mode == kPlain || R >= 32;  // the warp's 32 rows lie in one image: one mask row per chunk
        // two CTAs per SM run under a 96-register cap (and have 16 epilogue warps to hide latency with): there the
        // gate is loaded at the top of its own chunk instead of one chunk ahead
        constexpr bool kGateAhead = Cfg::kCtasPerSm == 1;
        int j = 0;
        TileCoord tc;
        int row0 = 0, img = 0, orow = 0;
        uint4 gq_next[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) gq_next[i] = make_uint4(0, 0, 0, 0);
        if (first_unit < total_tiles) {
            unit_rows(first_unit, tc, row0, img, orow);
            if (kGateAhead && args.gate) gate_load(tc, row0, orow, 0, gq_next);
        }
        for (int t = first_unit; t < total_tiles; t += unit_step, ++j) {
            const int acc = j & 1;
            // the unit after this one (its rows are needed to prefetch its first gate tile)
            TileCoord tc_n = tc;
            int row0_n = row0, img_n = img, orow_n = orow;
            const bool has_next = t + unit_step < total_tiles;
            if (kGateAhead && has_next) unit_rows(t + unit_step, tc_n, row0_n, img_n, orow_n);
            mbar_wait(&tfull_bar[acc], (j >> 1) & 1);
            tc_fence_after();
#pragma unroll(kStats ? NCH : 1)
            for (int ci = 0; ci < NCH; ++ci) {
                const int c0 = half * Cfg::kColsPerWarp + ci * 32;
                const int sub = kPair ? c0 / BN : 0;
                const int n_base = tc.tile_n * BN + (kPair ? c0 - sub * BN : c0);
                const int n8 = n_base + wr_k * 8;
                float b8[8], sc8[8], sh8[8], m8[8], gs8[8], gh8[8];
                if (ygate) {
                    load_vec8(args.gate_scale + n8, vec_ok, gs8);
                    load_vec8(args.gate_shift + n8, vec_ok, gh8);
                }
                if (!kGateStats && args.bias) load_vec8(args.bias + n8, vec_ok, b8);
                if (!kGateStats && args.scale) {
                    load_vec8(args.scale + n8, vec_ok, sc8);
                    load_vec8(args.shift + n8, vec_ok, sh8);
                }
                if (!kGateStats && args.mask && one_img) {
                    const int mi = row0 < args.M_total ? (mode != kPlain ? row0 >> lgR : 0) : 0;
                    load_vec8(args.mask + static_cast<size_t>(mi) * args.ldmask + n8, vec_ok, m8);
                }
                uint4 gq[4];
                if (!kGateAhead && args.gate) gate_load(tc, row0, orow, ci, gq);
                uint32_t v[32];
                tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * Cfg::kAccCols + c0, v);
                tmem_ld_wait();
                if (c0 + 32 >= (half + 1) * Cfg::kColsPerWarp) {
                    // last chunk of this warp is in registers: hand the accumulator back to the MMA issuer
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tempty_bar[acc]);
                }
                if (kGateAhead) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) gq[i] = gq_next[i];
                    if (args.gate) {
                        if (ci + 1 < NCH)
                            gate_load(tc, row0, orow, ci + 1, gq_next);
                        else if (has_next)
                            gate_load(tc_n, row0_n, orow_n, 0, gq_next);
                    }
                }
#pragma unroll
                for (int p = 0; p < 8; ++p)
                    sts128(stage_u32 + epi_off(lane, p), v[4 * p], v[4 * p + 1], v[4 * p + 2], v[4 * p + 3]);
                __syncwarp();
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int r = i * 8 + wr_row;
                    const int o = __shfl_sync(0xffffffffu, orow, r) + sub;
                    const float4 lo4 = lds128f(stage_u32 + epi_off(r, 2 * wr_k));
                    const float4 hi4 = lds128f(stage_u32 + epi_off(r, 2 * wr_k + 1));
                    float f[8] = {lo4.x, lo4.y, lo4.z, lo4.w, hi4.x, hi4.y, hi4.z, hi4.w};
                    if (!kGateStats && args.bias) {
#pragma unroll
                        for (int jj = 0; jj < 8; ++jj) f[jj] += b8[jj];
                    }
                    if (!kGateStats && args.scale) {
#pragma unroll
                        for (int jj = 0; jj < 8; ++jj) f[jj] = fmaf(f[jj], sc8[jj], sh8[jj]);
                    }
                    if (kGateStats) {
                    } else if (args.act == kActRelu) {
#pragma unroll
                        for (int jj = 0; jj < 8; ++jj) f[jj] = fmaxf(f[jj], 0.f);
                    } else if (args.act == kActLeaky) {
#pragma unroll
                        for (int jj = 0; jj < 8; ++jj) f[jj] = f[jj] > 0.f ? f[jj] : f[jj] * args.slope;
                    }
                    if (!kGateStats && args.mask) {
                        if (one_img) {
#pragma unroll
                            for (int jj = 0; jj < 8; ++jj) f[jj] *= m8[jj];
                        } else {  // tiny grids (4 x 4): the rows of a chunk span several images
                            float m[8];
                            const int mi = row0 + r < args.M_total ? (row0 + r) >> lgR : 0;
                            load_vec8(args.mask + static_cast<size_t>(mi) * args.ldmask + n8, vec_ok, m);
#pragma unroll
                            for (int jj = 0; jj < 8; ++jj) f[jj] *= m[jj];
                        }
                    }
                    if (args.gate) {
                        const uint32_t w4[4] = {gq[i].x, gq[i].y, gq[i].z, gq[i].w};
#pragma unroll
                        for (int tt = 0; tt < 4; ++tt) {
                            float g0 = bf16_lo(w4[tt]), g1 = bf16_hi(w4[tt]);
                            if (ygate) {
                                g0 = fmaf(g0, gs8[tt * 2], gh8[tt * 2]);
                                g1 = fmaf(g1, gs8[tt * 2 + 1], gh8[tt * 2 + 1]);
                            }
                            f[tt * 2] *= g0 > 0.f ? 1.f : args.slope;
                            f[tt * 2 + 1] *= g1 > 0.f ? 1.f : args.slope;
                        }
                    }
                    const uint4 d = make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]),
                                               pack_bf16(f[6], f[7]));
                    if (row0 + r < args.M_total) {
                        *reinterpret_cast<uint4*>(outp + static_cast<size_t>(o) * args.ldo + n8) = d;
                        if (kStats) {
                            const int cs = kStats ? ci : 0;
                            const uint32_t w4[4] = {d.x, d.y, d.z, d.w};
                            const uint32_t y4[4] = {gq[i].x, gq[i].y, gq[i].z, gq[i].w};
#pragma unroll
                            for (int tt = 0; tt < 4; ++tt) {
                                const float lo = bf16_lo(w4[tt]), hi = bf16_hi(w4[tt]);
                                // forward statistics: sum, sum of squares; BatchNorm gate: sum d, sum d * y
                                const float olo = ygate ? bf16_lo(y4[tt]) : lo, ohi = ygate ? bf16_hi(y4[tt]) : hi;
                                st_sum[cs][2 * tt] += lo;
                                st_sq[cs][2 * tt] = fmaf(lo, olo, st_sq[cs][2 * tt]);
                                st_sum[cs][2 * tt + 1] += hi;
                                st_sq[cs][2 * tt + 1] = fmaf(hi, ohi, st_sq[cs][2 * tt + 1]);
                            }
                        }
                    }
                }
                __syncwarp();
            }
            if (kGateAhead) {
                tc = tc_n;
                row0 = row0_n;
                img = img_n;
                orow = orow_n;
            } else if (has_next) {
                unit_rows(t + unit_step, tc, row0, img, orow);
            }
        }
        (void)img;
        (void)stage_u32;
        if (kStats) {
            // lanes with equal wr_k hold partial sums of the same 8 channels; this warp's slot = [2][kColsPerWarp]
            float* slot = reinterpret_cast<float*>(stage);
#pragma unroll
            for (int ci = 0; ci < NCH; ++ci)
#pragma unroll
                for (int jj = 0; jj < 8; ++jj) {
                    float a = st_sum[ci][jj], b = st_sq[ci][jj];
#pragma unroll
                    for (int o = 4; o < 32; o <<= 1) {
                        a += __shfl_xor_sync(0xffffffffu, a, o);
                        b += __shfl_xor_sync(0xffffffffu, b, o);
                    }
                    if (lane < 4) {
                        slot[ci * 32 + lane * 8 + jj] = a;
                        slot[Cfg::kColsPerWarp + ci * 32 + lane * 8 + jj] = b;
                    }
                }
        }
    }
    tc_fence_before();
    if (CL > 1) cluster_sync_all(); else __syncthreads();  // (cluster: the peer may still multicast into / arrive on this CTA)
    if (kStats && warp >= 2) {
        // fold the warps that covered a channel in a fixed order. Paired units (and accumulators narrower than 64
        // columns): every active epilogue warp covered all BN channels; otherwise the warps of column half h = ch /
        // kColsPerWarp (one per TMEM lane quarter) did.
        const int e = (warp - 2) * 32 + lane;
        if (e < 2 * BN) {
            const int which = e / BN, ch = e - which * BN;
            constexpr bool kAllCover = kPair || Cfg::kAccCols < 64;
            const int h = kAllCover ? 0 : ch / Cfg::kColsPerWarp;
            const int cc = ch - h * Cfg::kColsPerWarp;
            float tot = 0.f;
#pragma unroll
            for (int wi = 0; wi < (kAllCover ? Cfg::kActiveEpiWarps : 4); ++wi)
                tot += reinterpret_cast<const float*>(epi_smem + (h * 4 + wi) * kEpiStageBytes)[which * Cfg::kColsPerWarp + cc];
            args.stats_partial[(static_cast<size_t>(blockIdx.x) * 2 + which) * BN + ch] = tot;
        }
    }
    if (warp == 1) tmem_dealloc(tmem_base, Cfg::kTmemCols);
}

static int sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    }
    return n;
}

// total_tiles counts schedule units (one per cluster); the grid is that many clusters, capped by the SM count.
template <int BN, int BK, bool kPair>
static int conv_grid(int total_tiles) {
    using Cfg = ConvCfg<BN, BK, kPair>;
    const int slots = sm_count() * Cfg::kCtasPerSm / Cfg::kCluster;
    return (total_tiles < slots ? total_tiles : slots) * Cfg::kCluster;
}

template <int BN, int BK, bool kPair, bool kStats = false>
static int launch_cfg(const ConvGemmArgs& a, int total_tiles, cudaStream_t stream) {
    using Cfg = ConvCfg<BN, BK, kPair>;
    // statistics variants exist for the paired transposed convolutions and for the BatchNorm-gated data gradients of
    // the Generator's thin levels (sg_model.cu asks conv_gemm_gate_stats_chunks first)
    constexpr bool kHasStats = kPair || (BN == 32 && BK == 32) || (BN == 64 && BK == 32) || (BN == 128 && BK == 64);
    if (kHasStats && !kStats && a.stats_partial) return launch_cfg<BN, BK, kPair, kHasStats>(a, total_tiles, stream);
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(conv_umma_kernel<BN, BK, kPair, kStats>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
        if (e != cudaSuccess) SG_FAIL("cudaFuncSetAttribute(conv_umma<%d,%d>): %s", BN, BK, cudaGetErrorString(e));
        attr_set = true;
    }
    const int grid = conv_grid<BN, BK, kPair>(total_tiles);
    note_launch();
    cudaLaunchConfig_t lc;
    memset(&lc, 0, sizeof(lc));
    lc.gridDim = dim3(grid);
    lc.blockDim = dim3(kThreads);
    lc.dynamicSmemBytes = Cfg::kSmemBytes;
    lc.stream = stream;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = Cfg::kCluster;
    attr.val.clusterDim.y = 1;
    attr.val.clusterDim.z = 1;
    lc.attrs = &attr;
    lc.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&lc, conv_umma_kernel<BN, BK, kPair, kStats>, a);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) SG_FAIL("conv_umma<%d,%d> launch: %s", BN, BK, cudaGetErrorString(e));
    return 0;
}

static bool is_pow2(int x) { return x > 0 && (x & (x - 1)) == 0; }

// sg_convt4.cu
bool convt4_supported(int inH, int inW, int Cin, int Cout);
int convt4_grid(int nimg, int inH, int inW);
int launch_convt4(const __nv_bfloat16* in, const __nv_bfloat16* w_packed, int nimg, int inH, int inW, int Cin, int Cout,
                  __nv_bfloat16* out, const float* scale, const float* shift, float* stats_partial, cudaStream_t stream);

// Rows of stats_partial a kConvT launch will write (= its grid), or 0 when the configuration cannot fuse the statistics.
int conv_gemm_stats_chunks(int nimg, int inH, int inW, int Cin, int Cout) {
    if (convt4_supported(inH, inW, Cin, Cout)) return convt4_grid(nimg, inH, inW);
    if (!(Cout == 32 || Cout == 64 || Cout == 128) || Cin % 32 != 0) return 0;
    const int cl = Cout >= 128 ? 2 : 1;
    const int units = (((nimg * inH * inW + kTileM - 1) / kTileM) + cl - 1) / cl * 2;
    const int slots = sm_count() * (Cout <= 64 ? 2 : 1) / cl;
    return (units < slots ? units : slots) * cl;
}

// sg_conv2_umma.cu
bool conv2_supported(ConvMode mode, int inH, int inW, int Cin, int Cout);
int launch_conv2(ConvMode mode, const __nv_bfloat16* in, const __nv_bfloat16* w_packed, int nimg, int inH, int inW,
                 int Cin, int Cout, const ConvGemmArgs& e, cudaStream_t stream);
static bool conv2_enabled() {  // SIGGAN_CONV2=0 keeps every layer on the one-CTA kernels (A/B comparison)
    static const bool v = [] {
        const char* e = getenv("SIGGAN_CONV2");
        return !(e && e[0] == '0');
    }();
    return v;
}

// sg_convs2_thin.cu
bool convs2_thin_supported(int inH, int inW, int Cin, int Cout);
size_t convs2_thin_scratch_bytes();
int convs2_thin_ctas(int nimg, int inH);
int launch_convs2_thin(const __nv_bfloat16* in, const __nv_bfloat16* w_packed, int nimg, int inH, int inW,
                       __nv_bfloat16* out, const __nv_bfloat16* gate, float slope, void* scratch, cudaStream_t stream,
                       const float* gate_scale, const float* gate_shift, float* stats_partial);

// Rows of stats_partial a BatchNorm-gated kConvS2 launch writes (= its grid), 0 when the shape cannot fuse them.
int conv_gemm_gate_stats_chunks(int nimg, int inH, int inW, int Cin, int Cout) {
    if (convs2_thin_supported(inH, inW, Cin, Cout)) return convs2_thin_ctas(nimg, inH);
    const int BK = (Cin % 64 == 0) ? 64 : 32;
    if (!((Cout == 32 && BK == 32) || (Cout == 64 && BK == 32) || (Cout == 128 && BK == 64)) || Cin % BK != 0) return 0;
    const int GH = inH / 2, GW = inW / 2;
    if (!is_pow2(GH) || !is_pow2(GW) || GW > kTileM) return 0;
    const int cl = Cout >= 128 ? 2 : 1;
    const int tiles = (((nimg * GH * GW + kTileM - 1) / kTileM) + cl - 1) / cl;   // schedule units (one per cluster)
    const int slots = sm_count() * (Cout <= 64 ? 2 : 1) / cl;
    return (tiles < slots ? tiles : slots) * cl;
}
static void* thin_scratch() {  // stacked weights of the thin stride-2 kernel, one buffer per device
    static void* buf[16] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 16) return nullptr;
    if (!buf[dev] && cudaMalloc(&buf[dev], convs2_thin_scratch_bytes()) != cudaSuccess) buf[dev] = nullptr;
    return buf[dev];
}

int launch_conv_gemm(ConvMode mode, const __nv_bfloat16* in, const __nv_bfloat16* w_packed, int nimg, int inH, int inW,
                     int Cin, int Cout, ConvGemmArgs a, cudaStream_t stream) {
    if (mode == kConvS2 && convs2_thin_supported(inH, inW, Cin, Cout) && !a.bias && !a.scale && !a.mask &&
        (!a.stats_partial || (a.gate && a.gate_scale)) && a.act == kActNone && a.ldo == Cout) {
        // thin fine-grid tensors (generator's last block, data gradient): pixel-pair rows + col2im epilogue
        void* scratch = thin_scratch();
        if (!scratch) SG_FAIL("convs2_thin: cannot allocate the weight scratch");
        if (launch_convs2_thin(in, w_packed, nimg, inH, inW, static_cast<__nv_bfloat16*>(a.out), a.gate, a.slope, scratch,
                               stream, a.gate_scale, a.gate_shift, a.stats_partial))
            SG_FAIL("convs2_thin launch failed: %s (%s)", cudaGetErrorString(cudaGetLastError()), umma_last_error());
        return 0;
    }
    if (mode != kPlain && !a.stats_partial && !a.gate_scale && a.ldo == Cout && conv2_enabled() &&
        conv2_supported(mode, inH, inW, Cin, Cout)) {
        // CTA-pair kernel with shared shifted-input tiles (D conv1 forward, D conv1 / conv2 data gradients)
        if (launch_conv2(mode, in, w_packed, nimg, inH, inW, Cin, Cout, a, stream))
            SG_FAIL("conv2 launch failed: %s", cudaGetErrorString(cudaGetLastError()));
        return 0;
    }
    if (mode == kConvT && convt4_supported(inH, inW, Cin, Cout) && !a.bias && !a.mask && !a.gate && a.ldo == Cout &&
        (a.scale ? a.act == kActRelu : a.act == kActNone)) {
        // thin wide-grid generator blocks: all four output parities per tile, halo-shared A, resident weights
        if (launch_convt4(in, w_packed, nimg, inH, inW, Cin, Cout, static_cast<__nv_bfloat16*>(a.out), a.scale, a.shift,
                          a.stats_partial, stream))
            SG_FAIL("convt4 launch failed: %s", cudaGetErrorString(cudaGetLastError()));
        return 0;
    }
    a.mode = mode;
    a.nimg = nimg;
    a.Cin = Cin;
    a.N_total = Cout;
    const int BK = (Cin % 64 == 0) ? 64 : 32;
    if (Cin % BK != 0) SG_FAIL("conv_gemm: Cin=%d must be a multiple of 32", Cin);
    int BN = Cout >= 256 ? 256 : Cout;
    if (!(BN == 32 || BN == 64 || BN == 128 || BN == 256) || Cout % BN != 0)
        SG_FAIL("conv_gemm: unsupported Cout=%d", Cout);
    int taps;
    if (mode == kPlain) {
        a.GH = 1;
        a.GW = kTileM;
        a.M_total = nimg;
        taps = 1;
        if (make_map_2d(&a.amap[0], in, Cin, nimg, Cin, BK, kTileM)) return -1;
    } else {
        a.GH = (mode == kConvS2) ? inH / 2 : inH;
        a.GW = (mode == kConvS2) ? inW / 2 : inW;
        taps = 16;
        if (!is_pow2(a.GH) || !is_pow2(a.GW) || a.GW > kTileM) SG_FAIL("conv_gemm: grid %dx%d unsupported", a.GH, a.GW);
        a.M_total = nimg * a.GH * a.GW;
        const int R = a.GH * a.GW;
        const uint32_t bw = a.GW;
        const uint32_t bh = R >= kTileM ? kTileM / a.GW : a.GH;
        const uint32_t bn = R >= kTileM ? 1 : kTileM / R;
        if (mode == kConvS2) {
            for (int p = 0; p < 4; ++p)
                if (make_map_nhwc(&a.amap[p], in, nimg, inH, inW, Cin, 2, p >> 1, p & 1, BK, bw, bh, bn)) return -1;
        } else {
            if (make_map_nhwc(&a.amap[0], in, nimg, inH, inW, Cin, 1, 0, 0, BK, bw, bh, bn)) return -1;
        }
    }
    if (mode != kPlain) {  // producer schedule (see ConvGemmArgs::tab)
        const int cc_n = Cin / BK, num_k = taps * cc_n / (mode == kConvT ? 4 : 1);
        const int n_ent = (mode == kConvT ? 4 : 1) * num_k;
        if (n_ent > 256) SG_FAIL("conv_gemm: schedule of %d entries does not fit", n_ent);
        for (int e = 0; e < n_ent; ++e) {
            const int ph4 = e / num_k, it = e - ph4 * num_k;
            const int tap = it / cc_n, cc = it - tap * cc_n;
            int4 v;
            v.x = cc * BK;
            if (mode == kConvS2) {
                const int ky = tap >> 2, kx = tap & 3;
                const int yp = (ky + 1) & 1, xp = (kx + 1) & 1;
                const int dy = ((ky + 1) >> 1) - 1, dx = ((kx + 1) >> 1) - 1;
                v.y = (dx + 1) | ((dy + 1) << 2) | ((yp * 2 + xp) << 4);
                v.z = it * BK;
            } else {
                const int py = ph4 >> 1, px = ph4 & 1, ty = tap >> 1, tx = tap & 1;
                const int ky = (1 - py) + 2 * ty, kx = (1 - px) + 2 * tx;
                v.y = (px - tx + 1) | ((py - ty + 1) << 2);
                v.z = (ky * 4 + kx) * Cin + cc * BK;
            }
            v.w = 0;
            a.tab[e] = v;
        }
    }
    const int cl = BN >= 128 ? 2 : 1;  // = ConvCfg::kCluster
    if (make_map_2d(&a.bmap, w_packed, (uint64_t)taps * Cin, Cout, (uint64_t)taps * Cin, BK, BN / cl)) return -1;
    const bool pair = mode == kConvT && BN <= 128;
    if (a.stats_partial && !(BN == Cout && (pair || (mode == kConvS2 && a.gate && a.gate_scale &&
                                                      conv_gemm_gate_stats_chunks(nimg, inH, inW, Cin, Cout) > 0))))
        SG_FAIL("conv_gemm: fused statistics need a paired transposed convolution or a BatchNorm-gated data gradient");
    if (a.stats_partial && (a.bias || a.scale || a.mask || a.act != kActNone || (pair && a.gate)))
        SG_FAIL("conv_gemm: an epilogue with fused reductions has no bias / affine / activation / mask");
    const int total_tiles = (((a.M_total + kTileM - 1) / kTileM + cl - 1) / cl) * (Cout / BN) *
                            (mode == kConvT ? (pair ? 2 : 4) : 1);
#define SG_DISPATCH(bn, bk)                                                   \
    if (BN == bn && BK == bk)                                                 \
        return pair ? launch_cfg<bn, bk, (bn <= 128)>(a, total_tiles, stream) \
                    : launch_cfg<bn, bk, false>(a, total_tiles, stream);
    SG_DISPATCH(256, 64)
    SG_DISPATCH(128, 64)
    SG_DISPATCH(64, 64)
    SG_DISPATCH(32, 64)
    SG_DISPATCH(256, 32)
    SG_DISPATCH(128, 32)
    SG_DISPATCH(64, 32)
    SG_DISPATCH(32, 32)
#undef SG_DISPATCH
    SG_FAIL("conv_gemm: no kernel for BN=%d BK=%d", BN, BK);
}

// ----------------------------------------------------------------------------
// Weight gradient: dW[m][tap][n] = sum_pix coarse[pix][m] * fine[2*pix - 1 + tap][n]
// Both operands are "MN-major" (the contraction runs over rows of NHWC tensors).
// ----------------------------------------------------------------------------
constexpr int kWgK = 64;  // pixels per pipeline stage
constexpr int kWgThreads = 192;  // warp 0 producer, warp 1 MMA, warps 2..5 epilogue
// Fused bias gradient (Discriminator conv blocks, disc…:51-58): dbias[m] = sum_pix coarse[pix][m] is one more product of
// the staged coarse tile, with a constant all-ones B operand (16 pixel rows x 128 bytes of bf16 1.0 — all ones, so
// neither the swizzle nor the MN-major atom layout matters) accumulated into 16 extra TMEM columns by the CTAs of the
// CTAs (tile_m, y, split), y = 0..ny-1, for the K steps `it % ny == y` — spread over all tap groups, because the extra
// product re-reads the 4 KB A operand (~32 clk against the 128 clk of a 256-column product) and a launch is as slow as
// its slowest CTA. It replaces a separate column-reduction pass over dy (0.22 ms per step at B = 4096).
constexpr int kOnesBytes = 2048;
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template <int BN>
struct WgCfg {
    static constexpr int kAtomBytes = kWgK * 128;  // 64 pixel rows x 64 channels bf16
    static constexpr int kABytes = 2 * kAtomBytes;
    static constexpr int kBBytes = (BN / 64) * kAtomBytes;
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kStages = BN == 256 ? 4 : (BN == 128 ? 3 : 4);
    static constexpr int kTmemCols = BN;
    static constexpr int kTmemColsBias = 2 * BN;  // + 16 columns for the fused bias gradient (power of two)
    static constexpr int kSmemBytes = kStages * kStageBytes + kOnesBytes + 1024 + 256;
};

template <int BN>
__global__ void __launch_bounds__(kWgThreads) wgrad_umma_kernel(const __grid_constant__ WgradArgs args) {
    using Cfg = WgCfg<BN>;
    constexpr int STAGES = Cfg::kStages;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* ones = smem + STAGES * Cfg::kStageBytes;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(ones + kOnesBytes);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* accum_bar = empty_bar + STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

    const int warp = uniform_warp_idx();
    const int lane = threadIdx.x & 31;
    const int tile_m = blockIdx.x;
    const int n_tiles = (args.Nf + BN - 1) / BN;
    const int taps = args.plain ? 1 : 16;
    const bool do_bias = args.bias_partial != nullptr;
    const int ny = gridDim.y, by = blockIdx.y;
    const uint32_t tmem_cols = do_bias ? Cfg::kTmemColsBias : Cfg::kTmemCols;
    if (do_bias) {
        for (int i = threadIdx.x; i < kOnesBytes / 4; i += kWgThreads) reinterpret_cast<uint32_t*>(ones)[i] = 0x3F803F80u;
        fence_proxy_async_smem();
    }
    // Narrow fine tensors (Nf < BN): the N dimension stacks tpc = BN/Nf filter taps, [tap][channel] — a 128 x 256 UMMA
    // reads 96 B/clk of operands from shared memory where four 128 x 64 ones read 192 B/clk, the actual bound here.
    const int tpc = args.tpc;                    // taps per CTA
    const int apt = (BN / 64) / tpc;             // 64-channel atoms per tap
    const int tap = (blockIdx.y / n_tiles) * tpc;  // first tap of this CTA
    const int tile_n = blockIdx.y % n_tiles;
    const int split = blockIdx.z;
    const int kt_begin = static_cast<int>(static_cast<long>(args.k_tiles) * split / args.splits);
    const int kt_end = static_cast<int>(static_cast<long>(args.k_tiles) * (split + 1) / args.splits);
    const int num_k = kt_end - kt_begin;
    const int R = args.GH * args.GW;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&args.cmap);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(accum_bar, 1);
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ---------------- TMA producer (whole warp, uniform control flow; one elected lane issues) ----------------
        const bool issuer = elect_one();
        // per 64-channel atom of B: tap -> parity view, shift and channel block (loop invariant)
        const CUtensorMap* a_fm[BN / 64];
        int a_dy[BN / 64], a_dx[BN / 64], a_ch[BN / 64];
#pragma unroll
        for (int a = 0; a < BN / 64; ++a) {
            const int ta = tap + a / apt, cb = a % apt;
            const int ky = ta >> 2, kx = ta & 3;
            a_fm[a] = &args.fmap[((ky + 1) & 1) * 2 + ((kx + 1) & 1)];
            a_dy[a] = ((ky + 1) >> 1) - 1;
            a_dx[a] = ((kx + 1) >> 1) - 1;
            a_ch[a] = tile_n * BN + cb * 64;
        }
        const int lg_tpi = R >= kWgK ? 31 - __clz(R / kWgK) : 0, rows_per = kWgK / args.GW, ipk = R >= kWgK ? 1 : kWgK / R;
        int s = 0;
        uint32_t ph = 0;
        for (int it = 0; it < num_k; ++it) {
            mbar_wait(&empty_bar[s], ph ^ 1);
            if (issuer) {
                uint8_t* sa = smem + s * Cfg::kStageBytes;
                uint8_t* sb = sa + Cfg::kABytes;
                mbar_arrive_expect_tx(&full_bar[s], Cfg::kStageBytes);
                const int kt = kt_begin + it;
                int n0, y0;
                if (R >= kWgK) {
                    n0 = kt >> lg_tpi;
                    y0 = (kt & ((1 << lg_tpi) - 1)) * rows_per;
                } else {
                    n0 = kt * ipk;
                    y0 = 0;
                }
                tma_load_2d(sa, &args.cmap, &full_bar[s], tile_m * 128, kt * kWgK);
                tma_load_2d(sa + Cfg::kAtomBytes, &args.cmap, &full_bar[s], tile_m * 128 + 64, kt * kWgK);
                if (args.plain) {
#pragma unroll
                    for (int a = 0; a < BN / 64; ++a)
                        tma_load_2d(sb + a * Cfg::kAtomBytes, &args.fmap[0], &full_bar[s], tile_n * BN + a * 64, kt * kWgK);
                } else {
#pragma unroll
                    for (int a = 0; a < BN / 64; ++a)
                        tma_load_4d(sb + a * Cfg::kAtomBytes, a_fm[a], &full_bar[s], a_ch[a], a_dx[a], y0 + a_dy[a], n0);
                }
            }
            if (++s == STAGES) {
                s = 0;
                ph ^= 1;
            }
        }
    } else if (warp == 1) {
        // ---------------- MMA issuer (whole warp, uniform control flow) ----------------
        constexpr uint32_t idesc = make_idesc_bf16(128, BN, 1, 1);
        const bool issuer = elect_one();
        int s = 0;
        uint32_t ph = 0;
        for (int it = 0; it < num_k; ++it) {
            mbar_wait(&full_bar[s], ph);
            tc_fence_after();
            const uint32_t a_addr = smem_u32(smem + s * Cfg::kStageBytes);
            const uint32_t b_addr = a_addr + Cfg::kABytes;
            if (issuer) {
#pragma unroll
                for (int k = 0; k < kWgK / 16; ++k) {
                    // 16 pixel rows per MMA = 2 swizzle atoms of 8 rows (SBO); 64-channel column blocks are LBO apart.
                    const uint64_t da = make_smem_desc(a_addr + k * 2048, Cfg::kAtomBytes, 1024, kLayoutSW128);
                    const uint64_t db = make_smem_desc(b_addr + k * 2048, Cfg::kAtomBytes, 1024, kLayoutSW128);
                    umma_bf16_ss(tmem_base, da, db, idesc, (it | k) != 0);
                }
                if (do_bias && it % ny == by) {
                    constexpr uint32_t idesc_b = make_idesc_bf16(128, 16, 1, 1);
                    const uint64_t d1 = make_smem_desc(smem_u32(ones), Cfg::kAtomBytes, 1024, kLayoutSW128);
#pragma unroll
                    for (int k = 0; k < kWgK / 16; ++k) {
                        const uint64_t da = make_smem_desc(a_addr + k * 2048, Cfg::kAtomBytes, 1024, kLayoutSW128);
                        umma_bf16_ss(tmem_base + BN, da, d1, idesc_b, (it != by) || k != 0);
                    }
                }
                umma_commit(&empty_bar[s]);
            }
            if (++s == STAGES) {
                s = 0;
                ph ^= 1;
            }
        }
        if (issuer) umma_commit(accum_bar);
    } else {
        const int q = warp & 3;
        const int m = tile_m * 128 + q * 32 + lane;
        if (num_k > 0) {
            mbar_wait(accum_bar, 0);
            tc_fence_after();
        }
        if (do_bias) {
            uint32_t v[32];
            float b = 0.f;
            if (num_k > by) {  // this CTA issued at least one bias product (K step `by`)
                tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + BN, v);
                tmem_ld_wait();
                b = __uint_as_float(v[0]);
            }
            if (m < args.Mc) args.bias_partial[(static_cast<size_t>(split) * ny + by) * args.Mc + m] = b;
        }
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
            const int atom = c0 >> 6;
            const int ta = tap + atom / apt;
            float* dst = args.partial + ((static_cast<size_t>(split) * taps + ta) * args.Mc + m) * args.Nf;
            uint32_t v[32];
            if (num_k > 0) {
                tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + c0, v);
                tmem_ld_wait();
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = 0;
            }
            const int n_base = tile_n * BN + (atom % apt) * 64 + (c0 & 63);
            if (m < args.Mc && n_base < args.Nf) {
                float4* o = reinterpret_cast<float4*>(dst + n_base);
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    o[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                       __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, tmem_cols);
}


// ----------------------------------------------------------------------------
// CTA-pair weight gradient (Mc >= 256): one 256 x 256 UMMA (cta_group::2) per K step. Each CTA stages its own 128
// coarse channels and only HALF of the fine-tensor atoms, so a stage is 32 KB per SM per 512 MMA clocks (64 B/clk,
// the SM's ingest rate) instead of 48 KB: the one-CTA kernel above is ingest-bound at ~70 % of the tensor peak.
// ----------------------------------------------------------------------------
struct Wg2Cfg {
    static constexpr int kAtomBytes = kWgK * 128;
    static constexpr int kABytes = 2 * kAtomBytes;      // this CTA's 128 coarse channels
    static constexpr int kBBytes = 2 * kAtomBytes;      // this CTA's half (128 columns) of the 256-column B tile
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kStages = 6;
    static constexpr int kTmemCols = 256;
    static constexpr int kTmemColsBias = 512;
    static constexpr int kSmemBytes = kStages * kStageBytes + kOnesBytes + 1024 + 256;
};

__global__ void __launch_bounds__(kWgThreads) wgrad2_umma_kernel(const __grid_constant__ WgradArgs args) {
    using Cfg = Wg2Cfg;
    constexpr int STAGES = Cfg::kStages, BN = 256;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* ones = smem + STAGES * Cfg::kStageBytes;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(ones + kOnesBytes);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* accum_bar = empty_bar + STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

    const int warp = uniform_warp_idx();
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool do_bias = args.bias_partial != nullptr;
    const int ny = gridDim.y, by = blockIdx.y;
    const uint32_t tmem_cols = do_bias ? Cfg::kTmemColsBias : Cfg::kTmemCols;
    if (do_bias) {  // both CTAs of the pair: each provides its half of the (all-ones) B operand
        for (int i = threadIdx.x; i < kOnesBytes / 4; i += kWgThreads) reinterpret_cast<uint32_t*>(ones)[i] = 0x3F803F80u;
        fence_proxy_async_smem();
    }
    const int tile_m = blockIdx.x;  // = 2 * pair + rank
    const int n_tiles = (args.Nf + BN - 1) / BN;
    const int taps = 16;
    const int tpc = args.tpc;
    const int apt = (BN / 64) / tpc;
    const int tap = (blockIdx.y / n_tiles) * tpc;
    const int tile_n = blockIdx.y % n_tiles;
    const int split = blockIdx.z;
    const int kt_begin = static_cast<int>(static_cast<long>(args.k_tiles) * split / args.splits);
    const int kt_end = static_cast<int>(static_cast<long>(args.k_tiles) * (split + 1) / args.splits);
    const int num_k = kt_end - kt_begin;
    const int R = args.GH * args.GW;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&args.cmap);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(accum_bar, 1);
        mbar_fence_init();
    }
    if (warp == 1) tmem2_alloc(tmem_slot, tmem_cols);
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ---------------- TMA producer (both CTAs; completion counted on the leader's barriers) ----------------
        const bool issuer = elect_one();
        const CUtensorMap* a_fm[2];
        int a_dy[2], a_dx[2], a_ch[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int a = 2 * static_cast<int>(rank) + h;  // atom of the 256-column tile this CTA stages
            const int ta = tap + a / apt, cb = a % apt;
            const int ky = ta >> 2, kx = ta & 3;
            a_fm[h] = &args.fmap[((ky + 1) & 1) * 2 + ((kx + 1) & 1)];
            a_dy[h] = ((ky + 1) >> 1) - 1;
            a_dx[h] = ((kx + 1) >> 1) - 1;
            a_ch[h] = tile_n * BN + cb * 64;
        }
        const int lg_tpi = R >= kWgK ? 31 - __clz(R / kWgK) : 0, rows_per = kWgK / args.GW, ipk = R >= kWgK ? 1 : kWgK / R;
        int s = 0;
        uint32_t ph = 0;
        for (int it = 0; it < num_k; ++it) {
            mbar_wait(&empty_bar[s], ph ^ 1);
            if (issuer) {
                uint8_t* sa = smem + s * Cfg::kStageBytes;
                uint8_t* sb = sa + Cfg::kABytes;
                const uint32_t bar = mapa_rank(smem_u32(&full_bar[s]), 0);
                if (rank == 0) mbar_arrive_expect_tx(&full_bar[s], 2u * Cfg::kStageBytes);
                const int kt = kt_begin + it;
                int n0, y0;
                if (R >= kWgK) {
                    n0 = kt >> lg_tpi;
                    y0 = (kt & ((1 << lg_tpi) - 1)) * rows_per;
                } else {
                    n0 = kt * ipk;
                    y0 = 0;
                }
                tma2_load_2d(sa, &args.cmap, bar, tile_m * 128, kt * kWgK);
                tma2_load_2d(sa + Cfg::kAtomBytes, &args.cmap, bar, tile_m * 128 + 64, kt * kWgK);
#pragma unroll
                for (int h = 0; h < 2; ++h)
                    tma2_load_4d(sb + h * Cfg::kAtomBytes, a_fm[h], bar, a_ch[h], a_dx[h], y0 + a_dy[h], n0);
            }
            if (++s == STAGES) {
                s = 0;
                ph ^= 1;
            }
        }
    } else if (warp == 1) {
        if (rank == 0) {
            // ---------------- MMA issuer (leader CTA): 256 x 256 x 16 per instruction, both operands MN-major ----------------
            constexpr uint32_t idesc = make_idesc_bf16(256, BN, 1, 1);
            const bool issuer = elect_one();
            int s = 0;
            uint32_t ph = 0;
            for (int it = 0; it < num_k; ++it) {
                mbar_wait(&full_bar[s], ph);
                tc_fence_after();
                const uint32_t a_addr = smem_u32(smem + s * Cfg::kStageBytes);
                const uint32_t b_addr = a_addr + Cfg::kABytes;
                if (issuer) {
#pragma unroll
                    for (int k = 0; k < kWgK / 16; ++k) {
                        const uint64_t da = make_smem_desc(a_addr + k * 2048, Cfg::kAtomBytes, 1024, kLayoutSW128);
                        const uint64_t db = make_smem_desc(b_addr + k * 2048, Cfg::kAtomBytes, 1024, kLayoutSW128);
                        umma2_bf16_ss(tmem_base, da, db, idesc, (it | k) != 0);
                    }
                    if (do_bias && it % ny == by) {
                        constexpr uint32_t idesc_b = make_idesc_bf16(256, 16, 1, 1);
                        const uint64_t d1 = make_smem_desc(smem_u32(ones), Cfg::kAtomBytes, 1024, kLayoutSW128);
#pragma unroll
                        for (int k = 0; k < kWgK / 16; ++k) {
                            const uint64_t da = make_smem_desc(a_addr + k * 2048, Cfg::kAtomBytes, 1024, kLayoutSW128);
                            umma2_bf16_ss(tmem_base + BN, da, d1, idesc_b, (it != by) || k != 0);
                        }
                    }
                    umma2_commit(&empty_bar[s]);
                }
                if (++s == STAGES) {
                    s = 0;
                    ph ^= 1;
                }
            }
            if (issuer) umma2_commit(accum_bar);
        }
    } else {
        const int q = warp & 3;
        const int m = tile_m * 128 + q * 32 + lane;
        if (num_k > 0) {
            mbar_wait(accum_bar, 0);
            tc_fence_after();
        }
        if (do_bias) {
            uint32_t v[32];
            float b = 0.f;
            if (num_k > by) {
                tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + BN, v);
                tmem_ld_wait();
                b = __uint_as_float(v[0]);
            }
            if (m < args.Mc) args.bias_partial[(static_cast<size_t>(split) * ny + by) * args.Mc + m] = b;
        }
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
            const int atom = c0 >> 6;
            const int ta = tap + atom / apt;
            float* dst = args.partial + ((static_cast<size_t>(split) * taps + ta) * args.Mc + m) * args.Nf;
            uint32_t v[32];
            if (num_k > 0) {
                tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + c0, v);
                tmem_ld_wait();
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = 0;
            }
            const int n_base = tile_n * BN + (atom % apt) * 64 + (c0 & 63);
            if (m < args.Mc && n_base < args.Nf) {
                float4* o = reinterpret_cast<float4*>(dst + n_base);
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    o[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                       __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
            }
        }
    }
    tc_fence_before();
    cluster_sync_all();
    if (warp == 1) tmem2_dealloc(tmem_base, tmem_cols);
}

// partial [S][16][M][N] -> dW [M][N][16]
__global__ void wgrad_reduce_kernel(const float* __restrict__ partial, float* __restrict__ dW, int S, int M, int N,
                                    int accumulate, const float* __restrict__ bias_partial, float* __restrict__ dbias,
                                    int SB) {
    const long total = static_cast<long>(M) * N * 16;
    if (dbias && blockIdx.x == gridDim.x - 1) {  // fused bias gradient: [SB][M] partial sums
        for (int m = threadIdx.x; m < M; m += blockDim.x) {
            float acc = 0.f;
            for (int s = 0; s < SB; ++s) acc += bias_partial[static_cast<size_t>(s) * M + m];
            dbias[m] = acc;
        }
    }
    for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long>(gridDim.x) * blockDim.x) {
        // i enumerates (tap, m, n) so reads are coalesced; the 16x smaller output is written strided.
        const int n = static_cast<int>(i % N);
        const long t = i / N;
        const int m = static_cast<int>(t % M);
        const int tap = static_cast<int>(t / M);
        float acc = 0.f;
        for (int s = 0; s < S; ++s) acc += partial[(static_cast<size_t>(s) * 16 + tap) * M * N + static_cast<size_t>(m) * N + n];
        float* o = dW + (static_cast<size_t>(m) * N + n) * 16 + tap;
        *o = accumulate ? *o + acc : acc;
    }
}

// Split-K reductions on a side stream (wgrad_side_begin .. wgrad_side_end, set by the backward plans of sg_model.cu): the
// reduction of layer i's partials only feeds the gradient bucket, so it runs on `side` while the launch stream goes on
// with layer i's data gradient — a small grid that shares the SMs with the persistent tensor kernel instead of a ~25 us
// serial step (plus two kernel boundaries) per layer. The partial buffer is shared by all layers: the next weight-gradient
// kernel waits for the pending reduction (launch_wgrad), and wgrad_side_end joins the side stream back.
static thread_local struct {
    cudaStream_t side = nullptr;
    cudaEvent_t fork = nullptr, join = nullptr;
    bool pending = false;
    bool forked = false;  // wgrad_side_fork really opened the branch (wgrad_side_mark is a no-op otherwise)
} t_wside;
void wgrad_side_begin(cudaStream_t side, cudaEvent_t fork_ev, cudaEvent_t join_ev) {
    t_wside.side = side;
    t_wside.fork = fork_ev;
    t_wside.join = join_ev;
    t_wside.pending = false;
}
static void wgrad_side_wait(cudaStream_t stream) {  // `stream` may not touch the partial buffer / bucket before this
    if (t_wside.side && t_wside.pending) {
        cudaStreamWaitEvent(stream, t_wside.join, 0);
        t_wside.pending = false;
    }
}
// Other small kernels that only feed the gradient bucket may share the branch: fork returns the stream to launch them on
// (the side stream, ordered after everything `stream` holds so far; `stream` itself when no branch is open) and mark
// records them as pending.
cudaStream_t wgrad_side_fork(cudaStream_t stream) {
    static const bool extras = [] {  // SIGGAN_SIDE_EXTRA=0: only the split-K reductions use the branch (A/B comparison)
        const char* e = getenv("SIGGAN_SIDE_EXTRA");
        return !(e && e[0] == '0');
    }();
    t_wside.forked = false;
    if (!t_wside.side || !extras) return stream;
    cudaEventRecord(t_wside.fork, stream);
    cudaStreamWaitEvent(t_wside.side, t_wside.fork, 0);
    t_wside.forked = true;
    return t_wside.side;
}
void wgrad_side_mark() {
    if (!t_wside.side || !t_wside.forked) return;
    cudaEventRecord(t_wside.join, t_wside.side);
    t_wside.pending = true;
    t_wside.forked = false;
}
void wgrad_side_end(cudaStream_t stream) {
    wgrad_side_wait(stream);
    t_wside.side = nullptr;
}

void wgrad_reduce(const float* partial, float* dW, int S, int M, int N, int accumulate, cudaStream_t stream,
                  const float* bias_partial, float* dbias, int SB) {
    const long total = static_cast<long>(M) * N * 16;
    int blocks = static_cast<int>((total + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    note_launch();
    cudaStream_t on = stream;
    if (t_wside.side) {
        cudaEventRecord(t_wside.fork, stream);
        cudaStreamWaitEvent(t_wside.side, t_wside.fork, 0);
        on = t_wside.side;
    }
    wgrad_reduce_kernel<<<blocks, 256, 0, on>>>(partial, dW, S, M, N, accumulate, bias_partial, dbias, SB);
    if (t_wside.side) {
        cudaEventRecord(t_wside.join, t_wside.side);
        t_wside.pending = true;
    }
}

// sg_wgrad_thin.cu
bool wgrad_thin_supported(int cH, int cW, int Mc, int Nf);
int wgrad_thin_ctas(int nimg, int cH, int cW);
int launch_wgrad_thin(const __nv_bfloat16* coarse, const __nv_bfloat16* fine, int nimg, int cH, int cW, int Mc, int Nf,
                      float* partial, float* dW, int accumulate, cudaStream_t stream);

// sg_wgrad_pair.cu
bool wgrad_pair_supported(int cH, int cW, int Mc, int Nf);
int wgrad_pair_ctas(int nimg, int cH, int cW);
int launch_wgrad_pair(const __nv_bfloat16* coarse, const __nv_bfloat16* fine, int nimg, int cH, int cW, int Mc,
                      float* partial, float* dW, int accumulate, cudaStream_t stream);

static bool wgrad_pairs(int Mc, int Nf);
// Split-K factor: the launch is base * s CTAs (or CTA pairs) of one tile each on `slots` resident CTAs (pairs). Pick
// the s with the smallest modelled time = waves x K-steps per CTA x MMA time of a K step + the partial buffer's
// write + read (e.g. 32 pairs x 5 splits on 74 slots = 2.16 waves ran as 3; 23 splits fill the waves but move 193 MB).
static int wgrad_splits(int k_tiles, int m_tiles, int n_tiles, int taps = 16, bool pairs = false, int bn = 256,
                        double partial_bytes_per_split = 0.0) {
    const int base = (pairs ? m_tiles / 2 : m_tiles) * n_tiles * taps;
    const int slots = pairs ? 74 : 148;
    int max_s = k_tiles / 16 > 0 ? k_tiles / 16 : 1;
    if (max_s > 48) max_s = 48;
    const double t_k = 1.25 * 2.0 * bn / 1.9e9;  // seconds per 64-pixel K step (4 MMAs of bn/2 cycles, 80 % efficiency)
    int best = 1;
    double best_t = 1e30;
    for (int s = 1; s <= max_s; ++s) {
        const int waves = (base * s + slots - 1) / slots;
        const double t = waves * ((k_tiles + s - 1) / s) * t_k + 2.0 * s * partial_bytes_per_split / 5e12;
        if (t < best_t * 0.99) {
            best_t = t;
            best = s;
        }
    }
    return best;
}

static int wgrad_bn(int Nf) { return Nf >= 64 ? 256 : 64; }
// CTA pairs along M (cta_group::2, wgrad2_umma_kernel); SIGGAN_WGRAD2=0 keeps the one-CTA kernel (A/B comparison)
static bool wgrad_pairs(int Mc, int Nf) {
    static const bool pair_ok = [] {
        const char* e = getenv("SIGGAN_WGRAD2");
        return !(e && e[0] == '0');
    }();
    return pair_ok && wgrad_bn(Nf) == 256 && Mc % 256 == 0;
}
static int fc_wgrad_bn(int Kp) { return Kp >= 256 ? 256 : (Kp >= 128 ? 128 : 64); }
static int wgrad_tpc(int Nf) { return Nf == 64 ? 4 : (Nf == 128 ? 2 : 1); }  // filter taps stacked along N per CTA

size_t wgrad_partial_floats(int nimg, int cH, int cW, int Mc, int Nf) {
    if (wgrad_thin_supported(cH, cW, Mc, Nf)) return static_cast<size_t>(wgrad_thin_ctas(nimg, cH, cW)) * 16 * Mc * Nf;
    const int k_tiles = (nimg * cH * cW + kWgK - 1) / kWgK;
    const int BN = wgrad_bn(Nf);
    const int s = wgrad_splits(k_tiles, (Mc + 127) / 128, (Nf + BN - 1) / BN, 16 / wgrad_tpc(Nf), wgrad_pairs(Mc, Nf), BN,
                               64.0 * Mc * Nf);
    return static_cast<size_t>(s) * 16 * Mc * Nf + static_cast<size_t>(s) * 16 * Mc;  // + fused bias-gradient partials
}

template <int BN>
static int launch_wg(const WgradArgs& a, dim3 grid, cudaStream_t stream) {
    using Cfg = WgCfg<BN>;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(wgrad_umma_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             Cfg::kSmemBytes);
        if (e != cudaSuccess) SG_FAIL("cudaFuncSetAttribute(wgrad_umma<%d>): %s", BN, cudaGetErrorString(e));
        attr_set = true;
    }
    note_launch();
    wgrad_umma_kernel<BN><<<grid, kWgThreads, Cfg::kSmemBytes, stream>>>(a);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) SG_FAIL("wgrad_umma<%d> launch: %s", BN, cudaGetErrorString(e));
    return 0;
}

int launch_wgrad(const __nv_bfloat16* coarse, const __nv_bfloat16* fine, int nimg, int cH, int cW, int Mc, int Nf,
                 float* partial, size_t partial_floats, float* dW, int accumulate, cudaStream_t stream, float* dbias) {
    if (!is_pow2(cH) || !is_pow2(cW) || cW > kWgK) SG_FAIL("wgrad: coarse grid %dx%d unsupported", cH, cW);
    if (Mc % 8 != 0 || Nf % 8 != 0) SG_FAIL("wgrad: channels must be multiples of 8 (Mc=%d Nf=%d)", Mc, Nf);
    wgrad_side_wait(stream);  // the previous layer's reduction still reads the shared partial buffer
    if (wgrad_pair_supported(cH, cW, Mc, Nf)) {  // generator's last block: pixel-pair formulation on tcgen05
        if (dbias) SG_FAIL("wgrad: the pair kernel has no fused bias gradient");
        if (static_cast<size_t>(wgrad_pair_ctas(nimg, cH, cW)) * 16 * Mc * Nf > partial_floats)
            SG_FAIL("wgrad: partial workspace too small");
        if (launch_wgrad_pair(coarse, fine, nimg, cH, cW, Mc, partial, dW, accumulate, stream))
            SG_FAIL("wgrad_pair launch failed: %s (%s)", cudaGetErrorString(cudaGetLastError()), umma_last_error());
        return 0;
    }
    if (wgrad_thin_supported(cH, cW, Mc, Nf)) {  // thin layers: every pixel row staged once, mma.sync + ldmatrix
        if (wgrad_partial_floats(nimg, cH, cW, Mc, Nf) > partial_floats) SG_FAIL("wgrad: partial workspace too small");
        if (dbias) SG_FAIL("wgrad: the thin kernel has no fused bias gradient");
        if (launch_wgrad_thin(coarse, fine, nimg, cH, cW, Mc, Nf, partial, dW, accumulate, stream))
            SG_FAIL("wgrad_thin launch failed: %s", cudaGetErrorString(cudaGetLastError()));
        return 0;
    }
    WgradArgs a;
    memset(&a, 0, sizeof(a));
    a.GH = cH;
    a.GW = cW;
    a.nimg = nimg;
    a.Mc = Mc;
    a.Nf = Nf;
    const long pix = static_cast<long>(nimg) * cH * cW;
    a.k_tiles = static_cast<int>((pix + kWgK - 1) / kWgK);
    const int BN = wgrad_bn(Nf);
    const int m_tiles = (Mc + 127) / 128, n_tiles = (Nf + BN - 1) / BN;
    a.tpc = wgrad_tpc(Nf);
    a.splits = wgrad_splits(a.k_tiles, m_tiles, n_tiles, 16 / a.tpc, wgrad_pairs(Mc, Nf), BN, 64.0 * Mc * Nf);
    a.partial = partial;
    const size_t w_floats = static_cast<size_t>(a.splits) * 16 * Mc * Nf;
    if (w_floats + (dbias ? static_cast<size_t>(a.splits) * 16 * Mc : 0) > partial_floats)
        SG_FAIL("wgrad: partial workspace too small");
    a.bias_partial = dbias ? partial + w_floats : nullptr;
    if (make_map_2d(&a.cmap, coarse, Mc, pix, Mc, 64, kWgK)) return -1;
    const int R = cH * cW;
    const uint32_t bw = cW, bh = R >= kWgK ? kWgK / cW : cH, bn = R >= kWgK ? 1 : kWgK / R;
    for (int p = 0; p < 4; ++p)
        if (make_map_nhwc(&a.fmap[p], fine, nimg, 2 * cH, 2 * cW, Nf, 2, p >> 1, p & 1, 64, bw, bh, bn)) return -1;
    dim3 grid(m_tiles, (16 / a.tpc) * n_tiles, a.splits);
    int rc;
    if (wgrad_pairs(Mc, Nf)) {  // CTA pairs along M (cta_group::2)
        static bool attr_set = false;
        if (!attr_set) {
            cudaError_t e = cudaFuncSetAttribute(wgrad2_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                 Wg2Cfg::kSmemBytes);
            if (e != cudaSuccess) SG_FAIL("cudaFuncSetAttribute(wgrad2_umma): %s", cudaGetErrorString(e));
            attr_set = true;
        }
        note_launch();
        cudaLaunchConfig_t lc = {};
        lc.gridDim = grid;
        lc.blockDim = dim3(kWgThreads);
        lc.dynamicSmemBytes = Wg2Cfg::kSmemBytes;
        lc.stream = stream;
        cudaLaunchAttribute attr;
        attr.id = cudaLaunchAttributeClusterDimension;
        attr.val.clusterDim.x = 2;
        attr.val.clusterDim.y = 1;
        attr.val.clusterDim.z = 1;
        lc.attrs = &attr;
        lc.numAttrs = 1;
        cudaError_t e = cudaLaunchKernelEx(&lc, wgrad2_umma_kernel, a);
        if (e == cudaSuccess) e = cudaGetLastError();
        if (e != cudaSuccess) SG_FAIL("wgrad2_umma launch: %s", cudaGetErrorString(e));
        rc = 0;
    } else if (BN == 256)
        rc = launch_wg<256>(a, grid, stream);
    else if (BN == 128)
        rc = launch_wg<128>(a, grid, stream);
    else
        rc = launch_wg<64>(a, grid, stream);
    if (rc) return rc;
    wgrad_reduce(partial, dW, a.splits, Mc, Nf, accumulate, stream, a.bias_partial, dbias, a.splits * static_cast<int>(grid.y));
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) SG_FAIL("wgrad_reduce launch: %s", cudaGetErrorString(e));
    return 0;
}

// Generator fc weight gradient: dW[f(j)][i] = sum_b dy[b][j] * zp[b][i]  (j = NHWC column, f = NCHW feature).
__global__ void fc_wgrad_reduce_kernel(const float* __restrict__ partial, float* __restrict__ dW, int S, int F, int Kp,
                                       int C0, int latent) {
    const long total = static_cast<long>(F) * latent;
    for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long>(gridDim.x) * blockDim.x) {
        const int k = static_cast<int>(i % latent);
        const int j = static_cast<int>(i / latent);
        float acc = 0.f;
        for (int s = 0; s < S; ++s) acc += partial[(static_cast<size_t>(s) * F + j) * Kp + k];
        const int f = (j % C0) * 16 + j / C0;
        dW[static_cast<long>(f) * latent + k] = acc;
    }
}

size_t fc_wgrad_partial_floats(int B, int F, int Kp) {
    const int k_tiles = (B + kWgK - 1) / kWgK;
    const int BN = fc_wgrad_bn(Kp);
    const int s = wgrad_splits(k_tiles, (F + 127) / 128, (Kp + BN - 1) / BN, 1, false, BN, 4.0 * F * Kp);
    return static_cast<size_t>(s) * F * Kp;
}

int launch_fc_wgrad(const __nv_bfloat16* dy, const __nv_bfloat16* zp, int B, int C0, int Kp, int latent, float* partial,
                    size_t partial_floats, float* dW, cudaStream_t stream) {
    const int F = C0 * 16;
    wgrad_side_wait(stream);  // a pending split-K reduction of the layer above still reads the shared partial buffer
    WgradArgs a;
    memset(&a, 0, sizeof(a));
    a.plain = 1;
    a.tpc = 1;
    a.GH = 1;
    a.GW = kWgK;
    a.nimg = B;
    a.Mc = F;
    a.Nf = Kp;
    a.k_tiles = (B + kWgK - 1) / kWgK;
    const int BN = fc_wgrad_bn(Kp);
    const int m_tiles = (F + 127) / 128, n_tiles = (Kp + BN - 1) / BN;
    a.splits = wgrad_splits(a.k_tiles, m_tiles, n_tiles, 1, false, BN, 4.0 * F * Kp);
    a.partial = partial;
    if (static_cast<size_t>(a.splits) * F * Kp > partial_floats) SG_FAIL("fc_wgrad: partial workspace too small");
    if (make_map_2d(&a.cmap, dy, F, B, F, 64, kWgK)) return -1;
    if (make_map_2d(&a.fmap[0], zp, Kp, B, Kp, 64, kWgK)) return -1;
    dim3 grid(m_tiles, n_tiles, a.splits);
    int rc = BN == 256 ? launch_wg<256>(a, grid, stream) : (BN == 128 ? launch_wg<128>(a, grid, stream) : launch_wg<64>(a, grid, stream));
    if (rc) return rc;
    const long total = static_cast<long>(F) * latent;
    int blocks = static_cast<int>((total + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    note_launch();
    fc_wgrad_reduce_kernel<<<blocks, 256, 0, stream>>>(partial, dW, a.splits, F, Kp, C0, latent);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) SG_FAIL("fc_wgrad_reduce launch: %s", cudaGetErrorString(e));
    return 0;
}

}  // namespace sg
