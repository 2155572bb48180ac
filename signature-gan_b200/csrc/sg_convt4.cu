// sg_convt4.cu — phase-fused transposed convolution (ConvTranspose2d k4 s2 p1, gen…:46-54) for the Generator's thin,
// wide-grid blocks (Cout = 32, Cin = 32 or 64, input grid 16/32/64 wide), on tcgen05.
//
// The generic kernel (sg_conv_umma.cu) loads one 128-row A block per filter tap: 16 blocks per 128 input pixels,
// and with 64-byte rows the TMA unit, not the tensor core or HBM, is the limit (measured ~2.5 cycles per row).
// Here one schedule unit = ALL FOUR output parities of a 128-input-pixel tile:
//   * A: the tile's BH input rows plus one halo row above and below are loaded ONCE per horizontal shift
//     dx in {-1, 0, +1} (3 TMA boxes of (BH+2) x GW pixel rows; out-of-image rows/columns are zero-filled by TMA,
//     which is the padding). The operand of tap shift (dy, dx) is the dx buffer advanced by (1+dy) grid rows — a
//     whole number of swizzle atoms — so 3 loads feed all 16 (parity, tap) products: 4.6x fewer TMA rows.
//   * B: the 16 packed tap matrices [Cout x Cin] stay resident in shared memory for the whole launch.
//   * D: four accumulators [px][py][Cout] side by side in TMEM (128 columns), double-buffered, so the epilogue of a
//     tile overlaps the MMAs of the next; products that share a shifted input block are one MMA with the tap matrices
//     stacked along N: 11 x Cin/16 tcgen05.mma per tile instead of 16 x Cin/16, all issued by one thread.
//   * epilogue: warp (lane quarter q, py) owns 32 input pixels x {px = 0, 1} x 32 channels = for each input pixel
//     the 128 contiguous output bytes of pixels (2y+py, 2x), (2y+py, 2x+1); the warp's block is transposed through
//     swizzled shared memory and leaves as full 128-byte lines (a whole output row segment), optionally with the
//     eval-mode BatchNorm+ReLU folded in, or with the BatchNorm batch statistics accumulated per lane (training).
#include "sg_conv_umma.cuh"
#include "sg_kernels.cuh"
#include "sg_umma.cuh"

#include <cstdio>
#include <cstring>

namespace sg {

int sm_count_t4() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    }
    return n;
}

namespace {

constexpr int kT4Threads = 64 + 32 * 8;  // warp 0 producer, warp 1 MMA issuer, warps 2..9 epilogue
constexpr int kT4EpiBytes = 32 * 128;    // per epilogue warp: 32 rows x 128 bytes

struct ConvT4Args {
    CUtensorMap amap;  // input [C][W][H][N], box {BK, GW, BH+2, 1}
    CUtensorMap wmap;  // packed weights [Cout][16*Cin], box {BK, 32}
    int GH, GW, nimg, M_total;
    __nv_bfloat16* out;
    const float* scale;  // optional fused eval-mode BatchNorm (+ ReLU)
    const float* shift;
    float* stats_partial;  // optional [gridDim.x][2][32]
};

template <int BK>
struct T4Cfg {
    static constexpr int BN = 32;
    static constexpr int kRowBytes = BK * 2;
    static constexpr int kWTapBytes = BN * kRowBytes;
    static constexpr int kWBytes = 16 * kWTapBytes;
    static constexpr int kABufMax = (BK == 32 ? 4 * 64 : 10 * 16) * kRowBytes;  // (BH+2)*GW rows; see convt4_supported
    static constexpr int kStageBytes = 3 * kABufMax;
    static constexpr int kStages = BK == 32 ? 3 : 2;
    static constexpr int kTmemCols = 256;  // 2 x [px][py][32]
    static constexpr int kSmemBytes = kWBytes + kStages * kStageBytes + 8 * kT4EpiBytes + 1024 + 256;
    static constexpr uint32_t kLayout = (BK == 64) ? kLayoutSW128 : kLayoutSW64;
    static constexpr uint32_t kSBO = 8 * kRowBytes;
};

// Products of one tile grouped by input shift (dy, dx): parities that share a shifted input block are ONE MMA with
// their tap matrices stacked along N (the A block is read from shared memory once per MMA whatever N is, and with
// N = 32 those reads, not the math, bound the tensor pipe). Accumulator columns are ordered [px][py][32], so the
// dy = 0 shifts address 2 or 4 adjacent parities: 11 MMAs per k-step instead of 16.
struct T4Prod { int dy, dx, slot, n, col; };
__device__ constexpr T4Prod kT4Prods[11] = {
    {0, 0, 0, 128, 0},                                                           // all four parities
    {0, -1, 4, 64, 0},   {0, 1, 6, 64, 64},                                       // px = 0 / px = 1, both py
    {-1, 0, 8, 32, 0},   {-1, 0, 9, 32, 64},  {1, 0, 10, 32, 32}, {1, 0, 11, 32, 96},
    {-1, -1, 12, 32, 0}, {-1, 1, 13, 32, 64}, {1, -1, 14, 32, 32}, {1, 1, 15, 32, 96}};
// resident weight slot -> filter tap ky*4+kx (parity (py, px) with tap shift (dy, dx): ky = 1-py+2(py-dy), kx alike)
__device__ constexpr int kT4SlotTap[16] = {1 * 4 + 1, 2 * 4 + 1, 1 * 4 + 2, 2 * 4 + 2, 1 * 4 + 3, 2 * 4 + 3,
                                           1 * 4 + 0, 2 * 4 + 0, 3 * 4 + 1, 3 * 4 + 2, 0 * 4 + 1, 0 * 4 + 2,
                                           3 * 4 + 3, 3 * 4 + 0, 0 * 4 + 3, 0 * 4 + 0};

__device__ __forceinline__ float t4_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float t4_hi(uint32_t u) { return __uint_as_float(u & 0xFFFF0000u); }
__device__ __forceinline__ uint32_t t4_pack(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}
// 32 rows x 128 bytes; 16-byte piece k of row r is stored at piece k ^ (r & 7)
__device__ __forceinline__ uint32_t t4_off(int row, int k) { return row * 128 + ((k ^ (row & 7)) << 4); }

template <int BK, bool kStats>
__global__ void __launch_bounds__(kT4Threads, 1) convt4_kernel(const __grid_constant__ ConvT4Args args) {
    using Cfg = T4Cfg<BK>;
    constexpr int BN = Cfg::BN;
    constexpr int STAGES = Cfg::kStages;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* wsm = smem;
    uint8_t* ring = wsm + Cfg::kWBytes;
    uint8_t* epi_smem = ring + STAGES * Cfg::kStageBytes;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(epi_smem + 8 * kT4EpiBytes);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tfull_bar = empty_bar + STAGES;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint64_t* w_bar = tempty_bar + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_bar + 1);

    const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
    const int GW = args.GW, R = args.GH * GW;
    const int BH = 128 / GW;                          // input rows per tile
    const uint32_t abuf = static_cast<uint32_t>((BH + 2) * GW) * Cfg::kRowBytes;
    const int tpi = R / 128;                          // tiles per image
    const int total_tiles = args.M_total / 128;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&args.amap);
        tma_prefetch_desc(&args.wmap);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tfull_bar[a], 1);
            mbar_init(&tempty_bar[a], 8);
        }
        mbar_init(w_bar, 1);
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, Cfg::kTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ---------------- TMA producer: resident weights once, then 3 halo'd input blocks per tile ----------------
        // (whole warp in uniform control flow, one elected lane issues: see elect_one() in sg_umma.cuh)
        const bool issuer = elect_one();
        if (issuer) {
            mbar_arrive_expect_tx(w_bar, Cfg::kWBytes);
#pragma unroll
            for (int sl = 0; sl < 16; ++sl)
                tma_load_2d(wsm + sl * Cfg::kWTapBytes, &args.wmap, w_bar, kT4SlotTap[sl] * BK, 0);
        }
        const int lg_tpi = 31 - __clz(tpi);
        int s = 0;
        uint32_t ph = 0;
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
            const int n0 = t >> lg_tpi, y0 = (t & (tpi - 1)) * BH;
            mbar_wait(&empty_bar[s], ph ^ 1);
            if (issuer) {
                mbar_arrive_expect_tx(&full_bar[s], 3 * abuf);
                uint8_t* sa = ring + s * Cfg::kStageBytes;
#pragma unroll
                for (int d = 0; d < 3; ++d) tma_load_4d(sa + d * abuf, &args.amap, &full_bar[s], 0, d - 1, y0 - 1, n0);
            }
            if (++s == STAGES) {
                s = 0;
                ph ^= 1;
            }
        }
    } else if (warp == 1) {
        // ---------------- MMA issuer (whole warp, uniform control flow) ----------------
        const bool issuer = elect_one();
        mbar_wait(w_bar, 0);
        const uint32_t w_addr = smem_u32(wsm);
        const uint32_t row_step = static_cast<uint32_t>(GW) * Cfg::kRowBytes;  // one grid row of A
        uint32_t g = 0;
        int s = 0;
        uint32_t ph = 0;
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++g) {
            const int acc = g & 1;
            mbar_wait(&tempty_bar[acc], ((g >> 1) & 1) ^ 1);
            mbar_wait(&full_bar[s], ph);
            tc_fence_after();
            const uint32_t a_addr = smem_u32(ring + s * Cfg::kStageBytes);
            if (issuer) {
#pragma unroll
                for (int k = 0; k < BK / 16; ++k) {
#pragma unroll
                    for (int i = 0; i < 11; ++i) {
                        const uint32_t a0 = a_addr + (kT4Prods[i].dx + 1) * abuf + (1 + kT4Prods[i].dy) * row_step;
                        const uint32_t b0 = w_addr + kT4Prods[i].slot * Cfg::kWTapBytes;
                        const uint64_t da = make_smem_desc(a0 + k * 32, 0, Cfg::kSBO, Cfg::kLayout);
                        const uint64_t db = make_smem_desc(b0 + k * 32, 0, Cfg::kSBO, Cfg::kLayout);
                        umma_bf16_ss(tmem_base + acc * 128 + kT4Prods[i].col, da, db,
                                     make_idesc_bf16(128, kT4Prods[i].n, 0, 0), (i | k) != 0);
                    }
                }
                umma_commit(&empty_bar[s]);
                umma_commit(&tfull_bar[acc]);
            }
            if (++s == STAGES) {
                s = 0;
                ph ^= 1;
            }
        }
    } else {
        // ---------------- Epilogue ----------------
        const int q = warp & 3, py = (warp - 2) >> 2;
        uint8_t* stage = epi_smem + (warp - 2) * kT4EpiBytes;
        const uint32_t stage_u32 = smem_u32(stage);
        const int lgW = 31 - __clz(GW);
        const int wr_row = lane >> 3, wr_k = lane & 7;   // write-back role: 8 lanes per 128-byte row
        const int ch0 = (wr_k & 3) * 8;                  // channels of this lane's 16-byte piece
        float sc[8], sh[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            sc[j] = args.scale ? __ldg(args.scale + ch0 + j) : 1.f;
            sh[j] = args.scale ? __ldg(args.shift + ch0 + j) : 0.f;
        }
        float st_sum[8], st_sq[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) st_sum[j] = st_sq[j] = 0.f;
        uint32_t g = 0;
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++g) {
            const int acc = g & 1;
            const int n0 = t / tpi, y0 = (t - n0 * tpi) * BH;
            // output pixel (px = 0) of this lane's row in the tile
            const int rr = q * 32 + lane;
            const int yh = y0 + (rr >> lgW), xh = rr & (GW - 1);
            const int orow = (n0 * 2 * args.GH + 2 * yh + py) * 2 * GW + 2 * xh;
            mbar_wait(&tfull_bar[acc], (g >> 1) & 1);
            tc_fence_after();
#pragma unroll
            for (int px = 0; px < 2; ++px) {
                uint32_t v[32];
                tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * 128 + (px * 2 + py) * BN, v);
                tmem_ld_wait();
                if (px == 1) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tempty_bar[acc]);
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint4 pk = make_uint4(t4_pack(__uint_as_float(v[k * 8]), __uint_as_float(v[k * 8 + 1])),
                                                t4_pack(__uint_as_float(v[k * 8 + 2]), __uint_as_float(v[k * 8 + 3])),
                                                t4_pack(__uint_as_float(v[k * 8 + 4]), __uint_as_float(v[k * 8 + 5])),
                                                t4_pack(__uint_as_float(v[k * 8 + 6]), __uint_as_float(v[k * 8 + 7])));
                    sts128(stage_u32 + t4_off(lane, px * 4 + k), pk.x, pk.y, pk.z, pk.w);
                }
            }
            __syncwarp();
            // the BatchNorm/ReLU epilogue is applied in the write-back layout, where a lane's channels are fixed
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int r = i * 4 + wr_row;
                const int o = __shfl_sync(0xffffffffu, orow, r);
                const float4 df = lds128f(stage_u32 + t4_off(r, wr_k));
                uint4 d = make_uint4(__float_as_uint(df.x), __float_as_uint(df.y), __float_as_uint(df.z), __float_as_uint(df.w));
                uint32_t w4[4] = {d.x, d.y, d.z, d.w};
                if (args.scale) {
#pragma unroll
                    for (int tt = 0; tt < 4; ++tt) {
                        const float lo = fmaxf(fmaf(t4_lo(w4[tt]), sc[2 * tt], sh[2 * tt]), 0.f);
                        const float hi = fmaxf(fmaf(t4_hi(w4[tt]), sc[2 * tt + 1], sh[2 * tt + 1]), 0.f);
                        w4[tt] = t4_pack(lo, hi);
                    }
                    d = make_uint4(w4[0], w4[1], w4[2], w4[3]);
                }
                if (kStats) {
#pragma unroll
                    for (int tt = 0; tt < 4; ++tt) {
                        const float lo = t4_lo(w4[tt]), hi = t4_hi(w4[tt]);
                        st_sum[2 * tt] += lo;
                        st_sq[2 * tt] = fmaf(lo, lo, st_sq[2 * tt]);
                        st_sum[2 * tt + 1] += hi;
                        st_sq[2 * tt + 1] = fmaf(hi, hi, st_sq[2 * tt + 1]);
                    }
                }
                *reinterpret_cast<uint4*>(args.out + static_cast<size_t>(o) * BN + wr_k * 8) = d;
            }
            __syncwarp();
        }
        if (kStats) {
            // lanes with equal (wr_k & 3) hold partial sums of the same 8 channels
            float* slot = reinterpret_cast<float*>(stage);  // [2][32]
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float a = st_sum[j], b = st_sq[j];
                a += __shfl_xor_sync(0xffffffffu, a, 4);
                b += __shfl_xor_sync(0xffffffffu, b, 4);
#pragma unroll
                for (int o = 8; o < 32; o <<= 1) {
                    a += __shfl_xor_sync(0xffffffffu, a, o);
                    b += __shfl_xor_sync(0xffffffffu, b, o);
                }
                if (lane < 4) {
                    slot[lane * 8 + j] = a;
                    slot[32 + lane * 8 + j] = b;
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (kStats && warp >= 2 && warp < 4) {
        const int e = (warp - 2) * 32 + lane;  // [which][channel]
        float tot = 0.f;
#pragma unroll
        for (int wi = 0; wi < 8; ++wi) tot += reinterpret_cast<const float*>(epi_smem + wi * kT4EpiBytes)[e];
        args.stats_partial[static_cast<size_t>(blockIdx.x) * 64 + e] = tot;
    }
    if (warp == 1) tmem_dealloc(tmem_base, Cfg::kTmemCols);
}

template <int BK, bool kStats>
int launch_t4(const ConvT4Args& a, int grid, cudaStream_t stream) {
    using Cfg = T4Cfg<BK>;
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(convt4_kernel<BK, kStats>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 Cfg::kSmemBytes) != cudaSuccess)
            return -1;
        attr_set = true;
    }
    note_launch();
    convt4_kernel<BK, kStats><<<grid, kT4Threads, Cfg::kSmemBytes, stream>>>(a);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

}  // namespace

// Cout 32; Cin 32 with a 32- or 64-wide... the (BH+2)*GW row count of one A buffer must fit T4Cfg::kABufMax.
bool convt4_supported(int inH, int inW, int Cin, int Cout) {
    if (Cout != 32 || inH * inW < 128 || (inH * inW) % 128 != 0) return false;
    if (Cin == 32) return inW == 32 || inW == 64;   // 6 x 32 or 4 x 64 rows per buffer (128x128 images: last block)
    if (Cin == 64) return inW == 16;   // 10 x 16 rows per buffer
    return false;
}
int convt4_grid(int nimg, int inH, int inW) {
    const int tiles = nimg * inH * inW / 128;
    return tiles < sm_count_t4() ? tiles : sm_count_t4();
}

int launch_convt4(const __nv_bfloat16* in, const __nv_bfloat16* w_packed, int nimg, int inH, int inW, int Cin, int Cout,
                  __nv_bfloat16* out, const float* scale, const float* shift, float* stats_partial, cudaStream_t stream) {
    ConvT4Args a;
    memset(&a, 0, sizeof(a));
    a.GH = inH;
    a.GW = inW;
    a.nimg = nimg;
    a.M_total = nimg * inH * inW;
    a.out = out;
    a.scale = scale;
    a.shift = shift;
    a.stats_partial = stats_partial;
    const int BH = 128 / inW;
    if (make_map_nhwc(&a.amap, in, nimg, inH, inW, Cin, 1, 0, 0, Cin, inW, BH + 2, 1)) return -1;
    if (make_map_2d(&a.wmap, w_packed, 16ull * Cin, Cout, 16ull * Cin, Cin, 32)) return -1;
    const int grid = convt4_grid(nimg, inH, inW);
    if (Cin == 32) return stats_partial ? launch_t4<32, true>(a, grid, stream) : launch_t4<32, false>(a, grid, stream);
    return stats_partial ? launch_t4<64, true>(a, grid, stream) : launch_t4<64, false>(a, grid, stream);
}

}  // namespace sg
