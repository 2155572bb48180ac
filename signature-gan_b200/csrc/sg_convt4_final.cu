// sg_convt4_final.cu — the Generator's last upsample block and its tail in ONE kernel, for sampling and for the fakes of
// the D step (eval-mode BatchNorm, gen…:46-60 + :153-163):
//     ConvTranspose2d(32 -> 32, k4 s2 p1) -> BatchNorm2d(running stats) -> ReLU -> Conv3x3(32 -> 1) + bias -> tanh
// Unfused, the 32-channel full-resolution level (256 KB per 64x64 image in bf16) is written by the ConvT kernel and read
// back by the Conv3x3 kernel: 10 GB of the 13 GB a 16384-image sampling pass moves. Here it never leaves the SM: HBM sees
// the ConvT input (64 KB/image) and the image (16 KB fp32 and/or 4 KB uint8).
//
// Main loop = sg_convt4.cu (3 halo'd TMA boxes per 128-input-pixel tile, resident tap matrices, four parity
// accumulators in TMEM, double-buffered), except that a CTA walks whole images top to bottom (tiles of one image are
// consecutive), so that the 3x3 stencil's row halo is the previous tile's output, already in shared memory.
// Epilogue, per tile (16 warps; warp (q, py, px) owns one parity accumulator of 32 input pixels = every other pixel of
// half an output row, 32 channels each) — nothing but the tap products touches shared memory:
//   1. tcgen05.ld.16x256b returns the accumulator in the mma.sync C-fragment layout, where a lane's 8 channels are
//      fixed: running-stat BatchNorm (fp32) + ReLU (inside the bf16 conversion) turn it into the A fragment of
//   2. the tap GEMM P[pixel][tap] = sum_c a[pixel][c] * w3[c][tap] on mma.sync (N = 9 taps padded to 16): 8 MMAs per
//      warp instead of 288 FMAs per pixel;
//   3. P goes to a ring of output rows in shared memory (3 tiles deep: one named barrier per tile suffices);
//   4. after the barrier the 512 epilogue threads emit the rows whose lower neighbour is now known:
//      out = tanh(bias + sum of 9 shifted P entries), fp32 and/or the sampling egress uint8 (utils/inference.py:129).
// History (profiles/r01_ncu_tail_fused.txt): 8 epilogue warps with the activations staged in shared memory and a
// transposed BatchNorm pass: 0.49 ms per 4096 images, no faster than the two kernels it replaces; 16 warps, BatchNorm by
// broadcast loads: 0.39 ms, shared-memory bandwidth 85 % busy (tensor-core operand reads 41 %, LSU 44 %).
#include "sg_conv_umma.cuh"
#include "sg_kernels.cuh"
#include "sg_mma.cuh"
#include "sg_umma.cuh"

#include <cstdio>
#include <cstring>

namespace sg {

int sm_count_t4();  // sg_convt4.cu

namespace {

constexpr int kFThreads = 64 + 32 * 16;  // warp 0 producer, warp 1 MMA issuer, warps 2..17 epilogue
constexpr int kFC = 32;                 // channels in and out of the block (gen…:139,149)
constexpr int kFRowBytes = kFC * 2;
constexpr int kFWTapBytes = kFC * kFRowBytes;
constexpr int kFWBytes = 16 * kFWTapBytes;
constexpr int kFABufMax = 4 * 64 * kFRowBytes;  // (BH+2)*GW input pixel rows: 6 x 32 or 4 x 64
constexpr int kFStageBytes = 3 * kFABufMax;
constexpr int kFStages = 2;
constexpr int kFTmemCols = 256;                 // 2 x [px][py][32]
constexpr int kFPBytes = 24 * 9 * 84 * 4;       // P ring: 3 tiles x (8 rows x 9 x 84 | 4 rows x 9 x 148) floats
constexpr int kFSmemBytes = kFWBytes + kFStages * kFStageBytes + kFPBytes + 1024 + 256;
constexpr uint32_t kFSBO = 8 * kFRowBytes;

// Products of one tile, grouped by input shift (dy, dx) so that parities sharing a shifted input block are ONE MMA
// with their tap matrices stacked along N (the A block is read from shared memory once per MMA whatever N is, and with
// N = 32 those reads, not the math, bound the tensor pipe). Accumulator columns are ordered [px][py][32]: the dy = 0
// shifts then address 2 or 4 adjacent parities. 11 MMAs per k-step instead of 16.
struct T4Prod { int dy, dx, slot, n, col; };
__device__ constexpr T4Prod kProds[11] = {
    {0, 0, 0, 128, 0},                                                           // all four parities
    {0, -1, 4, 64, 0},   {0, 1, 6, 64, 64},                                       // px = 0 / px = 1, both py
    {-1, 0, 8, 32, 0},   {-1, 0, 9, 32, 64},  {1, 0, 10, 32, 32}, {1, 0, 11, 32, 96},
    {-1, -1, 12, 32, 0}, {-1, 1, 13, 32, 64}, {1, -1, 14, 32, 32}, {1, 1, 15, 32, 96}};
// resident weight slot -> filter tap ky*4+kx (parity (py, px) with tap shift (dy, dx) uses ky = 1-py+2(py-dy),
// kx likewise)
__device__ constexpr int kSlotTap[16] = {1 * 4 + 1, 2 * 4 + 1, 1 * 4 + 2, 2 * 4 + 2, 1 * 4 + 3, 2 * 4 + 3, 1 * 4 + 0, 2 * 4 + 0,
                                         3 * 4 + 1, 3 * 4 + 2, 0 * 4 + 1, 0 * 4 + 2, 3 * 4 + 3, 3 * 4 + 0, 0 * 4 + 3, 0 * 4 + 0};

struct ConvT4FinalArgs {
    CUtensorMap amap;  // input [C][W][H][N], box {32, GW, BH+2, 1}
    CUtensorMap wmap;  // packed ConvT weights [Cout][16*Cin], box {32, 32}
    int GH, GW, nimg;
    const float* scale;  // eval-mode BatchNorm of the block, folded: a = relu(y * scale + shift)
    const float* shift;
    const float* w3;     // Conv3x3 weight (1, 32, 3, 3) fp32
    const float* b3;     // its bias
    float* out;          // (nimg, 1, 2GH, 2GW) fp32, may be null
    uint8_t* out_u8;     // same shape, may be null
};

__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 512;" ::: "memory"); }
__device__ __forceinline__ void sts32f(uint32_t addr, float v) {
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ float lds32f(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
    return v;
}
// 16 TMEM lanes x 32 columns in the mma.sync accumulator layout (probed: tools/probes/tmem_ld_layout.cu):
// r[4n + j] = lane (t >> 2) + 8 * (j >> 1), column 8n + 2 * (t & 3) + (j & 1)
__device__ __forceinline__ void tmem_ld_16x256b_x4(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.16x256b.x4.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
// {relu(lo), relu(hi)} as packed bf16
__device__ __forceinline__ uint32_t pack2_relu(float lo, float hi) {
    uint32_t d;
    asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
    return d;
}

__global__ void __launch_bounds__(kFThreads, 1) convt4_final_kernel(const __grid_constant__ ConvT4FinalArgs args) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* wsm = smem;
    uint8_t* ring = wsm + kFWBytes;
    float* P = reinterpret_cast<float*>(ring + kFStages * kFStageBytes);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(P) + kFPBytes);
    uint64_t* empty_bar = full_bar + kFStages;
    uint64_t* tfull_bar = empty_bar + kFStages;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint64_t* w_bar = tempty_bar + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_bar + 1);

    const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
    const int GW = args.GW, R = args.GH * GW;
    const int BH = 128 / GW;                          // input rows per tile
    const uint32_t abuf = static_cast<uint32_t>((BH + 2) * GW) * kFRowBytes;
    const int tpi = R / 128;                          // tiles per image
    const int lg_tpi = 31 - __clz(tpi);
    const int my_imgs = (args.nimg - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) /
                        static_cast<int>(gridDim.x);
    const int my_tiles = my_imgs << lg_tpi;           // tile j of this CTA: image blockIdx.x + (j >> lg_tpi) * gridDim.x

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&args.amap);
        tma_prefetch_desc(&args.wmap);
        for (int s = 0; s < kFStages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tfull_bar[a], 1);
            mbar_init(&tempty_bar[a], 16);
        }
        mbar_init(w_bar, 1);
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, kFTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ---------------- TMA producer (whole warp in uniform control flow, one elected lane issues) ----------------
        const bool issuer = elect_one();
        if (issuer) {
            mbar_arrive_expect_tx(w_bar, kFWBytes);
#pragma unroll
            for (int sl = 0; sl < 16; ++sl)
                tma_load_2d(wsm + sl * kFWTapBytes, &args.wmap, w_bar, kSlotTap[sl] * kFC, 0);
        }
        int s = 0;
        uint32_t ph = 0;
        for (int j = 0; j < my_tiles; ++j) {
            const int n0 = blockIdx.x + (j >> lg_tpi) * gridDim.x, y0 = (j & (tpi - 1)) * BH;
            mbar_wait(&empty_bar[s], ph ^ 1);
            if (issuer) {
                mbar_arrive_expect_tx(&full_bar[s], 3 * abuf);
                uint8_t* sa = ring + s * kFStageBytes;
#pragma unroll
                for (int d = 0; d < 3; ++d) tma_load_4d(sa + d * abuf, &args.amap, &full_bar[s], 0, d - 1, y0 - 1, n0);
            }
            if (++s == kFStages) {
                s = 0;
                ph ^= 1;
            }
        }
    } else if (warp == 1) {
        // ---------------- MMA issuer (whole warp, uniform control flow) ----------------
        const bool issuer = elect_one();
        mbar_wait(w_bar, 0);
        const uint32_t w_addr = smem_u32(wsm);
        const uint32_t row_step = static_cast<uint32_t>(GW) * kFRowBytes;  // one grid row of A
        int s = 0;
        uint32_t ph = 0;
        for (int g = 0; g < my_tiles; ++g) {
            const int acc = g & 1;
            mbar_wait(&tempty_bar[acc], ((g >> 1) & 1) ^ 1);
            mbar_wait(&full_bar[s], ph);
            tc_fence_after();
            const uint32_t a_addr = smem_u32(ring + s * kFStageBytes);
            if (issuer) {
#pragma unroll
                for (int k = 0; k < kFC / 16; ++k) {
#pragma unroll
                    for (int i = 0; i < 11; ++i) {
                        const uint32_t a0 = a_addr + (kProds[i].dx + 1) * abuf + (1 + kProds[i].dy) * row_step;
                        const uint32_t b0 = w_addr + kProds[i].slot * kFWTapBytes;
                        const uint64_t da = make_smem_desc(a0 + k * 32, 0, kFSBO, kLayoutSW64);
                        const uint64_t db = make_smem_desc(b0 + k * 32, 0, kFSBO, kLayoutSW64);
                        umma_bf16_ss(tmem_base + acc * 128 + kProds[i].col, da, db,
                                     make_idesc_bf16(128, kProds[i].n, 0, 0), (i | k) != 0);
                    }
                }
                umma_commit(&empty_bar[s]);
                umma_commit(&tfull_bar[acc]);
            }
            if (++s == kFStages) {
                s = 0;
                ph ^= 1;
            }
        }
    } else {
        // ---------------- Epilogue: 16 warps, warp (q, py, px) owns one parity accumulator of 32 input pixels -------
        const int q = warp & 3, py = (warp - 2) >> 3, px = ((warp - 2) >> 2) & 1;
        const int te = (warp - 2) * 32 + lane;           // 0..511
        const int gid = lane >> 2, t4 = lane & 3;
        const uint32_t P_u32 = smem_u32(P);
        const int lgW = 31 - __clz(GW);
        const int OW = 2 * GW, OH = 2 * args.GH, lgOW = lgW + 1;
        // P row layout: even output columns at [0, GW), odd ones at [HP, HP + GW), HP == 16 (mod 32) and
        // pitch == 4 (mod 16): the accumulator stores (lanes = consecutive INPUT pixels of one parity) and the emission
        // loads (lanes = consecutive output pixels) are both bank-conflict-free
        const int HP = GW + 16, pitch = HP + GW + 4;
        const int RPT = 2 * BH;                          // output rows per tile
        // the warp's 32 input pixels inside the tile: output row and first input column
        const int row_in_tile = 2 * ((q * 32) >> lgW) + py;
        const int xin0 = (q * 32) & (GW - 1);
        // BatchNorm coefficients of this lane's 8 channels (accumulator fragment: channels 8n + 2*t4 + {0, 1})
        float sc[4][2], sh[4][2];
#pragma unroll
        for (int n = 0; n < 4; ++n)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                sc[n][e] = __ldg(args.scale + 8 * n + 2 * t4 + e);
                sh[n][e] = __ldg(args.shift + 8 * n + 2 * t4 + e);
            }
        // B fragments of the tap GEMM: n-block nb holds taps nb*8 + gid; k-step ks covers channels ks*16 + {2t4, 2t4+1,
        // 8+2t4, 8+2t4+1}
        uint32_t bw[2][2][2];
#pragma unroll
        for (int nb = 0; nb < 2; ++nb) {
            const int tap = nb * 8 + gid;
#pragma unroll
            for (int ks = 0; ks < 2; ++ks)
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    const int ch = ks * 16 + r * 8 + 2 * t4;
                    bw[nb][ks][r] =
                        tap < 9 ? pack2_bf16(__ldg(args.w3 + ch * 9 + tap), __ldg(args.w3 + (ch + 1) * 9 + tap)) : 0u;
                }
        }
        const float b3 = __ldg(args.b3);
        // emission role (fixed per thread: 512 = RPT * OW): output column e_x of tile-relative output row e_ro
        const int e_x = te & (OW - 1), e_ro = (te >> lgOW) - 1;
        const uint32_t row_bytes = 9 * pitch * 4;
        // byte offsets of (x-1, tap 0), (x, tap 1), (x+1, tap 2) inside a tap-row triple (parity-split columns)
        const uint32_t xo_l = (((e_x - 1) & 1) * HP + ((e_x - 1) >> 1)) * 4;
        const uint32_t xo_m = (pitch + (e_x & 1) * HP + (e_x >> 1)) * 4;
        const uint32_t xo_r = (2 * pitch + ((e_x + 1) & 1) * HP + ((e_x + 1) >> 1)) * 4;
        const bool has_l = e_x > 0, has_r = e_x < OW - 1;
        // source rows e_ro-1, e_ro, e_ro+1 with filter rows ky = 0, 1, 2; negative rows live in the previous tile's slots
        uint32_t e_off[3];
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
            const int rr = e_ro + ky - 1;
            e_off[ky] = static_cast<uint32_t>((rr < 0 ? RPT + rr : rr) * 9 + ky * 3) * pitch * 4;
        }
        auto tap_row = [&](uint32_t base) -> float {   // sum over kx of one source row's taps at columns x-1, x, x+1
            const float l = has_l ? lds32f(base + xo_l) : 0.f;
            const float r = has_r ? lds32f(base + xo_r) : 0.f;
            return l + lds32f(base + xo_m) + r;
        };
        auto put = [&](float a, int n, int yo) {
            // tanh(a) = 1 - 2 / (exp(2a) + 1): exact limits at +-inf, absolute error ~1e-7 (the output is an image)
            const float v = 1.f - __fdividef(2.f, __expf(2.f * a) + 1.f);
            const size_t o = (static_cast<size_t>(n) * OH + yo) * OW + e_x;
            if (args.out) args.out[o] = v;
            if (args.out_u8) {
                float qv = (v + 1.f) * 127.5f;
                qv = fminf(fmaxf(qv, 0.f), 255.f);
                args.out_u8[o] = static_cast<uint8_t>(qv);  // numpy astype(uint8) truncates (utils/inference.py:129)
            }
        };

        for (int g = 0; g < my_tiles; ++g) {
            const int acc = g & 1;
            const int n0 = blockIdx.x + (g >> lg_tpi) * gridDim.x, ti = g & (tpi - 1);
            const int slot0 = (g % 3) * RPT;                       // ring rows of this tile
            const int slotp = ((g + 2) % 3) * RPT;                 // ring rows of the previous tile
            mbar_wait(&tfull_bar[acc], (g >> 1) & 1);
            tc_fence_after();
            // accumulator -> registers in the mma.sync C-fragment layout (16 pixels x 32 channels per load): after
            // BatchNorm + ReLU and bf16 packing that IS the A fragment of the tap GEMM — no shared-memory staging
            uint32_t v[2][16];
            const uint32_t tsrc = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * 128 + (px * 2 + py) * kFC;
            tmem_ld_16x256b_x4(tsrc, v[0]);
            tmem_ld_16x256b_x4(tsrc + (16u << 16), v[1]);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[acc]);
            const uint32_t prow = P_u32 + ((slot0 + row_in_tile) * 9 * pitch + px * HP + xin0 + gid) * 4;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                uint32_t a[2][4];
#pragma unroll
                for (int n = 0; n < 4; ++n) {
                    const float lo0 = fmaf(__uint_as_float(v[h][4 * n]), sc[n][0], sh[n][0]);
                    const float hi0 = fmaf(__uint_as_float(v[h][4 * n + 1]), sc[n][1], sh[n][1]);
                    const float lo1 = fmaf(__uint_as_float(v[h][4 * n + 2]), sc[n][0], sh[n][0]);
                    const float hi1 = fmaf(__uint_as_float(v[h][4 * n + 3]), sc[n][1], sh[n][1]);
                    a[n >> 1][(n & 1) * 2] = pack2_relu(lo0, hi0);       // pixel gid
                    a[n >> 1][(n & 1) * 2 + 1] = pack2_relu(lo1, hi1);   // pixel gid + 8
                }
                float d0[4] = {0.f, 0.f, 0.f, 0.f}, d1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int ks = 0; ks < 2; ++ks) {
                    mma_bf16(d0, a[ks], bw[0][ks][0], bw[0][ks][1]);
                    mma_bf16(d1, a[ks], bw[1][ks][0], bw[1][ks][1]);
                }
                const uint32_t pr = prow + h * 64;
                sts32f(pr + (2 * t4) * pitch * 4, d0[0]);
                sts32f(pr + (2 * t4 + 1) * pitch * 4, d0[1]);
                sts32f(pr + (2 * t4) * pitch * 4 + 32, d0[2]);
                sts32f(pr + (2 * t4 + 1) * pitch * 4 + 32, d0[3]);
                if (t4 == 0) {
                    sts32f(pr + 8 * pitch * 4, d1[0]);
                    sts32f(pr + 8 * pitch * 4 + 32, d1[2]);
                }
            }
            epi_bar();
            // rows ti*RPT-1 .. ti*RPT+RPT-2 of the image are complete now: one pixel per thread (+ the last row at the
            // last tile)
            const uint32_t cur = P_u32 + slot0 * row_bytes, prev = P_u32 + slotp * row_bytes;
            if (ti > 0 || e_ro >= 0) {
                float a = b3;
                if (ti > 0 || e_ro > 0) a += tap_row((e_ro >= 1 ? cur : prev) + e_off[0]);
                a += tap_row((e_ro >= 0 ? cur : prev) + e_off[1]);
                a += tap_row(cur + e_off[2]);
                put(a, n0, ti * RPT + e_ro);
            }
            if (ti == tpi - 1 && te < OW) {
                const uint32_t last = cur + (RPT - 2) * row_bytes;
                put(b3 + tap_row(last) + tap_row(last + row_bytes + 3 * pitch * 4), n0, OH - 1);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, kFTmemCols);
}

}  // namespace

// The last block of either generator: 32 -> 32 channels on a 32- or 64-wide input grid (64x64 / 128x128 images).
bool convt4_final_supported(int inH, int inW, int Cin, int Cout) {
    return Cin == 32 && Cout == 32 && inH == inW && (inW == 32 || inW == 64);
}

int launch_convt4_final(const __nv_bfloat16* in, const __nv_bfloat16* w_packed, int nimg, int inH, int inW,
                        const float* scale, const float* shift, const float* w3, const float* b3, float* out,
                        uint8_t* out_u8, cudaStream_t stream) {
    if (!convt4_final_supported(inH, inW, 32, 32) || !scale || !shift || (!out && !out_u8)) return -1;
    ConvT4FinalArgs a;
    memset(&a, 0, sizeof(a));
    a.GH = inH;
    a.GW = inW;
    a.nimg = nimg;
    a.scale = scale;
    a.shift = shift;
    a.w3 = w3;
    a.b3 = b3;
    a.out = out;
    a.out_u8 = out_u8;
    const int BH = 128 / inW;
    if (make_map_nhwc(&a.amap, in, nimg, inH, inW, kFC, 1, 0, 0, kFC, inW, BH + 2, 1)) return -1;
    if (make_map_2d(&a.wmap, w_packed, 16ull * kFC, kFC, 16ull * kFC, kFC, 32)) return -1;
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(convt4_final_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFSmemBytes) !=
            cudaSuccess)
            return -1;
        attr_set = true;
    }
    const int grid = nimg < sm_count_t4() ? nimg : sm_count_t4();
    note_launch();
    convt4_final_kernel<<<grid, kFThreads, kFSmemBytes, stream>>>(a);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

}  // namespace sg
