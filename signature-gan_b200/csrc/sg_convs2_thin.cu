// sg_convs2_thin.cu — 4x4 stride-2 pad-1 convolution of a THIN fine-grid tensor (32 channels in, 32 out, coarse grid 32
// wide) on tcgen05: the data gradient of the Generator's last ConvTranspose2d block (gen…:46-54, 32 -> 32 channels at
// 64 x 64), the largest activation of the training step (1 GB at B = 4096).
//
// The generic kernel (sg_conv_umma.cu, kConvS2) gathers one 128-row A block per filter tap: 16 TMA boxes of 64-byte rows
// per 128 output pixels, a 4x re-read of every input element through the TMA unit (~2.5 clk per 64-byte row), which
// bound the layer at 3.9 TB/s of effective traffic. Here the contraction is re-associated so that every input row is
// staged TWICE (once per vertical tap pair) instead of four times, as 128-byte rows, and the horizontal taps cost
// nothing on the load side:
//   * a GEMM row is a horizontal PIXEL PAIR j of the fine grid, (2j, 2j+1), 64 contiguous values [px][c] = one 128-byte
//     swizzle row. For a vertical tap ky the A tile is the box {64, 32 pairs, 4 rows} of the row-parity plane that holds
//     fine row 2*iy - 1 + ky (out-of-image rows zero-filled by TMA = the vertical padding).
//   * pair j feeds THREE output pixels: ix = j (taps kx = 1 from px 0 and kx = 2 from px 1), ix = j - 1 (kx = 3 from
//     px 0) and ix = j + 1 (kx = 0 from px 1). So B stacks three 32-column targets, [t0 | t- | t+] (K rows of the
//     unused pixel half are zero), N = 96, K = 4 ky x 64: 16 tcgen05.mma (128 x 96 x 16) per 128 output pixels.
//   * the epilogue is a col2im: a warp owns one output row (32 pixels = 32 TMEM lanes), so the horizontal shifts are
//     two warp shuffles per column, out[ix] = P[ix][t0] + P[ix+1][t-] + P[ix-1][t+]; lanes 31 / 0 take zero = the
//     horizontal padding. Then the ReLU gate of the layer below and one 64-byte store per pixel.
#include "sg_conv_umma.cuh"
#include "sg_kernels.cuh"
#include "sg_umma.cuh"

#include <cstdio>
#include <cstring>

namespace sg {

int make_map_tiled(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box, int swizzle_bytes);

namespace {

constexpr int kS2Threads = 64 + 32 * 8;  // warp 0 TMA producer, warp 1 MMA issuer, warps 2..9 epilogue
constexpr int kS2BoxBytes = 128 * 128;   // 4 rows x 32 pairs x 128 bytes
constexpr int kS2StageBytes = 4 * kS2BoxBytes;
constexpr int kS2Stages = 2;
constexpr int kS2WBytes = 4 * 96 * 128;  // [ky][96 rows][64] bf16
constexpr int kS2TmemCols = 256;         // 2 accumulators of 96 columns, 128 apart
constexpr int kS2StatBytes = 8 * 32 * 4;  // per epilogue warp: (sum d, sum d*y) of its 16 channels
constexpr int kS2SmemBytes = kS2WBytes + kS2Stages * kS2StageBytes + 1024 + 256 + kS2StatBytes;

struct ConvS2ThinArgs {
    CUtensorMap xmap[2];  // fine tensor, row-parity planes: [64 (px,c)][32 pairs][GH rows][N]
    CUtensorMap wmap;     // stacked weights [4*96][64]
    int GH, nimg, total_tiles;
    __nv_bfloat16* out;          // [N][GH][32][32]
    const __nv_bfloat16* gate;   // saved activation of the layer below at the output position, or null
    float slope;
    // BatchNorm-gate form (training backward): `gate` is the layer below's PRE-BatchNorm output y; the activation
    // derivative is taken on y * gate_scale[c] + gate_shift[c] (the expression its forward applied), and the CTA also
    // writes stats_partial[blockIdx.x][2][32] = (sum d, sum d * y) over its stored outputs d — the raw reductions of that
    // layer's BatchNorm backward (bn_bwd_finalize), which then needs no pass of its own over d and y.
    const float* gate_scale;
    const float* gate_shift;
    float* stats_partial;
};

// Wt[ky][t*32 + n][px*32 + c] from the data-gradient pack w[n][ky*4+kx][c] (n = output channel of this convolution)
__global__ void convs2_thin_pack_kernel(const __nv_bfloat16* __restrict__ w, __nv_bfloat16* __restrict__ wt) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 4 * 96 * 64) return;
    const int k = i & 63, row = (i >> 6) % 96, ky = i / (96 * 64);
    const int t = row >> 5, n = row & 31, px = k >> 5, c = k & 31;
    // target 0 (ix = j): px0 -> kx 1, px1 -> kx 2; target 1 (ix = j - 1): px0 -> kx 3; target 2 (ix = j + 1): px1 -> kx 0
    int kx = -1;
    if (t == 0) kx = px == 0 ? 1 : 2;
    else if (t == 1) kx = px == 0 ? 3 : -1;
    else kx = px == 1 ? 0 : -1;
    wt[i] = kx < 0 ? __float2bfloat16(0.f) : w[(n * 16 + ky * 4 + kx) * 32 + c];
}

__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ uint32_t s2_pack(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}

__global__ void __launch_bounds__(kS2Threads, 1) convs2_thin_kernel(const __grid_constant__ ConvS2ThinArgs args) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* wsm = smem;
    uint8_t* ring = wsm + kS2WBytes;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(ring + kS2Stages * kS2StageBytes);
    uint64_t* empty_bar = full_bar + kS2Stages;
    uint64_t* tfull_bar = empty_bar + kS2Stages;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint64_t* w_bar = tempty_bar + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_bar + 1);
    float* stat_slots = reinterpret_cast<float*>(ring + kS2Stages * kS2StageBytes + 256);

    const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
    const int tpi = args.GH / 4;  // tiles per image
    const int total_tiles = args.total_tiles;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&args.xmap[0]);
        tma_prefetch_desc(&args.xmap[1]);
        tma_prefetch_desc(&args.wmap);
        for (int s = 0; s < kS2Stages; ++s) {
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
    if (warp == 1) tmem_alloc(tmem_slot, kS2TmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ---------------- TMA producer (whole warp, uniform control flow; one elected lane issues) ----------------
        const bool issuer = elect_one();
        if (issuer) {
            mbar_arrive_expect_tx(w_bar, kS2WBytes);
#pragma unroll
            for (int ky = 0; ky < 4; ++ky) tma_load_2d(wsm + ky * 96 * 128, &args.wmap, w_bar, 0, ky * 96);
        }
        int s = 0;
        uint32_t ph = 0;
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
            const int n0 = t / tpi, y0 = (t - n0 * tpi) * 4;
            mbar_wait(&empty_bar[s], ph ^ 1);
            if (issuer) {
                mbar_arrive_expect_tx(&full_bar[s], kS2StageBytes);
                uint8_t* sa = ring + s * kS2StageBytes;
                // fine row 2 iy - 1 + ky: ky 0 -> odd plane, row iy - 1; 1 -> even, iy; 2 -> odd, iy; 3 -> even, iy + 1
                tma_load_4d(sa + 0 * kS2BoxBytes, &args.xmap[1], &full_bar[s], 0, 0, y0 - 1, n0);
                tma_load_4d(sa + 1 * kS2BoxBytes, &args.xmap[0], &full_bar[s], 0, 0, y0, n0);
                tma_load_4d(sa + 2 * kS2BoxBytes, &args.xmap[1], &full_bar[s], 0, 0, y0, n0);
                tma_load_4d(sa + 3 * kS2BoxBytes, &args.xmap[0], &full_bar[s], 0, 0, y0 + 1, n0);
            }
            if (++s == kS2Stages) {
                s = 0;
                ph ^= 1;
            }
        }
    } else if (warp == 1) {
        // ---------------- MMA issuer ----------------
        constexpr uint32_t idesc = make_idesc_bf16(128, 96, 0, 0);
        const bool issuer = elect_one();
        mbar_wait(w_bar, 0);
        const uint32_t w_addr = smem_u32(wsm);
        uint32_t g = 0;
        int s = 0;
        uint32_t ph = 0;
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++g) {
            const int acc = g & 1;
            mbar_wait(&tempty_bar[acc], ((g >> 1) & 1) ^ 1);
            mbar_wait(&full_bar[s], ph);
            tc_fence_after();
            const uint32_t a_addr = smem_u32(ring + s * kS2StageBytes);
            if (issuer) {
#pragma unroll
                for (int ky = 0; ky < 4; ++ky) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint64_t da = make_smem_desc(a_addr + ky * kS2BoxBytes + k * 32, 0, 1024, kLayoutSW128);
                        const uint64_t db = make_smem_desc(w_addr + ky * 96 * 128 + k * 32, 0, 1024, kLayoutSW128);
                        umma_bf16_ss(tmem_base + acc * 128, da, db, idesc, (ky | k) != 0);
                    }
                }
                umma_commit(&empty_bar[s]);
                umma_commit(&tfull_bar[acc]);
            }
            if (++s == kS2Stages) {
                s = 0;
                ph ^= 1;
            }
        }
    } else {
        // ---------------- Epilogue: col2im by warp shuffles, gate, store ----------------
        const int q = warp & 3;              // TMEM lane quarter = output row of the tile
        const int half = (warp - 2) >> 2;    // which 16 of the 32 output channels
        const int c0 = half * 16;
        const bool ygate = args.gate && args.gate_scale;
        const bool stats = ygate && args.stats_partial;
        float gsc[16], gsh[16], s0[16], s1[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            gsc[j] = ygate ? __ldg(args.gate_scale + c0 + j) : 1.f;
            gsh[j] = ygate ? __ldg(args.gate_shift + c0 + j) : 0.f;
            s0[j] = s1[j] = 0.f;
        }
        uint32_t g = 0;
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++g) {
            const int acc = g & 1;
            const int n0 = t / tpi, y0 = (t - n0 * tpi) * 4;
            const size_t opix = (static_cast<size_t>(n0) * args.GH + y0 + q) * 32 + lane;
            uint4 gq[2] = {make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0)};
            if (args.gate) {
                const uint4* gp = reinterpret_cast<const uint4*>(args.gate + opix * 32 + c0);
                gq[0] = __ldg(gp);
                gq[1] = __ldg(gp + 1);
            }
            mbar_wait(&tfull_bar[acc], (g >> 1) & 1);
            tc_fence_after();
            const uint32_t tbase = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * 128 + c0;
            uint32_t v0[16], vm[16], vp[16];
            tmem_ld_32x16(tbase, v0);
            tmem_ld_32x16(tbase + 32, vm);
            tmem_ld_32x16(tbase + 64, vp);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[acc]);
            float f[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                float a = __shfl_down_sync(0xffffffffu, __uint_as_float(vm[j]), 1);  // P[ix + 1][t-]
                float b = __shfl_up_sync(0xffffffffu, __uint_as_float(vp[j]), 1);    // P[ix - 1][t+]
                if (lane == 31) a = 0.f;
                if (lane == 0) b = 0.f;
                f[j] = __uint_as_float(v0[j]) + a + b;
            }
            if (args.gate) {
                const uint32_t w8[8] = {gq[0].x, gq[0].y, gq[0].z, gq[0].w, gq[1].x, gq[1].y, gq[1].z, gq[1].w};
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    // plain gate: sign of the saved activation; BatchNorm gate: sign of y * scale + shift (ygate)
                    f[2 * j] *= fmaf(__uint_as_float(w8[j] << 16), gsc[2 * j], gsh[2 * j]) > 0.f ? 1.f : args.slope;
                    f[2 * j + 1] *=
                        fmaf(__uint_as_float(w8[j] & 0xFFFF0000u), gsc[2 * j + 1], gsh[2 * j + 1]) > 0.f ? 1.f : args.slope;
                }
            }
            uint32_t pk[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) pk[j] = s2_pack(f[2 * j], f[2 * j + 1]);
            uint4* op = reinterpret_cast<uint4*>(args.out + opix * 32 + c0);
            op[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            op[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
            if (stats) {  // of the STORED (bf16-rounded) values, as a separate reduction pass over `out` would see them
                const uint32_t w8[8] = {gq[0].x, gq[0].y, gq[0].z, gq[0].w, gq[1].x, gq[1].y, gq[1].z, gq[1].w};
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float dlo = __uint_as_float(pk[j] << 16), dhi = __uint_as_float(pk[j] & 0xFFFF0000u);
                    s0[2 * j] += dlo;
                    s0[2 * j + 1] += dhi;
                    s1[2 * j] = fmaf(dlo, __uint_as_float(w8[j] << 16), s1[2 * j]);
                    s1[2 * j + 1] = fmaf(dhi, __uint_as_float(w8[j] & 0xFFFF0000u), s1[2 * j + 1]);
                }
            }
        }
        if (stats) {  // warp totals (fixed shuffle tree: deterministic) -> this warp's slot [2][16]
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                float a = s0[j], b = s1[j];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    a += __shfl_xor_sync(0xffffffffu, a, o);
                    b += __shfl_xor_sync(0xffffffffu, b, o);
                }
                if (lane == 0) {
                    stat_slots[(warp - 2) * 32 + j] = a;
                    stat_slots[(warp - 2) * 32 + 16 + j] = b;
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (args.gate && args.gate_scale && args.stats_partial && threadIdx.x < 64) {
        // fold the four row-quarter warps of each channel half in a fixed order; one partial row per CTA
        const int which = threadIdx.x >> 5, ch = threadIdx.x & 31, hf = ch >> 4;
        float tot = 0.f;
#pragma unroll
        for (int qq = 0; qq < 4; ++qq) tot += stat_slots[(hf * 4 + qq) * 32 + which * 16 + (ch & 15)];
        args.stats_partial[(static_cast<size_t>(blockIdx.x) * 2 + which) * 32 + ch] = tot;
    }
    if (warp == 1) tmem_dealloc(tmem_base, kS2TmemCols);
}

int sm_count_s2() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    }
    return n;
}

}  // namespace

// fine tensor [nimg][inH][inW][32] with inW = 64 (coarse grid 32 wide), 32 output channels, inH / 2 a multiple of 4
bool convs2_thin_supported(int inH, int inW, int Cin, int Cout) {
    static const bool on = [] {
        const char* e = getenv("SIGGAN_CONVS2_THIN");
        return !(e && e[0] == '0');
    }();
    return on && Cin == 32 && Cout == 32 && inW == 64 && inH >= 8 && (inH / 2) % 4 == 0;
}
size_t convs2_thin_scratch_bytes() { return static_cast<size_t>(kS2WBytes); }

// w_packed: the data-gradient pack [Cout][16][Cin] (K contiguous). scratch: convs2_thin_scratch_bytes() of device
// memory for the stacked weights (rebuilt by every call: 24 K elements).
int convs2_thin_ctas(int nimg, int inH) {
    const int tiles = nimg * (inH / 2 / 4);
    return tiles < sm_count_s2() ? tiles : sm_count_s2();
}

int launch_convs2_thin(const __nv_bfloat16* in, const __nv_bfloat16* w_packed, int nimg, int inH, int inW,
                       __nv_bfloat16* out, const __nv_bfloat16* gate, float slope, void* scratch, cudaStream_t stream,
                       const float* gate_scale, const float* gate_shift, float* stats_partial) {
    ConvS2ThinArgs a;
    memset(&a, 0, sizeof(a));
    a.GH = inH / 2;
    a.nimg = nimg;
    a.total_tiles = nimg * (a.GH / 4);
    a.out = out;
    a.gate = gate;
    a.slope = slope;
    a.gate_scale = gate_scale;
    a.gate_shift = gate_shift;
    a.stats_partial = stats_partial;
    __nv_bfloat16* wt = static_cast<__nv_bfloat16*>(scratch);
    note_launch();
    convs2_thin_pack_kernel<<<(4 * 96 * 64 + 255) / 256, 256, 0, stream>>>(w_packed, wt);
    const uint64_t row_bytes = static_cast<uint64_t>(inW) * 32 * 2;  // one fine row
    for (int py = 0; py < 2; ++py) {
        const uint64_t dims[4] = {64, static_cast<uint64_t>(inW / 2), static_cast<uint64_t>(a.GH), static_cast<uint64_t>(nimg)};
        const uint64_t strides[3] = {128, 2 * row_bytes, static_cast<uint64_t>(inH) * row_bytes};
        const uint32_t box[4] = {64, 32, 4, 1};
        if (make_map_tiled(&a.xmap[py], reinterpret_cast<const char*>(in) + py * row_bytes, 4, dims, strides, box, 128))
            return -1;
    }
    if (make_map_2d(&a.wmap, wt, 64, 4 * 96, 64, 64, 96)) return -1;
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(convs2_thin_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kS2SmemBytes) !=
            cudaSuccess)
            return -1;
        attr_set = true;
    }
    const int grid = a.total_tiles < sm_count_s2() ? a.total_tiles : sm_count_s2();
    note_launch();
    convs2_thin_kernel<<<grid, kS2Threads, kS2SmemBytes, stream>>>(a);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

}  // namespace sg
