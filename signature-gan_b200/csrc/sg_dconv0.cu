// sg_dconv0.cu — the Discriminator's first block, Conv 4x4 s2 p1 from the 1-channel image to 64 channels
// (disc…:134-139), forward / weight gradient / image gradient, bf16 tensor-core mode.
//
// K = 16 taps x 1 channel: one k-step of a 16x8x16 MMA. All three passes are HBM-bound (16 KB of image against
// 128 KB of bf16 activations per 64x64 image; 2.1 MFLOP), but on the CUDA cores their 1 MFMA per image costs more
// issue slots than the memory system needs time, so the contraction runs on warp-level mma.sync with fragments built
// straight from global memory — the im2col gather of a 1-channel image is two floats per fragment register:
//   forward : A = im2col(x) [16 pixels x 16 taps] from scalar loads, B = w [16 taps x 8 channels] in registers;
//             bias + LeakyReLU + Dropout2d mask in the accumulator registers; tile transposed through swizzled
//             shared memory and written as one contiguous 2 KB run (16 pixels x 64 channels, NHWC).
//   wgrad   : dW[c][tap] = sum_pix dy[pix][c] * im2col(x)[pix][tap]. A = dy^T: the pixel-pair packing the fragment
//             layout wants is done with PRMT on 16-byte loads (rows of the MMA are a permutation of the channels,
//             undone when the partials are written); an extra all-ones B block yields the bias gradient.
//   dgrad   : T[pix][tap] = sum_c dy[pix][c] * w[c][tap] per band of output rows into shared memory (k permuted so
//             that A fragments are plain 16-byte loads), then dx = col2im(T): 4 reads per image pixel.
#include "sg_elem.cuh"
#include "sg_kernels.cuh"
#include "sg_mma.cuh"

namespace sg {
namespace {

constexpr int kC0 = 64;  // output channels one launch handles (the first block's width, disc…:134)
// LD = channels of the NHWC activation tensor the 64 channels live in: 64 for the reference's widths; the 2x-width variant
// (BASELINE configs[4]) runs two launches over the halves of a 128-channel tensor (pointers pre-offset by the caller).

template <int LD>
__global__ void __launch_bounds__(256)
dconv0_fwd_mma_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                      const float* __restrict__ mask, float slope, bf16* __restrict__ a, int B, int S) {
    __shared__ __align__(16) uint8_t stage_all[8][2048];  // per warp: 16 pixels x 128 bytes, chunk-swizzled
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, gid = lane >> 2, t4 = lane & 3;
    uint8_t* stage = stage_all[warp];
    const int O = S / 2, lgO = ilog2(O), tpr = O / 16;
    uint32_t bw[8][2];
    float bs[8][2];
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) {
        const float* wp = w + (nb * 8 + gid) * 16 + 2 * t4;
        bw[nb][0] = pack2_bf16(wp[0], wp[1]);
        bw[nb][1] = pack2_bf16(wp[8], wp[9]);
        bs[nb][0] = bias[nb * 8 + 2 * t4];
        bs[nb][1] = bias[nb * 8 + 2 * t4 + 1];
    }
    const int ky0 = t4 >> 1, kxo = 2 * (t4 & 1);
    // A warp walks a strip of 16 output columns down ONE image (O tiles). The kernel is bound by load issue / latency
    // (ncu: half of all stall samples wait on the image / mask loads at their first use), and in this order
    //   * the dropout keep-scales of the image are loaded once per strip instead of once per tile,
    //   * a tile's lower two input rows (h = 1: iy = 2 oy + 1 + ky0) are the next tile's upper two (h = 0), so only four
    //     image values per lane and tile are loaded instead of eight — issued one tile ahead of their use.
    constexpr int kSeg = 8;  // output rows per unit: short enough to balance the warps, long enough to amortise the set-up
    const int segs = O / kSeg, lg_tpr = ilog2(tpr), lg_segs = ilog2(segs);
    const int units = B * segs * tpr;  // (image, row segment, strip)
    const int wstride = static_cast<int>(gridDim.x) * 8;
    for (int u = static_cast<int>(blockIdx.x) * 8 + warp; u < units; u += wstride) {
        const int strip = u & (tpr - 1);
        const int oy_lo = ((u >> lg_tpr) & (segs - 1)) * kSeg;
        const long n = u >> (lg_tpr + lg_segs);
        const int ox0 = strip * 16;
        const float* xi = x + n * S * S;
        float2 mk[8];
#pragma unroll
        for (int nb = 0; nb < 8; ++nb)
            mk[nb] = mask ? __ldg(reinterpret_cast<const float2*>(mask + n * LD + nb * 8 + 2 * t4)) : make_float2(1.f, 1.f);
        int ixs[2];
        bool okl[2], okr[2];
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
            ixs[rr] = 2 * (ox0 + gid + 8 * rr) - 1 + kxo;
            okl[rr] = ixs[rr] >= 0;
            okr[rr] = ixs[rr] + 1 < S;
        }
        auto load_rows = [&](int iy, float (&v)[4]) {  // v[rr*2 + e] = x[iy][ixs[rr] + e] (0 outside the image)
            const bool yok = iy >= 0 && iy < S;
            const float* row = xi + iy * S;
#pragma unroll
            for (int rr = 0; rr < 2; ++rr) {
                v[rr * 2] = (yok && okl[rr]) ? __ldg(row + ixs[rr]) : 0.f;
                v[rr * 2 + 1] = (yok && okr[rr]) ? __ldg(row + ixs[rr] + 1) : 0.f;
            }
        };
        float up[4], lo[4], lo_next[4];
        load_rows(2 * oy_lo - 1 + ky0, up);      // first tile: rows 2 oy - 1 + ky0 and 2 oy + 1 + ky0
        load_rows(2 * oy_lo + 1 + ky0, lo);
        bf16* dst = a + ((n * O + oy_lo) * O + ox0) * LD;
        (void)lgO;
        for (int oy = oy_lo; oy < oy_lo + kSeg; ++oy, dst += static_cast<long>(O) * LD) {
            if (oy + 1 < oy_lo + kSeg) load_rows(2 * oy + 3 + ky0, lo_next);
            uint32_t af[4];
            af[0] = pack2_bf16(up[0], up[1]);
            af[1] = pack2_bf16(up[2], up[3]);
            af[2] = pack2_bf16(lo[0], lo[1]);
            af[3] = pack2_bf16(lo[2], lo[3]);
            float acc[8][4];
#pragma unroll
            for (int nb = 0; nb < 8; ++nb) {
                acc[nb][0] = acc[nb][1] = acc[nb][2] = acc[nb][3] = 0.f;
                mma_bf16(acc[nb], af, bw[nb][0], bw[nb][1]);
            }
#pragma unroll
            for (int nb = 0; nb < 8; ++nb) {
                const float m0 = mk[nb].x, m1 = mk[nb].y;
                float v[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float t = acc[nb][e] + bs[nb][e & 1];
                    v[e] = (t > 0.f ? t : t * slope) * ((e & 1) ? m1 : m0);
                }
                const uint32_t off = static_cast<uint32_t>((nb ^ gid) << 4) + t4 * 4;   // (gid + 8) & 7 == gid
                *reinterpret_cast<uint32_t*>(stage + gid * 128 + off) = pack2_bf16(v[0], v[1]);
                *reinterpret_cast<uint32_t*>(stage + (gid + 8) * 128 + off) = pack2_bf16(v[2], v[3]);
            }
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int row = j * 4 + (lane >> 3), chunk = lane & 7;
                const uint4 d = *reinterpret_cast<const uint4*>(stage + row * 128 + ((chunk ^ (row & 7)) << 4));
                *reinterpret_cast<uint4*>(dst + row * LD + chunk * 8) = d;
            }
            __syncwarp();
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                up[k] = lo[k];
                lo[k] = lo_next[k];
            }
        }
    }
}

// partial[block][64*16 + 64]: dW in (c, tap) order followed by dbias.
template <int LD>
__global__ void __launch_bounds__(256)
dconv0_wgrad_mma_kernel(const float* __restrict__ x, const bf16* __restrict__ dy, float* __restrict__ partial, int B,
                        int S) {
    constexpr int NOUT = kC0 * 17;
    __shared__ float red[8][NOUT];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, gid = lane >> 2, t4 = lane & 3;
    const int O = S / 2, lgO = ilog2(O), tpr = O / 16;
    float acc[4][3][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int nb = 0; nb < 3; ++nb) acc[i][nb][0] = acc[i][nb][1] = acc[i][nb][2] = acc[i][nb][3] = 0.f;
    const uint32_t ones = gid == 0 ? 0x3F803F80u : 0u;  // B column 0 of the third block = 1.0: row sums = dbias
    const int kx = gid & 3, kyb = gid >> 2;
    // 32-bit shift/mask index math (S and therefore tpr are powers of two): the 64-bit runtime division this replaces
    // cost more issue slots per 16-pixel tile than its 8 MMAs
    const int lg_tpr = ilog2(tpr);
    const int total = B * O * tpr;
    const int stride = static_cast<int>(gridDim.x) * 8;
    // latency-bound like the forward kernel (ncu: 64 % of the stall samples wait on the first use of a load): the
    // gradient and image loads of the NEXT tile are issued before the current tile's MMAs
    struct TileIn {
        uint4 L[4];   // dy: 8-channel chunk `gid` of pixels 2*t4, 2*t4+1, 2*t4+8, 2*t4+9 of the tile
        float v[8];   // im2col(x): taps gid (block 0) and 8 + gid (block 1) at the same four pixels
    };
    auto load_tile = [&](int tile, TileIn& in) {
        const int ox0 = (tile & (tpr - 1)) * 16;
        const int r = tile >> lg_tpr;
        const int oy = r & (O - 1);
        const long n = r >> lgO;
        const bf16* dp = dy + ((n * O + oy) * O + ox0) * LD + gid * 8;
        in.L[0] = __ldg(reinterpret_cast<const uint4*>(dp + (2 * t4) * LD));
        in.L[1] = __ldg(reinterpret_cast<const uint4*>(dp + (2 * t4 + 1) * LD));
        in.L[2] = __ldg(reinterpret_cast<const uint4*>(dp + (2 * t4 + 8) * LD));
        in.L[3] = __ldg(reinterpret_cast<const uint4*>(dp + (2 * t4 + 9) * LD));
        const float* xi = x + n * S * S;
#pragma unroll
        for (int nb = 0; nb < 2; ++nb) {
            const int iy = 2 * oy - 1 + nb * 2 + kyb;
            const bool yok = iy >= 0 && iy < S;
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                const int px = 2 * t4 + (p & 1) + 8 * (p >> 1);
                const int ix = 2 * (ox0 + px) - 1 + kx;
                in.v[nb * 4 + p] = (yok && ix >= 0 && ix < S) ? __ldg(xi + iy * S + ix) : 0.f;
            }
        }
    };
    TileIn nxt;
    int tile = static_cast<int>(blockIdx.x) * 8 + warp;
    if (tile < total) load_tile(tile, nxt);
    for (; tile < total; tile += stride) {
        const TileIn cur = nxt;
        if (tile + stride < total) load_tile(tile + stride, nxt);
        const uint4 L0 = cur.L[0], L1 = cur.L[1], L2 = cur.L[2], L3 = cur.L[3];
        uint32_t b[2][2];
#pragma unroll
        for (int nb = 0; nb < 2; ++nb) {
            b[nb][0] = pack2_bf16(cur.v[nb * 4], cur.v[nb * 4 + 1]);
            b[nb][1] = pack2_bf16(cur.v[nb * 4 + 2], cur.v[nb * 4 + 3]);
        }
        const uint32_t l0[4] = {L0.x, L0.y, L0.z, L0.w}, l1[4] = {L1.x, L1.y, L1.z, L1.w};
        const uint32_t l2[4] = {L2.x, L2.y, L2.z, L2.w}, l3[4] = {L3.x, L3.y, L3.z, L3.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            // MMA row gid <-> channel 8*gid + 2*i (low halves), row gid + 8 <-> channel 8*gid + 2*i + 1 (high halves)
            uint32_t af[4];
            af[0] = __byte_perm(l0[i], l1[i], 0x5410);
            af[1] = __byte_perm(l0[i], l1[i], 0x7632);
            af[2] = __byte_perm(l2[i], l3[i], 0x5410);
            af[3] = __byte_perm(l2[i], l3[i], 0x7632);
            mma_bf16(acc[i][0], af, b[0][0], b[0][1]);
            mma_bf16(acc[i][1], af, b[1][0], b[1][1]);
            mma_bf16(acc[i][2], af, ones, ones);
        }
    }
    float* rw = red[warp];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int c = 8 * gid + 2 * i;
#pragma unroll
        for (int nb = 0; nb < 2; ++nb) {
            const int tap = nb * 8 + 2 * t4;
            rw[c * 16 + tap] = acc[i][nb][0];
            rw[c * 16 + tap + 1] = acc[i][nb][1];
            rw[(c + 1) * 16 + tap] = acc[i][nb][2];
            rw[(c + 1) * 16 + tap + 1] = acc[i][nb][3];
        }
        if (t4 == 0) {
            rw[kC0 * 16 + c] = acc[i][2][0];
            rw[kC0 * 16 + c + 1] = acc[i][2][2];
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < NOUT; i += blockDim.x) {
        float s = 0.f;
#pragma unroll
        for (int wv = 0; wv < 8; ++wv) s += red[wv][i];
        partial[static_cast<long>(blockIdx.x) * NOUT + i] = s;
    }
}

constexpr int kTPitch = 17;  // floats per pixel of the tap-sum tile (16 taps + 1: conflict-free col2im reads)

template <int LD>
__global__ void __launch_bounds__(256, 3)
dconv0_dgrad_mma_kernel(const bf16* __restrict__ dy, const float* __restrict__ w, float* __restrict__ dx, int B, int S,
                        int accumulate) {
    extern __shared__ float T[];  // [(RB + 2)][O][kTPitch]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, gid = lane >> 2, t4 = lane & 3;
    const int O = S / 2, tpr = O / 16, lgS = ilog2(S);
    const int RB = 1024 / O;  // output rows per band (the whole 32x32 grid at 64x64)
    const int bands = O / RB, units = B * bands;
    // B fragments: k index 2*t4+e of step s <-> channel 8*t4 + 2*s + e; k index 8+2*t4+e <-> channel 32 + 8*t4 + 2*s + e
    uint32_t bw[4][2][2];
#pragma unroll
    for (int s = 0; s < 4; ++s)
#pragma unroll
        for (int nb = 0; nb < 2; ++nb) {
            const int c = 8 * t4 + 2 * s, tap = nb * 8 + gid;
            bw[s][nb][0] = pack2_bf16(w[c * 16 + tap], w[(c + 1) * 16 + tap]);
            bw[s][nb][1] = pack2_bf16(w[(32 + c) * 16 + tap], w[(33 + c) * 16 + tap]);
        }
    for (int u = blockIdx.x; u < units; u += gridDim.x) {
        const long n = u / bands;
        const int r0 = (u - static_cast<int>(n) * bands) * RB;
        // ---- phase 1: T rows r0-1 .. r0+RB
        const int ntile = (RB + 2) * tpr;
        // two tiles per iteration: the 8 independent 16-byte loads of both are in flight before the first MMA
        for (int tl = warp; tl < ntile; tl += 16) {
            uint4 A[2][4];
            int rrs[2], oxs[2];
            bool live[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int t = tl + 8 * h;
                rrs[h] = t / tpr;
                oxs[h] = (t - rrs[h] * tpr) * 16;
                const int oy = r0 - 1 + rrs[h];
                live[h] = t < ntile && oy >= 0 && oy < O;
                if (live[h]) {
                    const bf16* dp = dy + ((n * O + oy) * O + oxs[h] + gid) * LD + t4 * 8;
                    A[h][0] = __ldg(reinterpret_cast<const uint4*>(dp));
                    A[h][1] = __ldg(reinterpret_cast<const uint4*>(dp + 32));
                    A[h][2] = __ldg(reinterpret_cast<const uint4*>(dp + 8 * LD));
                    A[h][3] = __ldg(reinterpret_cast<const uint4*>(dp + 8 * LD + 32));
                }
            }
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                if (tl + 8 * h >= ntile) break;
                float acc[2][4];
#pragma unroll
                for (int nb = 0; nb < 2; ++nb) acc[nb][0] = acc[nb][1] = acc[nb][2] = acc[nb][3] = 0.f;
                if (live[h]) {
                    const uint32_t lo0[4] = {A[h][0].x, A[h][0].y, A[h][0].z, A[h][0].w};
                    const uint32_t hi0[4] = {A[h][1].x, A[h][1].y, A[h][1].z, A[h][1].w};
                    const uint32_t lo8[4] = {A[h][2].x, A[h][2].y, A[h][2].z, A[h][2].w};
                    const uint32_t hi8[4] = {A[h][3].x, A[h][3].y, A[h][3].z, A[h][3].w};
#pragma unroll
                    for (int s = 0; s < 4; ++s) {
                        const uint32_t af[4] = {lo0[s], lo8[s], hi0[s], hi8[s]};
                        mma_bf16(acc[0], af, bw[s][0][0], bw[s][0][1]);
                        mma_bf16(acc[1], af, bw[s][1][0], bw[s][1][1]);
                    }
                }
                float* t0 = T + (rrs[h] * O + oxs[h] + gid) * kTPitch + 2 * t4;
#pragma unroll
                for (int nb = 0; nb < 2; ++nb) {
                    t0[nb * 8] = acc[nb][0];
                    t0[nb * 8 + 1] = acc[nb][1];
                    t0[8 * kTPitch + nb * 8] = acc[nb][2];
                    t0[8 * kTPitch + nb * 8 + 1] = acc[nb][3];
                }
            }
        }
        __syncthreads();
        // ---- phase 2: dx rows 2*r0 .. 2*(r0+RB)-1; dx[iy][ix] = sum of the 2x2 (ky, kx) taps that reach it
        const int npix = 2 * RB * S;
        for (int i = threadIdx.x; i < npix; i += blockDim.x) {
            const int ix = i & (S - 1), iyl = i >> lgS, iy = 2 * r0 + iyl;
            const int kya = (iy + 1) & 1, kxa = (ix + 1) & 1;
            float s = 0.f;
#pragma unroll
            for (int ty = 0; ty < 2; ++ty) {
                const int ky = kya + 2 * ty;
                const int trow = ((iy + 1 - ky) >> 1) - (r0 - 1);   // 0 .. RB+1
#pragma unroll
                for (int tx = 0; tx < 2; ++tx) {
                    const int kx = kxa + 2 * tx;
                    const int ox = (ix + 1 - kx) >> 1;
                    if (ox >= 0 && ox < O) s += T[(trow * O + ox) * kTPitch + ky * 4 + kx];
                }
            }
            float* o = dx + (n * S + iy) * S + ix;
            *o = accumulate ? *o + s : s;   // second channel half of a wide first block adds to the first
        }
        __syncthreads();
    }
}

}  // namespace

void dconv0_fwd_mma(const float* x, const float* w, const float* bias, const float* mask, float slope, bf16* a, int B,
                    int S, cudaStream_t s, int ld) {
    const long units = static_cast<long>(B) * (S / 2 / 8) * (S / 32);  // (image, 8-row segment, 16-column strip) per warp
    static int per_sm = 0;
    if (per_sm == 0 &&
        (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, dconv0_fwd_mma_kernel<64>, 256, 0) != cudaSuccess || per_sm < 1))
        per_sm = 2;
    note_launch();
    if (ld == 64)
        dconv0_fwd_mma_kernel<64><<<blocks_for(units, 8, 148 * per_sm), 256, 0, s>>>(x, w, bias, mask, slope, a, B, S);
    else
        dconv0_fwd_mma_kernel<128><<<blocks_for(units, 8, 148 * per_sm), 256, 0, s>>>(x, w, bias, mask, slope, a, B, S);
}
int dconv0_wgrad_mma(const float* x, const bf16* dy, float* partial, int B, int S, cudaStream_t s, int ld) {
    const long tiles = static_cast<long>(B) * (S / 2) * (S / 32);
    const int blocks = blocks_for(tiles, 8, 148 * 4);
    note_launch();
    if (ld == 64)
        dconv0_wgrad_mma_kernel<64><<<blocks, 256, 0, s>>>(x, dy, partial, B, S);
    else
        dconv0_wgrad_mma_kernel<128><<<blocks, 256, 0, s>>>(x, dy, partial, B, S);
    return blocks;
}
void dconv0_dgrad_mma(const bf16* dy, const float* w, float* dx, int B, int S, cudaStream_t s, int ld, int accumulate) {
    const int O = S / 2, RB = 1024 / O;
    const int units = B * (O / RB);
    const int smem = (RB + 2) * O * kTPitch * 4;
    static bool ok = cudaFuncSetAttribute(dconv0_dgrad_mma_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024) ==
                         cudaSuccess &&
                     cudaFuncSetAttribute(dconv0_dgrad_mma_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024) ==
                         cudaSuccess;
    (void)ok;
    static int per_sm[2] = {0, 0};  // by image size: the band buffer of a 128 x 128 image is a little larger
    int& ps = per_sm[S == 64 ? 0 : 1];
    if (ps == 0 &&
        (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ps, dconv0_dgrad_mma_kernel<64>, 256, smem) != cudaSuccess || ps < 1))
        ps = 2;
    note_launch();
    const int grid = units < 148 * ps ? units : 148 * ps;
    if (ld == 64)
        dconv0_dgrad_mma_kernel<64><<<grid, 256, smem, s>>>(dy, w, dx, B, S, accumulate);
    else
        dconv0_dgrad_mma_kernel<128><<<grid, 256, smem, s>>>(dy, w, dx, B, S, accumulate);
}

}  // namespace sg
