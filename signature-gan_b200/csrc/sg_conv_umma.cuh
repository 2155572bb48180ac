// sg_conv_umma.cuh — host-side interface of the tcgen05 implicit-GEMM kernels.
//
// Activations are NHWC bf16. Three GEMM "row gather" modes share one kernel:
//   kConvS2 : 4x4 stride-2 pad-1 convolution (Discriminator forward, disc…:51-58;
//             also the data-gradient of ConvTranspose2d, gen…:46-54).
//   kConvT  : 4x4 stride-2 pad-1 transposed convolution, one output parity phase per
//             grid.z (Generator forward, gen…:46-54; also the data-gradient of Conv2d).
//   kPlain  : row-major [M][K] x [N][K]^T (Generator fc, gen…:125).
// The weight-gradient kernel contracts over pixels with both operands MN-major.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>

namespace sg {

enum ConvMode : int { kConvS2 = 0, kConvT = 1, kPlain = 2 };
enum EpiAct : int { kActNone = 0, kActRelu = 1, kActLeaky = 2 };

// Arguments of the forward/dgrad implicit GEMM. `GH x GW` is the grid the GEMM rows
// enumerate: the OUTPUT grid for kConvS2, the INPUT grid for kConvT.
struct ConvGemmArgs {
    CUtensorMap amap[4];  // kConvS2: the four (row,col) parity views of the input; else [0]
    CUtensorMap bmap;     // packed weights [N_total][taps*Cin] bf16, K contiguous
    int mode;
    int GH, GW, nimg;
    int M_total;  // nimg*GH*GW (kPlain: rows)
    int N_total;  // output channels
    int Cin;      // channels per tap (K = taps*Cin)
    // epilogue: v = acc (+bias[n]) ; (v = v*scale[n]+shift[n]) ; act ; (*mask[img][n]) ; (*gate')
    void* out;  // bf16 rows of `ldo` elements
    int ldo;
    const float* bias;
    const float* scale;
    const float* shift;
    int act;
    float slope;
    const float* mask;  // [nimg][ldmask] dropout keep-scale, or null
    int ldmask;
    const __nv_bfloat16* gate;  // saved activation at the output position: v *= (g>0 ? 1 : slope)
    float* stats_partial;       // optional per-CTA partial rows [chunks][2][N_total]: paired kConvT — sum / sum of squares of
                                // the stored outputs (conv_gemm_stats_chunks); BatchNorm gate — (sum d, sum d*y), see below
    // BatchNorm gate (training backward of the Generator): `gate` is the PRE-BatchNorm output y of the layer below and the
    // activation derivative is taken on y*gate_scale[n] + gate_shift[n] (the expression that layer's forward applied). With
    // stats_partial the epilogue also accumulates (sum d, sum d*y) of its stored outputs d — the raw reductions of that
    // layer's BatchNorm backward (bn_bwd_finalize), rows = conv_gemm_gate_stats_chunks() — replacing a pass over d and y.
    const float* gate_scale;
    const float* gate_shift;
    // Producer schedule (filled by launch_conv_gemm): entry e of a section = {A channel offset, packed (dx+1) | (dy+1)<<2 |
    // parity view<<4, B k-offset, 0} for the e-th K step; transposed convolutions keep one section per output parity.
    // It lives in the parameter (constant) bank so that the producer warp reads it with uniform loads.
    int4 tab[256];  // 256 entries: 16 taps x 16 K blocks (the widest layer of the 2x-width variant: 1024 channels); > 4 KB of
                    // kernel parameters needs CUDA 12.1+ (large kernel parameters)
};

struct WgradArgs {
    CUtensorMap cmap;     // coarse-grid tensor as [pixels][Cc] (2-D)
    CUtensorMap fmap[4];  // the four parity views of the fine-grid (2x) tensor
    int GH, GW, nimg;     // coarse grid
    int Mc, Nf;           // channels of coarse / fine tensors
    int k_tiles;          // ceil(nimg*GH*GW / 64)
    int splits;
    int plain;            // 1: no taps, fmap[0] is a 2-D [pixels][Nf] map (Generator fc weight gradient)
    int tpc;              // filter taps stacked along N per CTA (1, 2 or 4)
    float* partial;  // [splits][16][Mc][Nf]
    float* bias_partial;  // optional [splits][Mc]: sum over pixels of the coarse tensor (fused bias gradient)
};

// Tensor-map builders (host). Return 0 on success.
int make_map_2d(CUtensorMap* m, const void* base, uint64_t inner, uint64_t outer, uint64_t row_stride_elems,
                uint32_t box_inner, uint32_t box_outer);
// NHWC tensor [N][H][W][C] viewed with (row,col) parity (yp,xp) and step `step` (1 = plain view, 2 = parity view).
int make_map_nhwc(CUtensorMap* m, const void* base, int N, int H, int W, int C, int step, int yp, int xp,
                  uint32_t box_c, uint32_t box_w, uint32_t box_h, uint32_t box_n);

// Fills maps + geometry and launches. `w_packed` is [Cout][taps][Cin] bf16 with taps = 16 (ky*4+kx) for conv
// modes and 1 for kPlain. For kConvT all four phases are launched (grid.z = 4).
int launch_conv_gemm(ConvMode mode, const __nv_bfloat16* in, const __nv_bfloat16* w_packed, int nimg, int inH, int inW,
                     int Cin, int Cout, ConvGemmArgs epi /* only epilogue fields read */, cudaStream_t stream);

// Fused eval-mode tail of the Generator (sg_convt4_final.cu): last ConvT block (32 -> 32) + running-stat BatchNorm + ReLU
// + Conv3x3(32 -> 1) + bias + tanh; out (fp32) and out_u8 are (nimg, 1, 2 inH, 2 inW), either may be null.
bool convt4_final_supported(int inH, int inW, int Cin, int Cout);
int launch_convt4_final(const __nv_bfloat16* in, const __nv_bfloat16* w_packed, int nimg, int inH, int inW,
                        const float* scale, const float* shift, const float* w3, const float* b3, float* out,
                        uint8_t* out_u8, cudaStream_t stream);
// Number of partial rows a kConvT launch with `stats_partial` set writes, 0 if this shape cannot fuse the statistics.
int conv_gemm_stats_chunks(int nimg, int inH, int inW, int Cin, int Cout);
// Same for a kConvS2 launch with a BatchNorm gate (gate_scale / gate_shift / stats_partial): 0 = not fusable.
int conv_gemm_gate_stats_chunks(int nimg, int inH, int inW, int Cin, int Cout);

// dW[m][n][ky][kx] (fp32, PyTorch (M,N,4,4) layout) = sum_pix coarse[pix][m] * fine[2*pix-1+k][n].
// `partial` must hold splits*16*Mc*Nf floats. `accumulate` adds into dW instead of overwriting.
// dbias (optional, Mc floats): sum over pixels of `coarse` — the bias gradient of a Conv2d whose output gradient is
// `coarse` — computed by the same kernel (one more product with an all-ones operand); fat layers only.
int launch_wgrad(const __nv_bfloat16* coarse, const __nv_bfloat16* fine, int nimg, int cH, int cW, int Mc, int Nf,
                 float* partial, size_t partial_floats, float* dW, int accumulate, cudaStream_t stream,
                 float* dbias = nullptr);
size_t wgrad_partial_floats(int nimg, int cH, int cW, int Mc, int Nf);
// Between begin and end (same host thread) every launch_wgrad puts its split-K reduction on `side` (ordered after the
// weight-gradient kernel by `fork_ev`), so that it overlaps whatever the launch stream does next; the next launch_wgrad
// and wgrad_side_end make the launch stream wait for it (`join_ev`). Works inside stream capture (fork / join branches).
void wgrad_side_begin(cudaStream_t side, cudaEvent_t fork_ev, cudaEvent_t join_ev);
void wgrad_side_end(cudaStream_t stream);
cudaStream_t wgrad_side_fork(cudaStream_t stream);  // stream for other bucket-only kernels (ordered after `stream` so far)
void wgrad_side_mark();                             // ... which are pending until the next join

// Generator fc weight gradient (plain MN-major GEMM over the batch), un-permuting rows into dW (F, latent).
size_t fc_wgrad_partial_floats(int B, int F, int Kp);
int launch_fc_wgrad(const __nv_bfloat16* dy, const __nv_bfloat16* zp, int B, int C0, int Kp, int latent, float* partial,
                    size_t partial_floats, float* dW, cudaStream_t stream);

const char* umma_last_error();

}  // namespace sg
