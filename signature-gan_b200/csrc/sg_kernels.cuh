// sg_kernels.cuh — launchers of the CUDA-core kernels: everything on the path that is not a fat
// implicit GEMM (BatchNorm statistics/apply/backward, the 1-channel convolutions at either end of the
// networks, classifier + sigmoid + BCE, dropout masks, weight packing, Adam) plus the direct fp32
// convolutions used by the validation mode. Activations are NHWC, element type T in {float, bf16}.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>

namespace sg {

using bf16 = __nv_bfloat16;

// Fused epilogue shared by the direct convolutions (mirrors the tcgen05 epilogue in sg_conv_umma.cu).
struct Epi {
    const float* bias = nullptr;
    const float* scale = nullptr;
    const float* shift = nullptr;
    int act = 0;  // 0 none, 1 relu, 2 leaky(slope)
    float slope = 0.f;
    const float* mask = nullptr;  // [nimg][ldmask]
    int ldmask = 0;
    const void* gate = nullptr;  // T activation at the output position: v *= (g > 0 ? 1 : slope)
};

// Every launcher bumps this (host-side) counter once per kernel it enqueues; bench.py reports it as gpu_launches.
extern unsigned long long g_launches;
inline void note_launch(int n = 1) { g_launches += static_cast<unsigned long long>(n); }

constexpr int kMaxChunks = 592;  // 148 SMs x 4: upper bound on partial-sum rows of the column reductions

// ---- packing ---------------------------------------------------------------------------------
// src fp32 [A][B][16] -> dstAB bf16 [A][16][B], dstBA bf16 [B][16][A] (either may be null)
void pack_w16(const float* src, bf16* dstAB, bf16* dstBA, int A, int B, cudaStream_t s);
// Generator fc: rows permuted from NCHW feature f = c*16+hw to NHWC j = hw*C0+c, K padded to Kp.
void pack_fc(const float* W, const float* bias, bf16* Wp, float* biasp, int C0, int latent, int Kp, cudaStream_t s);
// Classifier weight (1, C*16) NCHW order -> NHWC order fp32
void pack_classifier(const float* w, float* wp, int C, cudaStream_t s);

// All packs of one network in ONE launch (<= 8 segments). kind 0: pack_w16, 1: pack_fc, 2: pack_classifier.
struct PackSeg {
    const float* src;
    const float* src2;  // fc bias
    bf16* ab;           // w16: [A][16][B]; fc: Wp
    bf16* ba;           // w16: [B][16][A]
    float* fdst;        // fc: permuted bias; classifier: permuted weight
    int A, B, Kp, kind, tile0;
};
struct PackPlan {
    PackSeg seg[8];
    int nseg = 0;
    int total_tiles = 0;
};
int pack_plan_add_w16(PackPlan& p, const float* src, bf16* ab, bf16* ba, int A, int B);
int pack_plan_add_fc(PackPlan& p, const float* W, const float* bias, bf16* Wp, float* biasp, int C0, int latent, int Kp);
int pack_plan_add_classifier(PackPlan& p, const float* w, float* wp, int C);
void pack_plan_launch(const PackPlan& p, cudaStream_t s);

// ---- generator side --------------------------------------------------------------------------
template <typename T>
void cast_pad_z(const float* z, T* zp, int B, int latent, int Kp, cudaStream_t s);

// Column reductions over a [rows][C] matrix, two-stage and deterministic.
//   mode 0: p0 = sum y,            p1 = sum y^2
//   mode 1: p0 = sum d,            p1 = sum d * (y - mean[c]) * rstd[c]       (d = first operand)
//   mode 2: p0 = sum y * rs[row],  p1 unused                                   (rs = per-row scale)
// partial must hold chunks*2*C floats; returns the number of chunks used.
template <typename T>
int col_reduce(int mode, const T* a, const T* y, const float* mean, const float* rstd, const float* rowscale,
               long rows, int C, float* partial, cudaStream_t s);

// BatchNorm forward bookkeeping. perm_c0 > 0: parameter index of column j is (j % perm_c0)*16 + j / perm_c0.
void bn_finalize(const float* partial, int chunks, long rows, int C, const float* gamma, const float* beta,
                 float* running_mean, float* running_var, float momentum, float eps, int batch_stats, int perm_c0,
                 float* mean, float* rstd, float* scale, float* shift, cudaStream_t s);
// Eval-mode coefficients (mean, rstd, scale, shift from the running statistics) of up to 8 BatchNorm layers in one launch.
struct BnEvalLayer {
    const float *gamma, *beta, *running_mean, *running_var;
    float *mean, *rstd, *scale, *shift;
    int first, perm_c0;  // first global channel index of the layer; perm_c0 as in bn_finalize
};
struct BnEvalPlan {
    BnEvalLayer layer[8];
    int n = 0, total = 0;
};
void bn_eval_all(const BnEvalPlan& plan, float eps, cudaStream_t s);
template <typename T>
void bn_apply_relu(const T* y, const float* scale, const float* shift, T* a, long rows, int C, float act_slope,
                   cudaStream_t s);  // act_slope 0 = ReLU, else LeakyReLU(act_slope)
// BatchNorm backward bookkeeping: dgamma/dbeta (parameter order) and the per-channel coefficients.
// raw_mean != null: the second partial is sum d*y (not d*xhat) and is converted here with mean/rstd.
void bn_bwd_finalize(const float* partial, int chunks, long rows, int C, const float* gamma, const float* rstd,
                     const float* raw_mean, int batch_stats, int perm_c0, float* dgamma, float* dbeta, float* k1,
                     float* k2, float* k3, cudaStream_t s);
template <typename T>
void bn_bwd_apply(const T* d, const T* y, const float* mean, const float* rstd, const float* k1, const float* k2,
                  const float* k3, T* dy, long rows, int C, cudaStream_t s);
// Generic: out0[pidx] = sum_chunks p0, out1[pidx] = sum_chunks p1 (null = skip); accumulate adds into out.
void col_finalize(const float* partial, int chunks, int C, int perm_c0, float* out0, float* out1, cudaStream_t s);

// Conv3x3(32->1) + bias + tanh (gen…:153-163) over the last NHWC level (sg_gfinal.cu). With scale/shift non-null `in`
// is the PRE-BatchNorm convolution output and relu(in*scale+shift) is applied while loading, so the normalised
// activation of the last upsample block is never written to HBM. out fp32 (B,1,S,S); out_u8 optional.
template <typename T>
void final_conv_tanh(const T* in, const float* scale, const float* shift, const float* w, const float* bias, float* out,
                     uint8_t* out_u8, int B, int S, int C, float act_slope, cudaStream_t s);
// Backward of the above in one pass over y: dbn = relu'(a) * convT3x3(dout*(1-out^2)), dW, dbias, plus the
// BatchNorm-backward partial sums of the last block, part_bn[chunk][2][C] = (sum dbn, sum dbn*y) (raw form, see
// bn_bwd_finalize). part_w must hold chunks*(C*9+1) floats. Returns the number of chunks (<= kMaxChunks).
template <typename T>
int final_conv_bwd(const float* dout, const float* out, const T* y, const float* scale, const float* shift,
                   const float* w, T* dbn, float* dW, float* dbias, float* part_w, float* part_bn, int B, int S, int C,
                   float act_slope, cudaStream_t s);
void vec_finalize(const float* partial, int chunks, int n, float* out_a, int na, float* out_b, cudaStream_t s);
// bf16 two-pass form (sg_gfinal_mma.cu): final_conv_bwd with dbn == nullptr computes only the reductions (nothing is
// written per pixel); after bn_bwd_finalize, final_conv_bwd_apply recomputes d from y and writes the
// BatchNorm-backward result dy = k1*(d - k2 - xhat*k3) directly — 3 activation-sized HBM passes instead of 5.
bool final_conv_bwd_two_pass(int S, int C, float act_slope);
void final_conv_bwd_apply(const float* dout, const float* out, const bf16* y, const float* scale, const float* shift,
                          const float* w, const float* mean, const float* rstd, const float* k1, const float* k2,
                          const float* k3, bf16* dy, int B, int S, int C, cudaStream_t s);
// records "<what>: C=<C> unsupported" for kernels_check (the next SG_KCHECK fails with it)
void note_unsupported(const char* what, int C);

// ---- discriminator side ----------------------------------------------------------------------
// Conv 4x4 s2 p1, 1 -> C channels (disc…:134-139) with bias + LeakyReLU + dropout mask.
template <typename T>
void d_conv0(const float* x, const float* w, const float* bias, const float* mask, float slope, T* a, int B, int S,
             int C, cudaStream_t s);
template <typename T>
void d_conv0_wgrad(const float* x, const T* dy, float* dW, float* partial, int B, int S, int C, cudaStream_t s);
template <typename T>
void d_conv0_dgrad(const T* dy, const float* w, float* dx, int B, int S, int C, cudaStream_t s);
// bf16 mode implementations of the three calls above (sg_dconv0.cu, warp-level mma.sync): 64 channels per launch inside an
// NHWC tensor of `ld` channels (64, or 128 for the 2x-width variant: two launches, pointers pre-offset per half).
void dconv0_fwd_mma(const float* x, const float* w, const float* bias, const float* mask, float slope, bf16* a, int B,
                    int S, cudaStream_t s, int ld = 64);
int dconv0_wgrad_mma(const float* x, const bf16* dy, float* partial, int B, int S, cudaStream_t s, int ld = 64);  // returns chunks
void dconv0_dgrad_mma(const bf16* dy, const float* w, float* dx, int B, int S, cudaStream_t s, int ld = 64,
                      int accumulate = 0);
// logit = <a, wp> + b ; prob = sigmoid(logit). a is [B][F] NHWC-flattened.
template <typename T>
void classifier_sigmoid(const T* a, const float* wp, const float* bias, float* prob, int B, int F, cudaStream_t s);
template <typename T>
void features_nchw(const T* a, float* feat, int B, int C, cudaStream_t s);  // (B, C*16) in reference order
// dlogit = dprob * p * (1-p)
void sigmoid_bwd(const float* prob, const float* dprob, float* dlogit, int B, cudaStream_t s);
// dy[b][j] = dlogit[b] * wp[j] * mask[b][c] * leaky'(a[b][j])
template <typename T>
void classifier_bwd_dy(const float* dlogit, const float* wp, const float* mask, const T* a, float slope, T* dy, int B,
                       int C, cudaStream_t s);
void sum_vector(const float* v, int n, float* out, cudaStream_t s);  // out[0] = sum v (deterministic)
void dropout_masks(uint64_t seed, uint64_t offset, long n, float p, float* out, cudaStream_t s);

// ---- loss ------------------------------------------------------------------------------------
void bce_forward(const float* prob, const float* target, int n, float* loss, cudaStream_t s);
void bce_backward(const float* prob, const float* target, int n, const float* grad_loss, float* dprob, cudaStream_t s);
// Fused metrics + dlogit for the training step. prob holds [real B | fake B].
void d_loss_metrics(const float* prob, int B, float smoothing, float* metrics, float* dlogit, cudaStream_t s);
void g_loss_metrics(const float* prob, int B, float* metrics, float* dlogit, cudaStream_t s);

// ---- optimizer -------------------------------------------------------------------------------
void adam_step(float* p, const float* g, float* m, float* v, long n, float lr, float b1, float b2, float eps, long step,
               cudaStream_t s);

// Device-resident per-step state of the fused training step (see step_prep_kernel).
struct StepCounters {
    long long g_step, d_step;
    unsigned long long dropout_offset;
    float adam_g[2], adam_d[2];  // {lr / (1 - b1^t), 1 / sqrt(1 - b2^t)}
};
void step_prep(StepCounters* c, int which /* 0 = G, 1 = D */, float lr, float b1, float b2,
               unsigned long long dropout_advance, cudaStream_t s);
void step_set(StepCounters* c, int field /* 0 g_step, 1 d_step, 2 dropout_offset */, long long value, cudaStream_t s);
void adam_step_dev(float* p, const float* g, float* m, float* v, long n, float b1, float b2, float eps, const float* scal,
                   float grad_scale, cudaStream_t s);
// Adam over the whole flat buffer AND the network's bf16 / permuted weight packs in ONE launch (the
// optimizer emits the packed copies): `plan` lists the pack segments whose sources lie inside [p, p + n); the blocks of
// a segment update their parameters and write the packs from the updated values, the rest of the buffer is updated by
// plain element-wise blocks. Same arithmetic, element by element, as adam_step_dev followed by pack_plan_launch.
// Returns 0, or -1 when the plan does not tile the buffer (caller falls back to the two launches).
int adam_pack_step_dev(const PackPlan& plan, float* p, const float* g, float* m, float* v, long n, float b1, float b2,
                       float eps, const float* scal, float grad_scale, cudaStream_t s);
void dropout_masks_dev(uint64_t seed, const unsigned long long* offset_ptr, long n, float p, float* out, cudaStream_t s);

// ---- direct convolutions (validation mode / fallback). Weights are the fp32 master tensors in PyTorch
// layout, addressed as w[o*so + i*si + ky*4 + kx] where o = output channel of THIS op, i = input channel.
template <typename T>
void conv_s2_direct(const T* in, const float* w, long so, long si, const Epi& e, T* out, int B, int inH, int inW,
                    int Cin, int Cout, cudaStream_t s);
template <typename T>
void convT_direct(const T* in, const float* w, long so, long si, const Epi& e, T* out, int B, int inH, int inW, int Cin,
                  int Cout, cudaStream_t s);
// dW[m][n][16] = sum coarse[pix][m] * fine[2*pix-1+tap][n]
template <typename T>
void wgrad_direct(const T* coarse, const T* fine, float* dW, int B, int cH, int cW, int Mc, int Nf, cudaStream_t s);
// y[b][j] = sum_i zp[b][i] * W[f(j)][i] + bias[f(j)]  (fp32 master weights, NHWC-permuted output)
template <typename T>
void fc_direct(const T* zp, int Kp, const float* W, const float* bias, T* y, int B, int C0, int latent,
               cudaStream_t s);
// dW[f(j)][i] = sum_b dy[b][j] * zp[b][i]
template <typename T>
void fc_wgrad_direct(const T* dy, const T* zp, int Kp, float* dW, int B, int C0, int latent, cudaStream_t s);
template <typename T>
void fc_dz_direct(const T* dy, const float* W, float* dz, int B, int C0, int latent, cudaStream_t s);

const char* kernels_last_error();
// ---- spectral normalisation of a (rows, cols) weight matrix (sg_spectral.cu) -----------------------
// scratch: at least max(rows, 512) floats.
void spectral_norm_weight(const float* w_orig, float* u, float* v, int rows, int cols, int power_iterations, float eps,
                          float* w_out, float* sigma, float* scratch, cudaStream_t s);
void spectral_norm_backward(const float* w_eff, const float* u, const float* v, const float* sigma, int rows, int cols,
                            float* grad, float* scratch, cudaStream_t s);

// ---- input-pipeline augmentation on a device-resident uint8 pool (sg_augment.cu) --------------------
int augment_batch(const uint8_t* pool, const int* index, const int* rot, const double* sc, const uint8_t* flip, int batch,
                  int size, float* out, cudaStream_t s);
void augment_params(const double* angles, const double* scales, int n, int size, int* rot, double* sc);  // host only

// ---- per-image ink statistics (sg_metrics.cu): counts of x < t and of (x + 1) / 2 < t, and the minimum ------
void ink_stats(const float* images, int n_images, int pixels, float threshold, int* count_raw, int* count_rescaled,
               float* minimum, cudaStream_t s);

int kernels_check(const char* what);  // cudaGetLastError -> 0 / -1 (message kept)

}  // namespace sg
