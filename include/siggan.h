/* siggan.h — C ABI of libsiggan.so, the B200 (sm_100a) implementation of signature-Gan's
 * adversarial-training and sampling hot path.
 *
 * The reference (Nobita421/signature-Gan) is pure Python/PyTorch and has no FFI of its own; the
 * "operator interface" this library replaces is the set of torch.nn / torch.optim calls made by
 *   src/generator_vanilla_gan.py:189-209      Generator.forward            -> sg_g_forward / sg_g_backward
 *   src/discriminator_vanilla_gan.py:241-274  Discriminator.forward(_features) -> sg_d_forward / sg_d_backward
 *   src/vanilla_gan_model.py:107              nn.BCELoss                   -> sg_bce_forward / sg_bce_backward
 *   src/vanilla_gan_model.py:110-120          optim.Adam(...).step()       -> sg_adam_step
 *   src/vanilla_gan_model.py:180-336          train_*_step / train_step    -> sg_train_step
 *   src/utils/inference.py:129                ((x+1)*127.5).clip -> uint8  -> `out_u8` of sg_g_forward
 * The Python drop-in modules (the .py files in signature-gan_b200) bind these with ctypes; INTEGRATION.md shows
 * the stub.
 *
 * Conventions
 *  - Every pointer is a DEVICE pointer unless named host_*. Images are (B,1,S,S) fp32, latents (B,latent) fp32.
 *  - Parameters / gradients / Adam moments are flat fp32 buffers laid out in the reference's
 *    `named_parameters()` order (sg_tensor_info gives offset + shape); BatchNorm running statistics are
 *    a flat fp32 buffer laid out as [running_mean | running_var] per BN layer in module order.
 *  - All calls only ENQUEUE work on `stream` (a cudaStream_t passed as void*); no call synchronises unless
 *    it has to grow library-owned scratch (first call at a new maximum batch).
 *  - Return value: 0 = ok, negative = error (message from sg_last_error(), thread-local). Nothing throws
 *    or exits across this boundary. One sg_ctx per (process, device); a ctx is not re-entrant.
 *  - There is no CPU implementation behind these symbols.
 */
#ifndef SIGGAN_H
#define SIGGAN_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SG_ABI_VERSION 3

enum { SG_PREC_BF16 = 0, SG_PREC_FP32 = 1 };
enum { SG_NET_G = 0, SG_NET_D = 1 };

typedef struct sg_ctx sg_ctx;

typedef struct sg_config {
    int image_size;   /* 64 or 128 (gen…:106-107, disc…:121-122) */
    int latent_dim;   /* default 100 */
    int precision;    /* SG_PREC_BF16: bf16 operands, fp32 accumulate, tcgen05; SG_PREC_FP32: validation mode */
    float leaky_slope; /* 0.2 (disc…:46) */
    float bn_eps;      /* 1e-5 */
    float bn_momentum; /* 0.1 */
    float g_act_slope; /* Generator activation: 0 = ReLU (gen…:60,127); 0.2 = the ablation's ConfigurableGenerator with
                          activation="leaky_relu" (ablation…:204-207, 265-268) */
    int width_mult;    /* 0 / 1: the reference's channel ladders (gen…:131-149, disc…:131-194); 2: every hidden width doubled —
                          the "2x hidden width" of the width / resolution sweep. Not a configuration of the reference (its
                          base_features argument is inert): the same blocks with doubled channel counts; bf16 mode, ReLU */
} sg_config;

int sg_abi_version(void);
const char* sg_last_error(void);

/* Kernels enqueued by this library in this process so far (host-side count; bench.py's gpu_launches). */
unsigned long long sg_launch_count(void);

int sg_create(const sg_config* cfg, sg_ctx** out);
void sg_destroy(sg_ctx* ctx);

/* ---- layout of the flat buffers -------------------------------------------------------------- */
int sg_num_tensors(const sg_ctx* ctx, int net);
/* name is the reference state_dict key; shape has up to 4 dims (unused = 0). */
int sg_tensor_info(const sg_ctx* ctx, int net, int index, const char** name, long long* offset, int shape[4]);
long long sg_param_count(const sg_ctx* ctx, int net);
long long sg_g_stat_count(const sg_ctx* ctx);          /* floats in the running-stat buffer */
int sg_g_num_bn(const sg_ctx* ctx);
int sg_g_bn_info(const sg_ctx* ctx, int index, long long* mean_offset, long long* var_offset, int* channels);
size_t sg_g_workspace_bytes(const sg_ctx* ctx, int batch);   /* saved activations of one G forward */
size_t sg_d_workspace_bytes(const sg_ctx* ctx, int batch);   /* saved activations of one D forward */
long long sg_d_mask_count(const sg_ctx* ctx, int batch);     /* floats: sum_i batch*C_i */
/* element offset of the last conv block's weight in the flat D parameter / gradient buffer: [offset, count) is the
 * part of the bucket whose gradients the D backward finishes first (all-reduce overlap, see sg_train_step) */
long long sg_d_grad_tail_offset(const sg_ctx* ctx);
long long sg_d_feature_count(const sg_ctx* ctx);             /* 512*4*4 */

/* ---- Generator ------------------------------------------------------------------------------ */
/* bn_batch_stats != 0: training-mode BatchNorm (batch statistics; running stats updated in place).
 * ws may be NULL when no backward will follow (sampling). out_u8 (optional) receives the
 * ((x+1)*127.5).clip(0,255) uint8 image of utils/inference.py:129. */
int sg_g_forward(sg_ctx* ctx, const float* params, float* running_stats, const float* z, int batch,
                 int bn_batch_stats, void* ws, float* out_image, uint8_t* out_u8, void* stream);
/* grads_out: flat fp32, overwritten. dz_out optional (batch x latent). Must follow sg_g_forward with the same ws. */
int sg_g_backward(sg_ctx* ctx, const float* params, const void* ws, const float* grad_image, int batch,
                  int bn_batch_stats, float* grads_out, float* dz_out, void* stream);

/* ---- Discriminator -------------------------------------------------------------------------- */
/* masks: Dropout2d keep-scale per (n,c) for each block (sg_d_mask_count floats) or NULL for eval mode. */
int sg_d_forward(sg_ctx* ctx, const float* params, const float* x, int batch, const float* masks, void* ws,
                 float* prob_out, float* features_out, void* stream);
/* grads_out NULL => skip weight gradients; dx_out NULL => skip the image gradient. */
int sg_d_backward(sg_ctx* ctx, const float* params, const float* x, const void* ws, const float* masks,
                  const float* grad_prob, int batch, float* grads_out, float* dx_out, void* stream);
/* Byte offset of one saved tensor inside a forward workspace (parity tests read per-layer activations there and
 * replace them by the oracle's before running a single backward unit), or -1. Generator kinds: 0 padded latents,
 * 1 fc output before BatchNorm, 2 fc activation, 3 block `index` ConvT output before BatchNorm, 4 block `index`
 * activation, 5 image (fp32), 6 / 7 / 8 / 9 BatchNorm `index` mean / rstd / scale / shift (fp32; index 0 = fc, in NHWC
 * column order). Discriminator kinds: 0 block `index` activation, 1 probabilities (fp32). */
long long sg_ws_offset(const sg_ctx* ctx, int net, int batch, int kind, int index);

/* ---- one layer of a backward pass (parity tests: tests/test_gpu_layers.py feeds each layer the oracle's exact upstream
 * gradient). These run the same launchers as sg_d_backward / sg_g_backward on ONE unit of the chain; `ws` must come from
 * the matching forward. Activation-shaped tensors are NHWC in the context's activation type (bf16, or fp32 in
 * validation mode). Only the gradient entries of the unit's own parameters are written to grads_out.
 * Discriminator units: layer = number of conv blocks -> classifier (dz_in = d loss / d probability, fp32 (batch));
 * layer 1.. -> conv block `layer`; layer 0 -> first block (dx_out = image gradient, fp32). dz_in / dz_prev_out = gradient
 * w.r.t. the block's convolution output, i.e. after the LeakyReLU derivative and the dropout mask (disc…:51-75). */
int sg_d_backward_layer(sg_ctx* ctx, const float* params, const float* x, const void* ws, const float* masks, int layer,
                        const void* dz_in, int batch, float* grads_out, void* dz_prev_out, float* dx_out, void* stream);
/* Generator units: level = number of upsample blocks - 1 -> final Conv3x3 + tanh and the last block together (d_in =
 * d loss / d image, fp32); 0 <= level below that -> upsample block `level`; -1 -> the fc stage. d_in / d_prev_out =
 * gradient w.r.t. the stage's BatchNorm output with the ReLU derivative applied (gen…:58-60, 126-127). */
int sg_g_backward_layer(sg_ctx* ctx, const float* params, const void* ws, int level, const void* d_in, int batch,
                        int bn_batch_stats, float* grads_out, void* d_prev_out, void* stream);

/* Fills masks with {0, 1/(1-p)} from a counter-based generator (seed, offset). */
int sg_dropout_masks(sg_ctx* ctx, uint64_t seed, uint64_t offset, int batch, float p, float* masks_out, void* stream);

/* ---- nn.BCELoss on probabilities (log clamp -100, grad eps 1e-12) ----------------------------- */
int sg_bce_forward(const float* prob, const float* target, int n, float* loss_out, void* stream);
/* dprob = grad_loss[0] * (p - y) / max(p(1-p), 1e-12) / n */
int sg_bce_backward(const float* prob, const float* target, int n, const float* grad_loss, float* dprob_out,
                    void* stream);

/* ---- torch.optim.Adam (wd = 0, amsgrad = False) over a flat buffer ---------------------------- */
int sg_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, long long n, float lr,
                 float beta1, float beta2, float eps, long long step, void* stream);

/* ---- torch.nn.utils.spectral_norm on a Discriminator weight (disc…:61-62 Conv2d, :201-202 Linear) ---- */
/* w_orig = weight_orig viewed as (rows = Cout, cols = Cin*kh*kw); u (rows) / v (cols) = weight_u / weight_v,
 * updated IN PLACE by `power_iterations` power iterations (1 in training mode, 0 in eval mode):
 *   v <- W^T u / max(|W^T u|, eps);  u <- W v / max(|W v|, eps);  sigma = u.(W v);  w_out = W / sigma.
 * sigma_out: 1 float (device), kept by the caller for the backward. scratch: >= max(rows, 512) floats. */
int sg_spectral_norm_weight(const float* w_orig, float* u, float* v, int rows, int cols, int power_iterations,
                            float eps, float* w_out, float* sigma_out, float* scratch, void* stream);
/* grad holds dL/dw_eff on entry and dL/dw_orig = (grad - <grad, w_eff> u v^T) / sigma on return (u, v, sigma as
 * the matching sg_spectral_norm_weight call left them). */
int sg_spectral_norm_backward(const float* w_eff, const float* u, const float* v, const float* sigma, int rows,
                              int cols, float* grad, float* scratch, void* stream);

/* ---- Fused D step + G step (vanilla…:308-336 with n_critic = 1) ------------------------------- */
typedef struct sg_train_state {
    float* g_params; float* g_running_stats; float* g_exp_avg; float* g_exp_avg_sq;
    float* d_params; float* d_exp_avg; float* d_exp_avg_sq;
    long long g_step; long long d_step;            /* Adam step counts BEFORE this call */
    float g_lr, d_lr, beta1, beta2, eps;
    float label_smoothing;                         /* 0.9 */
    float dropout_p;                               /* 0.25; <= 0 disables */
    uint64_t seed; uint64_t offset;                /* dropout RNG counter (ignored when masks given) */
    const float* masks_real; const float* masks_fake; /* optional injected masks (parity tests) */
    int world_size;                                /* >1: caller all-reduces the gradient buckets between phases */
    float grad_scale;                              /* the Adam phases multiply the bucket by this (1/world_size when the
                                                      caller SUMs over ranks; 0 or 1 = gradients used as they are) */
} sg_train_state;

/* metrics_out (device, 12 floats): d_loss, d_loss_real, d_loss_fake, d_real_acc, d_fake_acc, d_real_mean,
 * d_fake_mean, g_loss, g_fake_mean, 0, 0, 0.  phase: 0 = whole step; 1 = D forward+backward only (grads in
 * d_grads), 2 = D Adam, 3 = G forward+backward (grads in g_grads), 4 = G Adam — phases let a data-parallel
 * caller all-reduce the flat gradient buckets between backward and update.
 * Overlapping the D all-reduce with backward: phase 11 = D forward + the classifier / last conv block part of the
 * backward (their gradients, the contiguous tail of the bucket from sg_d_grad_tail_offset() on — 76 % of it — are
 * final when it returns), phase 12 = the rest of the D backward; 11 followed by 12 equals phase 1. Likewise phase 31 =
 * the G step down to upsample block 0 (bucket final from sg_g_grad_tail_offset() on), phase 32 = the fc stage; 31
 * followed by 32 equals 3. With a communicator (sg_comm_init, world_size > 1) phase 0 runs the whole data-parallel
 * step: every gradient group is averaged over the ranks on the communication stream while the rest of its backward pass
 * runs, and both Adam updates use the averaged buckets. */
int sg_train_step(sg_ctx* ctx, sg_train_state* st, const float* real, const float* noise_d, const float* noise_g,
                  int batch, float* d_grads, float* g_grads, float* metrics_out, int phase, void* stream);

/* ---- input pipeline: augmentation of a device-resident 8-bit image pool --------------- */
/* Replaces the per-image PIL work of src/data_loader_signatures.py:154-219 `get_train_transforms` (applied in
 * SignatureDataset.__getitem__, :120-135): RandomRotation(fill=255) -> RandomAffine(degrees=0, scale, fill=255)
 * [-> horizontal flip] -> ToTensor -> Normalize(0.5, 0.5). Nearest-neighbour resampling on 8-bit pixels: the output
 * is bit-identical to torchvision + Pillow for the same sampled (angle, scale, flip).
 * sg_augment_params (HOST pointers, no GPU work): per image, Pillow's 16.16 fixed-point rotation coefficients
 * (6 ints: a0 a1 a2 a3 a4 a5 of `affine_fixed`, half-pixel offset folded into a2 / a5) and the start / step of the
 * scaling coordinates (4 doubles: a0, xo, a4, yo of `ImagingScaleAffine`). angles in degrees; NULL = no rotation /
 * no scaling. */
int sg_augment_params(const double* host_angles, const double* host_scales, int n, int image_size, int* host_rot_fixed,
                      double* host_scale_affine);
/* pool: (N, S, S) uint8; index: `batch` pool indices or NULL (= 0..batch-1); rot_fixed / scale_affine: device copies
 * of the tables above; flip: `batch` bytes or NULL; out: (batch, 1, S, S) fp32 in [-1, 1]. */
int sg_augment_batch(const uint8_t* pool, const int* index, const int* rot_fixed, const double* scale_affine,
                     const uint8_t* flip, int batch, int image_size, float* out, void* stream);

/* ---- evaluation: per-image ink statistics --------------------------------------------- */
/* One pass over (n_images, 1, H, W) fp32 images for src/utils/metrics.py:118-174 calculate_stroke_density /
 * calculate_foreground_ratio: per image, count_raw = #(x < threshold), count_rescaled = #((x + 1) / 2 < threshold)
 * (float32, as torch evaluates it) and the minimum. The reference rescales when the minimum of the WHOLE batch is
 * negative: the caller reduces `minimum` and picks the column. Integer counts: bit-exact. */
int sg_ink_stats(const float* images, int n_images, int pixels_per_image, float threshold, int* count_raw,
                 int* count_rescaled, float* minimum, void* stream);

/* ---- synchronised BatchNorm for data-parallel runs ------------------------------- */
/* With a callback set, every training-mode BatchNorm of the Generator (gen…:58,126) normalises with the statistics of
 * the GLOBAL batch: the per-channel sums (sum x, sum x^2 in forward; sum d, sum d*xhat in backward; 2*C floats) are
 * placed in `buf`, `fn(user, buf, count, stream)` must sum them in place over the `world_size` ranks (enqueued on
 * `stream`, e.g. an NCCL all-reduce; return 0 on success), and the row count is scaled by world_size. BatchNorm weight
 * / bias gradients stay local sums — the caller averages the gradient bucket as for every other parameter — so that
 * W ranks x B images reproduce one process at W*B images. buf: caller-owned device memory, >= 4 * the largest BatchNorm
 * channel count floats (16 * init_channels for the fc BatchNorm1d). fn == NULL switches back to local statistics. */
typedef int (*sg_allreduce_fn)(void* user, float* buf, long long count, void* stream);
int sg_set_sync_batchnorm(sg_ctx* ctx, sg_allreduce_fn fn, void* user, int world_size, float* buf,
                          long long buf_floats);

/* ---- data-parallel replicas: library-owned NCCL communicator for the flat gradient buckets ---- */
/* The reference is single-process (train…:494-502); a data-parallel launcher runs one process per GPU with identical
 * replicas and averages the flat fp32 gradient buckets over the ranks before each Adam update (losses are batch means,
 * vanilla…:107). NCCL is bound at run time (the libnccl.so.2 the host process already loaded, else the system one).
 * Rank 0 calls sg_comm_unique_id (128 bytes, HOST memory) and ships the id to the other ranks by any host channel;
 * every rank then calls sg_comm_init (collective: blocks until all `world_size` ranks arrive). */
int sg_comm_nccl_version(void);                               /* NCCL version code, -1 when libnccl.so.2 cannot be loaded */
int sg_comm_unique_id(void* host_id_out, size_t cap);         /* cap >= 128 */
int sg_comm_init(sg_ctx* ctx, const void* host_id, size_t id_bytes, int rank, int world_size);
int sg_comm_destroy(sg_ctx* ctx);
int sg_comm_world_size(const sg_ctx* ctx);                    /* 0 without a communicator */
/* grads[offset, offset+count) (elements of the `which` = SG_NET_G / SG_NET_D bucket; count < 0 = to its end) <- mean over
 * the ranks, in place. async = 0: enqueued in order on `stream`. async != 0: the reduction waits for what `stream` holds
 * so far and runs on the library's communication stream, overlapping whatever is enqueued on `stream` next (the rest of
 * a backward pass); sg_allreduce_join makes `stream` wait for every reduction started so far. */
int sg_allreduce_grads(sg_ctx* ctx, int which, float* grads, long long offset, long long count, int async, void* stream);
int sg_allreduce_join(sg_ctx* ctx, void* stream);
/* element offset of upsample block 0's weight in the flat G bucket: [offset, count) is final once the G backward has
 * passed block 0 (only the fc stage's gradients, [0, offset), are still to come) */
long long sg_g_grad_tail_offset(const sg_ctx* ctx);

/* ---- measurement aid: per-operation device time (CUDA events on the launch stream) --------------- */
/* on != 0 starts recording (and clears earlier records); every op of the plans is bracketed by two events. */
int sg_profile_enable(sg_ctx* ctx, int on);
/* Synchronises, then writes "name<TAB>ms<TAB>algorithmic flops<TAB>algorithmic bytes" lines; returns bytes written. */
long long sg_profile_dump(sg_ctx* ctx, char* buf, size_t cap);

#ifdef __cplusplus
}
#endif
#endif /* SIGGAN_H */
