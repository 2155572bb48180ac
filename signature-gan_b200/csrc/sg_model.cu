// sg_model.cu — the Generator / Discriminator execution plans and the extern "C" boundary (include/siggan.h).
//
// Data layout in HBM
//   activations : NHWC, bf16 (SG_PREC_BF16) or fp32 (SG_PREC_FP32); images (C = 1) are the caller's fp32 (B,1,S,S)
//   parameters  : the caller's flat fp32 buffers in reference named_parameters() order (PyTorch layouts);
//                 bf16 tensor-core packs [Cout][ky*4+kx][Cin] are rebuilt from them at the start of every
//                 forward (3.9 M parameters: a few microseconds) so external optimizers / load_state_dict just work
//   gradients   : flat fp32 buffers with the same offsets as the parameters
//   workspace   : per-forward saved activations (caller-allocated so autograd owns their lifetime)
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <type_traits>
#include <vector>

#include "../../include/siggan.h"
#include "sg_comm.cuh"
#include "sg_conv_umma.cuh"
#include "sg_kernels.cuh"

using sg::bf16;

static thread_local char t_err[768] = "";
static int fail(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_err, sizeof(t_err), fmt, ap);
    va_end(ap);
    return -1;
}
#define SG_TRY(expr)                 \
    do {                             \
        int rc_ = (expr);            \
        if (rc_ != 0) return rc_;    \
    } while (0)
#define SG_UMMA(expr)                                                  \
    do {                                                               \
        if ((expr) != 0) return fail("%s", sg::umma_last_error());     \
    } while (0)
#define SG_COMM(expr)                                                  \
    do {                                                               \
        if ((expr) != 0) return fail("%s", sg::comm_last_error());     \
    } while (0)
#define SG_KCHECK(what)                                                   \
    do {                                                                  \
        if (sg::kernels_check(what) != 0) return fail("%s", sg::kernels_last_error()); \
    } while (0)

namespace {

constexpr size_t kAlign = 256;
inline size_t align_up(size_t v) { return (v + kAlign - 1) / kAlign * kAlign; }

struct TensorInfo {
    std::string name;
    long long offset;
    int shape[4];
    long long numel() const {
        long long n = 1;
        for (int i = 0; i < 4; ++i)
            if (shape[i] > 0) n *= shape[i];
        return n;
    }
};

// Bumped whenever a library-owned buffer moves: captured CUDA graphs hold the old addresses and must be rebuilt.
static unsigned long long g_alloc_epoch = 0;

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes) {
        if (bytes <= cap) return 0;
        ++g_alloc_epoch;
        if (p) cudaFree(p);  // implicit device sync: in-flight users are done before the old block goes away
        p = nullptr;
        cap = 0;
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e != cudaSuccess) return fail("cudaMalloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
        cap = bytes;
        return 0;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

// Optional per-operation device timing (CUDA events on the launch stream) with the algorithmic work of each op.
struct ProfRec {
    std::string name;
    double flops, bytes;
    cudaEvent_t a, b;
};
struct Profiler {
    bool on = false;
    std::vector<ProfRec> recs;
    std::vector<cudaEvent_t> pool;
    cudaEvent_t get() {
        if (!pool.empty()) {
            cudaEvent_t e = pool.back();
            pool.pop_back();
            return e;
        }
        cudaEvent_t e;
        cudaEventCreate(&e);
        return e;
    }
};
struct ProfScope {
    Profiler* p;
    cudaStream_t s;
    size_t idx = 0;
    ProfScope(Profiler* prof, cudaStream_t stream, const char* name, double flops, double bytes) : p(prof), s(stream) {
        if (!p->on) return;
        ProfRec r{name, flops, bytes, p->get(), p->get()};
        cudaEventRecord(r.a, s);
        idx = p->recs.size();
        p->recs.push_back(r);
    }
    ~ProfScope() {
        if (p->on) cudaEventRecord(p->recs[idx].b, s);
    }
};
#define PROF(name, flops, bytes) ProfScope prof_scope_(&c->prof, s, name, flops, bytes)

struct BNInfo {
    int C;
    long long gamma_off, beta_off;  // into the flat G parameter buffer
    long long mean_off, var_off;    // into the flat running-stat buffer
};

// Carved view of one Generator forward's saved state.
struct GWs {
    char* zp;
    char* fc_y;
    char* fc_a;
    char* y[6];
    char* a[6];
    float* out;
    float* mean[7];
    float* rstd[7];
    float* scale[7];
    float* shift[7];
    size_t bytes;
};
struct DWs {
    char* a[6];
    float* prob;
    size_t bytes;
};

// One captured phase of the fused training step. The key holds everything the captured launches bake in.
struct GraphKey {
    int phase, B, flags;
    const void* ptr[10];
    float f[8];
    unsigned long long seed;
    bool operator==(const GraphKey& o) const { return memcmp(this, &o, sizeof(GraphKey)) == 0; }
};
struct GraphEntry {
    GraphKey key;
    cudaGraphExec_t exec = nullptr;
    int seen = 0, n_launches = 0;
    bool bad = false;
    unsigned long long epoch = 0, last = 0;
};

}  // namespace

struct sg_ctx {
    sg_config cfg;
    int S, L, ND, Kp, es;  // image size, #upsample blocks, #downsample blocks, padded latent, activation element size
    int gch[7], dch[7];
    std::vector<TensorInfo> gt, dt;
    long long g_count, d_count, g_stats;
    BNInfo bn[7];
    // tensor indices
    int g_fc_w, g_fc_b, g_up_w[6], g_final_w, g_final_b;
    int d_conv_w[6], d_conv_b[6], d_cls_w, d_cls_b;
    // library-owned device memory
    DevBuf packs;    // bf16 weight packs + permuted fp32 vectors
    bf16 *fc_Wp, *g_packF[6], *g_packB[6], *d_packF[6], *d_packB[6];
    float *fc_biasp, *cls_wp;
    DevBuf bufA, bufB, dpre, wpart, cpart, small, g_ws, d_ws, x2, masks2, dximg, gws_tmp;
    const float* d_pack_src = nullptr;  // parameter buffer the Discriminator packs were last built from
    const float* g_pack_src = nullptr;  // same for the Generator (several modules may share one context)
    // Inside ONE sg_train_step(phase 0) call nobody but the library writes the parameters, so the G-step phase can reuse
    // the Generator packs of the D-step phase (G is only updated at the very end) and the Discriminator packs its fused
    // Adam + pack kernel just wrote: set by the phase-0 sequence around phase 3 / 31, false everywhere else
    bool skip_g_pack = false, skip_d_pack = false;
    int device = 0;                     // CUDA device ordinal this context was created on
    // SyncBN (sg_set_sync_batchnorm): per-channel BatchNorm sums are all-reduced over the ranks through the caller's callback
    sg_allreduce_fn sync_fn = nullptr;
    void* sync_user = nullptr;
    int sync_world = 1;
    float* sync_buf = nullptr;  // caller-owned, >= 4 * (largest BatchNorm channel count) floats
    long long sync_cap = 0;
    bool u8_only = false;  // current sg_g_forward call wants the uint8 image only (no caller-visible fp32 image)
    float *dlogit, *k1, *k2, *k3;
    int scratch_batch = 0;
    Profiler prof;
    // fused training step: device-resident step counters + CUDA graphs of the phases (see sg_train_step)
    DevBuf counters_buf;
    sg::StepCounters* counters = nullptr;
    long long mirror_g_step = -1, mirror_d_step = -1;     // host mirror of the device counters (-1 = unknown)
    unsigned long long mirror_drop = ~0ull;
    bool mirror_drop_known = false;
    std::vector<GraphEntry> graphs;
    unsigned long long graph_tick = 0;
    cudaStream_t cap_stream = nullptr;
    // side stream + events for the split-K reductions of the weight gradients (wgrad_side_begin)
    cudaStream_t red_stream = nullptr;
    cudaEvent_t red_fork = nullptr, red_join = nullptr;
    // data-parallel replica: library-owned NCCL communicator (sg_comm_init) for the flat gradient buckets
    sg::Comm comm;
};

namespace {

void add_tensor(std::vector<TensorInfo>& v, long long& off, const std::string& name, int a, int b = 0, int c = 0,
                int d = 0) {
    TensorInfo t;
    t.name = name;
    t.offset = off;
    t.shape[0] = a;
    t.shape[1] = b;
    t.shape[2] = c;
    t.shape[3] = d;
    off += t.numel();
    v.push_back(t);
}

int g_spatial(const sg_ctx* c, int i) { return 8 << i; }        // output H=W of upsample block i
int d_spatial(const sg_ctx* c, int i) { return c->S >> (i + 1); }  // output H=W of downsample block i

GWs carve_g(const sg_ctx* c, void* base, int B) {
    GWs w;
    size_t off = 0;
    auto take = [&](size_t bytes) {
        char* p = base ? static_cast<char*>(base) + off : nullptr;
        off += align_up(bytes);
        return p;
    };
    const size_t es = c->es;
    const int F0 = c->gch[0] * 16;
    w.zp = take(static_cast<size_t>(B) * c->Kp * es);
    w.fc_y = take(static_cast<size_t>(B) * F0 * es);
    w.fc_a = take(static_cast<size_t>(B) * F0 * es);
    for (int i = 0; i < c->L; ++i) {
        const size_t n = static_cast<size_t>(B) * g_spatial(c, i) * g_spatial(c, i) * c->gch[i + 1] * es;
        w.y[i] = take(n);
        w.a[i] = take(n);
    }
    w.out = reinterpret_cast<float*>(take(static_cast<size_t>(B) * c->S * c->S * 4));
    for (int i = 0; i <= c->L; ++i) {
        const size_t n = static_cast<size_t>(c->bn[i].C) * 4;
        w.mean[i] = reinterpret_cast<float*>(take(n));
        w.rstd[i] = reinterpret_cast<float*>(take(n));
        w.scale[i] = reinterpret_cast<float*>(take(n));
        w.shift[i] = reinterpret_cast<float*>(take(n));
    }
    w.bytes = off;
    return w;
}

DWs carve_d(const sg_ctx* c, void* base, int B) {
    DWs w;
    size_t off = 0;
    auto take = [&](size_t bytes) {
        char* p = base ? static_cast<char*>(base) + off : nullptr;
        off += align_up(bytes);
        return p;
    };
    for (int i = 0; i < c->ND; ++i)
        w.a[i] = take(static_cast<size_t>(B) * d_spatial(c, i) * d_spatial(c, i) * c->dch[i + 1] * c->es);
    w.prob = reinterpret_cast<float*>(take(static_cast<size_t>(B) * 4));
    w.bytes = off;
    return w;
}

long long mask_offset(const sg_ctx* c, int batch, int layer) {
    long long off = 0;
    for (int i = 0; i < layer; ++i) off += static_cast<long long>(batch) * c->dch[i + 1];
    return off;
}

int ensure_scratch(sg_ctx* c, int B) {
    if (B <= c->scratch_batch) return 0;
    const size_t lvl = static_cast<size_t>(B) * c->S * c->S * c->gch[c->L] * c->es;  // largest activation level
    SG_TRY(c->bufA.ensure(lvl));
    SG_TRY(c->bufB.ensure(lvl));
    SG_TRY(c->dpre.ensure(static_cast<size_t>(B) * c->S * c->S * 4));
    size_t wp = 0;
    for (int i = 0; i < c->L; ++i) {
        const int h = g_spatial(c, i) / 2;
        const size_t n = sg::wgrad_partial_floats(B, h, h, c->gch[i], c->gch[i + 1]);
        if (n > wp) wp = n;
    }
    for (int i = 1; i < c->ND; ++i) {
        const int h = d_spatial(c, i);
        const size_t n = sg::wgrad_partial_floats(B, h, h, c->dch[i + 1], c->dch[i]);
        if (n > wp) wp = n;
    }
    {
        const size_t n = sg::fc_wgrad_partial_floats(B, c->gch[0] * 16, c->Kp);
        if (n > wp) wp = n;
    }
    SG_TRY(c->wpart.ensure(wp * 4));
    SG_TRY(c->cpart.ensure(static_cast<size_t>(sg::kMaxChunks) * 2 * 2048 * 4));
    const size_t b_pad = (static_cast<size_t>(B) + 63) / 64 * 64;  // keeps k1..k3 16-byte aligned (float4 loads)
    const size_t kmax = static_cast<size_t>(c->gch[0]) * 16;  // the widest BatchNorm: the fc stage's BatchNorm1d
    SG_TRY(c->small.ensure((b_pad + 3 * kmax) * 4));
    c->dlogit = static_cast<float*>(c->small.p);
    c->k1 = c->dlogit + b_pad;
    c->k2 = c->k1 + kmax;
    c->k3 = c->k2 + kmax;
    c->scratch_batch = B;
    return 0;
}

// ------------------------------------------------------------------------------------------------
// SyncBN: the per-CTA partial rows [chunks][2][C] of a BatchNorm reduction are folded into one row, summed over the
// ranks by the caller's all-reduce and handed to the same finalize kernels with the global row count.
// ------------------------------------------------------------------------------------------------
inline bool sync_bn_on(const sg_ctx* c) { return c->sync_fn != nullptr && c->sync_world > 1; }

int sync_partials(sg_ctx* c, const float*& partial, int& chunks, long& rows, int C, cudaStream_t s) {
    if (!sync_bn_on(c)) return 0;
    if (4LL * C > c->sync_cap) return fail("SyncBN buffer too small: %lld floats for %d channels", c->sync_cap, C);
    sg::col_finalize(partial, chunks, C, 0, c->sync_buf, c->sync_buf + C, s);
    SG_KCHECK("sync_partials");
    if (c->sync_fn(c->sync_user, c->sync_buf, 2LL * C, s) != 0) return fail("SyncBN: the all-reduce callback failed");
    partial = c->sync_buf;
    chunks = 1;
    rows *= c->sync_world;
    return 0;
}

// Backward: dgamma / dbeta stay the LOCAL sums (the gradient bucket's all-reduce averages them like every other
// parameter gradient); the coefficients of the data gradient (k2, k3 = batch means of d and d*xhat) become global.
int sync_bn_bwd_coefficients(sg_ctx* c, const float* partial, int chunks, long rows, int C, const float* gamma,
                             const float* rstd, const float* raw_mean, int perm_c0, cudaStream_t s) {
    if (!sync_bn_on(c)) return 0;
    SG_TRY(sync_partials(c, partial, chunks, rows, C, s));
    sg::bn_bwd_finalize(partial, chunks, rows, C, gamma, rstd, raw_mean, 1, perm_c0, c->sync_buf + 2L * C,
                        c->sync_buf + 3L * C, c->k1, c->k2, c->k3, s);
    return 0;
}

// ------------------------------------------------------------------------------------------------
// weight packs
// ------------------------------------------------------------------------------------------------
int pack_generator(sg_ctx* c, const float* params, cudaStream_t s) {
    if (c->skip_g_pack && c->g_pack_src == params) return 0;
    c->g_pack_src = params;
    if (c->cfg.precision != SG_PREC_BF16) return 0;
    PROF("g.pack", 0, 10.0 * c->g_count);
    sg::PackPlan plan;
    int rc = sg::pack_plan_add_fc(plan, params + c->gt[c->g_fc_w].offset, params + c->gt[c->g_fc_b].offset, c->fc_Wp,
                                  c->fc_biasp, c->gch[0], c->cfg.latent_dim, c->Kp);
    for (int i = 0; i < c->L; ++i)  // W (Cin, Cout, 4, 4): forward pack [Cout][16][Cin] = BA, dgrad pack [Cin][16][Cout] = AB
        rc |= sg::pack_plan_add_w16(plan, params + c->gt[c->g_up_w[i]].offset, c->g_packB[i], c->g_packF[i], c->gch[i],
                                    c->gch[i + 1]);
    if (rc) return fail("pack_generator: pack plan overflow");
    sg::pack_plan_launch(plan, s);
    SG_KCHECK("pack_generator");
    return 0;
}
int discriminator_pack_plan(sg_ctx* c, const float* params, sg::PackPlan& plan) {
    int rc = sg::pack_plan_add_classifier(plan, params + c->dt[c->d_cls_w].offset, c->cls_wp, c->dch[c->ND]);
    if (c->cfg.precision == SG_PREC_BF16) {
        for (int i = 1; i < c->ND; ++i)  // W (Cout, Cin, 4, 4): forward pack [Cout][16][Cin] = AB, dgrad pack = BA
            rc |= sg::pack_plan_add_w16(plan, params + c->dt[c->d_conv_w[i]].offset, c->d_packF[i], c->d_packB[i],
                                        c->dch[i + 1], c->dch[i]);
    }
    return rc;
}
int pack_discriminator(sg_ctx* c, const float* params, cudaStream_t s) {
    if (c->skip_d_pack && c->d_pack_src == params) return 0;
    PROF("d.pack", 0, 10.0 * c->d_count);
    c->d_pack_src = params;
    sg::PackPlan plan;
    if (discriminator_pack_plan(c, params, plan)) return fail("pack_discriminator: pack plan overflow");
    sg::pack_plan_launch(plan, s);
    SG_KCHECK("pack_discriminator");
    return 0;
}

sg::ConvGemmArgs epi_args(void* out, int ldo) {
    sg::ConvGemmArgs a;
    memset(&a, 0, sizeof(a));
    a.out = out;
    a.ldo = ldo;
    return a;
}

// ------------------------------------------------------------------------------------------------
// Generator
// ------------------------------------------------------------------------------------------------
// SIGGAN_FUSED_TAIL=0 keeps the last block and the Conv3x3 as two kernels (A/B comparison)
static bool fused_tail_enabled() {
    static int on = -1;
    if (on < 0) {
        const char* e = getenv("SIGGAN_FUSED_TAIL");
        on = (e && e[0] == '0') ? 0 : 1;
    }
    return on != 0;
}

// Scope of a backward plan: with SIGGAN_SIDE_REDUCE=1 the split-K reductions of the weight gradients (and the other
// bucket-only kernels that use wgrad_side_fork) run on the context's side stream; by default they stay in line on the
// launch stream. The destructor joins the side stream back into the launch stream.
struct SideReduce {
    cudaStream_t s;
    bool on = false;
    SideReduce(sg_ctx* c, cudaStream_t stream) : s(stream) {
        static int enabled = -1;
        if (enabled < 0) {
            const char* e = getenv("SIGGAN_SIDE_REDUCE");
            enabled = (e && e[0] == '1') ? 1 : 0;  // opt-in: measured neutral within the +-0.05 ms box-to-box noise
        }
        if (!enabled || c->cfg.precision != SG_PREC_BF16) return;
        if (!c->red_stream) {
            if (cudaStreamCreateWithFlags(&c->red_stream, cudaStreamNonBlocking) != cudaSuccess ||
                cudaEventCreateWithFlags(&c->red_fork, cudaEventDisableTiming) != cudaSuccess ||
                cudaEventCreateWithFlags(&c->red_join, cudaEventDisableTiming) != cudaSuccess) {
                cudaGetLastError();
                c->red_stream = nullptr;
                return;
            }
        }
        sg::wgrad_side_begin(c->red_stream, c->red_fork, c->red_join);
        on = true;
    }
    ~SideReduce() {
        if (on) sg::wgrad_side_end(s);
    }
};

// SIGGAN_FUSED_BN_REDUCE=0: BatchNorm-backward reductions as separate passes (A/B comparison)
static bool fused_bn_reduce_enabled() {
    static int on = -1;
    if (on < 0) {
        const char* e = getenv("SIGGAN_FUSED_BN_REDUCE");
        on = (e && e[0] == '0') ? 0 : 1;
    }
    return on != 0;
}

template <typename T>
int g_forward_t(sg_ctx* c, const float* params, float* stats, const float* z, int B, int train, void* ws_ptr,
                float* out_image, uint8_t* out_u8, bool save, cudaStream_t s) {
    constexpr bool kTC = std::is_same<T, bf16>::value;
    GWs w = carve_g(c, ws_ptr, B);
    const int F0 = c->gch[0] * 16;
    const int latent = c->cfg.latent_dim;
    const bool fused_eval = !train && !save;  // fold running-stat BN + ReLU into the producing GEMM's epilogue
    const float gs = c->cfg.g_act_slope;      // 0: ReLU (gen…:60,127); > 0: the ablation's LeakyReLU generator (ablation…:204-207)
    const int g_act = gs != 0.f ? sg::kActLeaky : sg::kActRelu;
    SG_TRY(pack_generator(c, params, s));
    const double es = c->es;
    if (z) {  // z == nullptr: the caller already cast the latents into the workspace (fused step, outside its graph)
        PROF("g.cast_z", 0, B * (4.0 * latent + es * c->Kp));
        sg::cast_pad_z<T>(z, reinterpret_cast<T*>(w.zp), B, latent, c->Kp, s);
    }
    if (!train) {
        PROF("g.bn_eval", 0, 0);
        sg::BnEvalPlan plan;
        for (int i = 0; i <= c->L; ++i) {
            sg::BnEvalLayer& l = plan.layer[plan.n++];
            l.gamma = params + c->bn[i].gamma_off;
            l.beta = params + c->bn[i].beta_off;
            l.running_mean = stats + c->bn[i].mean_off;
            l.running_var = stats + c->bn[i].var_off;
            l.mean = w.mean[i];
            l.rstd = w.rstd[i];
            l.scale = w.scale[i];
            l.shift = w.shift[i];
            l.first = plan.total;
            l.perm_c0 = i == 0 ? c->gch[0] : 0;
            plan.total += c->bn[i].C;
        }
        sg::bn_eval_all(plan, c->cfg.bn_eps, s);
    }
    // ---- fc + BN1d + ReLU (gen…:124-128), output already NHWC (B,4,4,C0)
    {
        T* dst = reinterpret_cast<T*>(fused_eval ? w.fc_a : w.fc_y);
        {
        PROF("g.fc", 2.0 * B * latent * F0, es * B * (double)(F0 + c->Kp));
        if (kTC) {
            sg::ConvGemmArgs e = epi_args(dst, F0);
            e.bias = c->fc_biasp;
            if (fused_eval) {
                e.scale = w.scale[0];
                e.shift = w.shift[0];
                e.act = g_act;
                e.slope = gs;
            }
            SG_UMMA(sg::launch_conv_gemm(sg::kPlain, reinterpret_cast<const bf16*>(w.zp), c->fc_Wp, B, 1, 1, c->Kp, F0, e,
                                         s));
        } else {
            sg::fc_direct<T>(reinterpret_cast<const T*>(w.zp), c->Kp, params + c->gt[c->g_fc_w].offset,
                             params + c->gt[c->g_fc_b].offset, reinterpret_cast<T*>(w.fc_y), B, c->gch[0], latent, s);
            dst = reinterpret_cast<T*>(w.fc_y);
        }
        }
        if (!(kTC && fused_eval)) {
            if (train) {
                PROF("g.fc.bn_stats", 0, es * B * (double)F0);
                int chunks = sg::col_reduce<T>(0, reinterpret_cast<const T*>(w.fc_y), nullptr, nullptr, nullptr,
                                               nullptr, B, F0, static_cast<float*>(c->cpart.p), s);
                const float* part = static_cast<float*>(c->cpart.p);
                long nrows = B;
                SG_TRY(sync_partials(c, part, chunks, nrows, F0, s));
                sg::bn_finalize(part, chunks, nrows, F0, params + c->bn[0].gamma_off,
                                params + c->bn[0].beta_off, stats + c->bn[0].mean_off, stats + c->bn[0].var_off,
                                c->cfg.bn_momentum, c->cfg.bn_eps, 1, c->gch[0], w.mean[0], w.rstd[0], w.scale[0],
                                w.shift[0], s);
            }
            PROF("g.fc.bn_apply", 0, 2.0 * es * B * (double)F0);
            sg::bn_apply_relu<T>(reinterpret_cast<const T*>(w.fc_y), w.scale[0], w.shift[0],
                                 reinterpret_cast<T*>(w.fc_a), B, F0, gs, s);
        }
    }
    // ---- upsample blocks: ConvT 4x4 s2 p1 (no bias) + BN2d + ReLU (gen…:46-60)
    const char* in = w.fc_a;
    float* out = save ? w.out : out_image;
    bool tail_done = false;
    for (int i = 0; i < c->L; ++i) {
        const int ih = g_spatial(c, i) / 2, Cin = c->gch[i], Cout = c->gch[i + 1];
        if (kTC && fused_eval && i == c->L - 1 && gs == 0.f && fused_tail_enabled() &&
            sg::convt4_final_supported(ih, ih, Cin, Cout)) {
            // last block + Conv3x3 + tanh in one kernel: the full-resolution 32-channel level never reaches HBM
            const bool u8_only = out_u8 && c->u8_only;
            PROF("g.tail_fused", 2.0 * B * ih * ih * 16.0 * Cin * Cout + 2.0 * B * c->S * c->S * 9.0 * Cout,
                 (double)B * (es * ih * ih * Cin + c->S * c->S * ((u8_only ? 0.0 : 4.0) + (out_u8 ? 1.0 : 0.0))));
            if (sg::launch_convt4_final(reinterpret_cast<const bf16*>(in), c->g_packF[i], B, ih, ih, w.scale[i + 1],
                                        w.shift[i + 1], params + c->gt[c->g_final_w].offset,
                                        params + c->gt[c->g_final_b].offset, u8_only ? nullptr : out, out_u8, s))
                return fail("g_forward: fused tail launch failed: %s", cudaGetErrorString(cudaGetLastError()));
            tail_done = true;
            break;
        }
        const long rows = static_cast<long>(B) * 4 * ih * ih;
        const bool fuse = kTC && fused_eval;
        T* dst = reinterpret_cast<T*>(fuse ? w.a[i] : w.y[i]);
        const std::string nm = "g.up" + std::to_string(i);
        const double cflops = 2.0 * B * ih * ih * 16.0 * Cin * Cout;
        int stat_chunks = 0;
        {
        PROF(nm.c_str(), cflops, es * ((double)B * ih * ih * Cin + (double)rows * Cout));
        if (kTC) {
            sg::ConvGemmArgs e = epi_args(dst, Cout);
            if (fuse) {
                e.scale = w.scale[i + 1];
                e.shift = w.shift[i + 1];
                e.act = g_act;
                e.slope = gs;
            } else if (train) {
                stat_chunks = sg::conv_gemm_stats_chunks(B, ih, ih, Cin, Cout);  // batch statistics from the epilogue
                if (stat_chunks > 0) e.stats_partial = static_cast<float*>(c->cpart.p);
            }
            SG_UMMA(sg::launch_conv_gemm(sg::kConvT, reinterpret_cast<const bf16*>(in), c->g_packF[i], B, ih, ih, Cin,
                                         Cout, e, s));
        } else {
            sg::Epi e;
            sg::convT_direct<T>(reinterpret_cast<const T*>(in), params + c->gt[c->g_up_w[i]].offset, 16,
                                static_cast<long>(Cout) * 16, e, dst, B, ih, ih, Cin, Cout, s);
        }
        }
        if (!fuse) {
            if (train) {
                PROF((nm + ".bn_stats").c_str(), 0, stat_chunks > 0 ? 0.0 : es * (double)rows * Cout);
                int chunks = stat_chunks > 0
                                 ? stat_chunks
                                 : sg::col_reduce<T>(0, reinterpret_cast<const T*>(w.y[i]), nullptr, nullptr, nullptr,
                                                     nullptr, rows, Cout, static_cast<float*>(c->cpart.p), s);
                const float* part = static_cast<float*>(c->cpart.p);
                long nrows = rows;
                SG_TRY(sync_partials(c, part, chunks, nrows, Cout, s));
                sg::bn_finalize(part, chunks, nrows, Cout, params + c->bn[i + 1].gamma_off,
                                params + c->bn[i + 1].beta_off, stats + c->bn[i + 1].mean_off,
                                stats + c->bn[i + 1].var_off, c->cfg.bn_momentum, c->cfg.bn_eps, 1, 0, w.mean[i + 1],
                                w.rstd[i + 1], w.scale[i + 1], w.shift[i + 1], s);
            }
            if (i < c->L - 1) {  // the last level is normalised on the fly inside final_conv_tanh
                PROF((nm + ".bn_apply").c_str(), 0, 2.0 * es * (double)rows * Cout);
                sg::bn_apply_relu<T>(reinterpret_cast<const T*>(w.y[i]), w.scale[i + 1], w.shift[i + 1],
                                     reinterpret_cast<T*>(w.a[i]), rows, Cout, gs, s);
            }
        }
        in = (fuse || i < c->L - 1) ? w.a[i] : w.y[i];
    }
    // ---- Conv3x3 + tanh (gen…:153-163)
    if (!tail_done) {
    PROF("g.final", 2.0 * B * c->S * c->S * 9.0 * c->gch[c->L], (double)B * c->S * c->S * (es * c->gch[c->L] + 4.0));
    const bool affine = !(kTC && fused_eval);
    sg::final_conv_tanh<T>(reinterpret_cast<const T*>(in), affine ? w.scale[c->L] : nullptr,
                           affine ? w.shift[c->L] : nullptr, params + c->gt[c->g_final_w].offset,
                           params + c->gt[c->g_final_b].offset, out, out_u8, B, c->S, c->gch[c->L], gs, s);
    }
    if (save && out_image) {
        cudaError_t e = cudaMemcpyAsync(out_image, w.out, static_cast<size_t>(B) * c->S * c->S * 4,
                                        cudaMemcpyDeviceToDevice, s);
        if (e != cudaSuccess) return fail("g_forward: copy of the image failed: %s", cudaGetErrorString(e));
    }
    SG_KCHECK("g_forward");
    return 0;
}

// only_level = kAllLevels: the whole backward. Otherwise ONE unit of it (sg_g_backward_layer, parity tests): level
// L-1 = final Conv3x3 + tanh together with the last block (input: grad_image), 0 <= level < L-1 = that block and -1 = the
// fc stage (input `inject`: the ReLU'-masked gradient w.r.t. the stage's BatchNorm output, NHWC, T); the masked gradient
// handed to the stage below is copied to d_prev_out.
constexpr int kAllLevels = -100;
template <typename T>
int g_backward_t(sg_ctx* c, const float* params, const void* ws_ptr, const float* grad_image, int B, int train,
                 float* grads, float* dz, cudaStream_t s, int only_level = kAllLevels, const void* inject = nullptr,
                 void* d_prev_out = nullptr, int stage = 0) {
    constexpr bool kTC = std::is_same<T, bf16>::value;
    GWs w = carve_g(c, const_cast<void*>(ws_ptr), B);
    SideReduce side_reduce(c, s);
    float* cpart = static_cast<float*>(c->cpart.p);
    float* wpart = static_cast<float*>(c->wpart.p);
    char* cur = static_cast<char*>(c->bufA.p);
    char* nxt = static_cast<char*>(c->bufB.p);
    const int L = c->L;
    const double es = c->es;
    const bool single = only_level != kAllLevels;
    const int lvl_hi = single ? only_level : L - 1, lvl_lo = single ? (only_level < 0 ? 0 : only_level) : 0;
    if (single && only_level != L - 1) {
        const int C = only_level < 0 ? c->gch[0] * 16 : c->gch[only_level + 1];
        const size_t px = only_level < 0 ? 1 : static_cast<size_t>(g_spatial(c, only_level)) * g_spatial(c, only_level);
        cudaMemcpyAsync(cur, inject, static_cast<size_t>(B) * px * C * sizeof(T), cudaMemcpyDeviceToDevice, s);
    }
    // stage 1: everything down to upsample block 0 (those gradients, the bucket from sg_g_grad_tail_offset() on, are
    // final afterwards); stage 2: the fc stage, continuing from the scratch buffer stage 1 left; 0: both
    const bool run_final = (!single || only_level == L - 1) && stage != 2;
    const bool run_fc = (!single || only_level < 0) && stage != 1;
    if (stage == 2 && (L & 1)) {  // stage 1 swapped the scratch buffers once per block
        char* t = cur;
        cur = nxt;
        nxt = t;
    }
    // One pass over the last block's pre-BatchNorm output: d/d(bn output), final-conv dW/dbias, and the
    // BatchNorm-backward reductions of the last block (so that block needs no separate reduction pass below).
    float* part_bn = cpart + static_cast<size_t>(sg::kMaxChunks) * (9 * c->gch[L] + 1);
    int last_chunks = 0;
    // bf16 with batch statistics: pass 1 = reductions only, pass 2 (below) recomputes d and applies BatchNorm backward
    const float gs = c->cfg.g_act_slope;
    const bool two_pass = kTC && train && sg::final_conv_bwd_two_pass(c->S, c->gch[L], gs);
    if (run_final) {
    PROF("g.final_bwd", 4.0 * B * c->S * c->S * 9.0 * c->gch[L],
         (double)B * c->S * c->S * (8.0 + (two_pass ? 1.0 : 2.0) * es * c->gch[L]));
    last_chunks = sg::final_conv_bwd<T>(grad_image, w.out, reinterpret_cast<const T*>(w.y[L - 1]), w.scale[L], w.shift[L],
                                        params + c->gt[c->g_final_w].offset, two_pass ? nullptr : reinterpret_cast<T*>(cur),
                                        grads + c->gt[c->g_final_w].offset, grads + c->gt[c->g_final_b].offset, cpart,
                                        part_bn, B, c->S, c->gch[L], gs, s);
    }
    // BatchNorm-backward reductions of the level about to be processed that the data-gradient kernel of the level above
    // already produced in its epilogue (raw form: sum d, sum d*y per CTA); 0 = none, run the reduction pass
    int fused_chunks = 0;
    float* part_fused = cpart;
    for (int i = ((only_level < 0 && single) || stage == 2) ? -1 : lvl_hi; i >= lvl_lo; --i) {
        const int oh = g_spatial(c, i), ih = oh / 2, Cin = c->gch[i], Cout = c->gch[i + 1];
        const long rows = static_cast<long>(B) * oh * oh;
        const BNInfo& bn = c->bn[i + 1];
        const std::string nm = "g.up" + std::to_string(i);
        const double cflops = 2.0 * B * ih * ih * 16.0 * Cin * Cout;
        // BatchNorm2d backward (cur holds relu'-masked d/d(bn output)); result dy overwrites cur
        if (i == L - 1) {
            sg::bn_bwd_finalize(part_bn, last_chunks, rows, Cout, params + bn.gamma_off, w.rstd[i + 1], w.mean[i + 1],
                                train, 0, grads + bn.gamma_off, grads + bn.beta_off, c->k1, c->k2, c->k3, s);
            if (train)
                SG_TRY(sync_bn_bwd_coefficients(c, part_bn, last_chunks, rows, Cout, params + bn.gamma_off, w.rstd[i + 1],
                                                w.mean[i + 1], 0, s));
        } else if (fused_chunks > 0) {
            sg::bn_bwd_finalize(part_fused, fused_chunks, rows, Cout, params + bn.gamma_off, w.rstd[i + 1], w.mean[i + 1],
                                train, 0, grads + bn.gamma_off, grads + bn.beta_off, c->k1, c->k2, c->k3, s);
            if (train)
                SG_TRY(sync_bn_bwd_coefficients(c, part_fused, fused_chunks, rows, Cout, params + bn.gamma_off,
                                                w.rstd[i + 1], w.mean[i + 1], 0, s));
        } else {
            PROF((nm + ".bn_bwd_reduce").c_str(), 0, 2.0 * es * (double)rows * Cout);
            const int chunks = sg::col_reduce<T>(1, reinterpret_cast<const T*>(cur), reinterpret_cast<const T*>(w.y[i]),
                                                 w.mean[i + 1], w.rstd[i + 1], nullptr, rows, Cout, cpart, s);
            sg::bn_bwd_finalize(cpart, chunks, rows, Cout, params + bn.gamma_off, w.rstd[i + 1], nullptr, train, 0,
                                grads + bn.gamma_off, grads + bn.beta_off, c->k1, c->k2, c->k3, s);
            if (train)
                SG_TRY(sync_bn_bwd_coefficients(c, cpart, chunks, rows, Cout, params + bn.gamma_off, w.rstd[i + 1], nullptr,
                                                0, s));
        }
        if (i == L - 1 && two_pass) {
            PROF((nm + ".bn_bwd_apply").c_str(), 4.0 * B * c->S * c->S * 9.0 * Cout, 2.0 * es * (double)rows * Cout);
            sg::final_conv_bwd_apply(grad_image, w.out, reinterpret_cast<const bf16*>(w.y[i]), w.scale[L], w.shift[L],
                                     params + c->gt[c->g_final_w].offset, w.mean[L], w.rstd[L], c->k1, c->k2, c->k3,
                                     reinterpret_cast<bf16*>(cur), B, c->S, Cout, s);
        } else {
        PROF((nm + ".bn_bwd_apply").c_str(), 0, 3.0 * es * (double)rows * Cout);
        sg::bn_bwd_apply<T>(reinterpret_cast<const T*>(cur), reinterpret_cast<const T*>(w.y[i]), w.mean[i + 1],
                            w.rstd[i + 1], c->k1, c->k2, c->k3, reinterpret_cast<T*>(cur), rows, Cout, s);
        }
        const char* xin = i == 0 ? w.fc_a : w.a[i - 1];
        float* dW = grads + c->gt[c->g_up_w[i]].offset;
        if (kTC) {
            {
            PROF((nm + ".wgrad").c_str(), cflops, es * ((double)B * ih * ih * Cin + (double)rows * Cout));
            SG_UMMA(sg::launch_wgrad(reinterpret_cast<const bf16*>(xin), reinterpret_cast<const bf16*>(cur), B, ih, ih,
                                     Cin, Cout, wpart, c->wpart.cap / 4, dW, 0, s));
            }
            PROF((nm + ".dgrad").c_str(), cflops, es * (2.0 * B * ih * ih * Cin + (double)rows * Cout));
            sg::ConvGemmArgs e = epi_args(nxt, Cin);
            e.gate = reinterpret_cast<const bf16*>(xin);  // ReLU' (slope 0) / LeakyReLU' of the previous block
            e.slope = gs;
            // BatchNorm gate: the block below is an upsample block whose forward applied act(y * scale + shift) to its
            // saved pre-BatchNorm output y: gate on that expression and let the epilogue produce that block's
            // BatchNorm-backward reductions (sum d, sum d*y), so that no separate pass over d and y is needed
            fused_chunks = (i >= 1 && fused_bn_reduce_enabled()) ? sg::conv_gemm_gate_stats_chunks(B, oh, oh, Cout, Cin) : 0;
            if (fused_chunks > sg::kMaxChunks) fused_chunks = 0;
            if (fused_chunks > 0) {
                e.gate = reinterpret_cast<const bf16*>(w.y[i - 1]);
                e.gate_scale = w.scale[i];
                e.gate_shift = w.shift[i];
                e.stats_partial = part_fused;
            }
            SG_UMMA(sg::launch_conv_gemm(sg::kConvS2, reinterpret_cast<const bf16*>(cur), c->g_packB[i], B, oh, oh, Cout,
                                         Cin, e, s));
        } else {
            sg::wgrad_direct<T>(reinterpret_cast<const T*>(xin), reinterpret_cast<const T*>(cur), dW, B, ih, ih, Cin,
                                Cout, s);
            sg::Epi e;
            e.gate = xin;
            e.slope = gs;
            sg::conv_s2_direct<T>(reinterpret_cast<const T*>(cur), params + c->gt[c->g_up_w[i]].offset,
                                  static_cast<long>(Cout) * 16, 16, e, reinterpret_cast<T*>(nxt), B, oh, oh, Cout, Cin,
                                  s);
        }
        char* t = cur;
        cur = nxt;
        nxt = t;
    }
    if (single && only_level >= 0 && d_prev_out) {
        const int Cp = c->gch[only_level];
        const size_t px = static_cast<size_t>(g_spatial(c, only_level) / 2) * (g_spatial(c, only_level) / 2);
        cudaMemcpyAsync(d_prev_out, cur, static_cast<size_t>(B) * px * Cp * sizeof(T), cudaMemcpyDeviceToDevice, s);
    }
    if (!run_fc) {
        SG_KCHECK("g_backward");
        return 0;
    }
    // ---- fc: BatchNorm1d backward, weight / bias gradients
    const int F0 = c->gch[0] * 16, latent = c->cfg.latent_dim;
    const BNInfo& bn = c->bn[0];
    PROF("g.fc.bwd", 2.0 * B * F0 * latent, 6.0 * es * B * (double)F0);
    int chunks = sg::col_reduce<T>(1, reinterpret_cast<const T*>(cur), reinterpret_cast<const T*>(w.fc_y), w.mean[0],
                                   w.rstd[0], nullptr, B, F0, cpart, s);
    sg::bn_bwd_finalize(cpart, chunks, B, F0, params + bn.gamma_off, w.rstd[0], nullptr, train, c->gch[0],
                        grads + bn.gamma_off, grads + bn.beta_off, c->k1, c->k2, c->k3, s);
    if (train)
        SG_TRY(sync_bn_bwd_coefficients(c, cpart, chunks, B, F0, params + bn.gamma_off, w.rstd[0], nullptr, c->gch[0], s));
    sg::bn_bwd_apply<T>(reinterpret_cast<const T*>(cur), reinterpret_cast<const T*>(w.fc_y), w.mean[0], w.rstd[0], c->k1,
                        c->k2, c->k3, reinterpret_cast<T*>(cur), B, F0, s);
    chunks = sg::col_reduce<T>(0, reinterpret_cast<const T*>(cur), nullptr, nullptr, nullptr, nullptr, B, F0, cpart, s);
    sg::col_finalize(cpart, chunks, F0, c->gch[0], grads + c->gt[c->g_fc_b].offset, nullptr, s);
    float* dWfc = grads + c->gt[c->g_fc_w].offset;
    if (kTC) {
        SG_UMMA(sg::launch_fc_wgrad(reinterpret_cast<const bf16*>(cur), reinterpret_cast<const bf16*>(w.zp), B, c->gch[0],
                                    c->Kp, latent, wpart, c->wpart.cap / 4, dWfc, s));
    } else {
        sg::fc_wgrad_direct<T>(reinterpret_cast<const T*>(cur), reinterpret_cast<const T*>(w.zp), c->Kp, dWfc, B,
                               c->gch[0], latent, s);
    }
    if (dz) sg::fc_dz_direct<T>(reinterpret_cast<const T*>(cur), params + c->gt[c->g_fc_w].offset, dz, B, c->gch[0], latent, s);
    SG_KCHECK("g_backward");
    return 0;
}

// ------------------------------------------------------------------------------------------------
// Discriminator
// ------------------------------------------------------------------------------------------------
template <typename T>
int d_forward_t(sg_ctx* c, const float* params, const float* x, int B, const float* masks, void* ws_ptr,
                float* prob_out, float* feat_out, cudaStream_t s) {
    constexpr bool kTC = std::is_same<T, bf16>::value;
    DWs w = carve_d(c, ws_ptr, B);
    const float slope = c->cfg.leaky_slope;
    SG_TRY(pack_discriminator(c, params, s));
    const double es = c->es;
    {
    const int o0 = c->S / 2;
    PROF("d.c0", 2.0 * B * o0 * o0 * 16.0 * c->dch[1], (double)B * (4.0 * c->S * c->S + es * o0 * o0 * c->dch[1]));
    sg::d_conv0<T>(x, params + c->dt[c->d_conv_w[0]].offset, params + c->dt[c->d_conv_b[0]].offset,
                   masks ? masks + mask_offset(c, B, 0) : nullptr, slope, reinterpret_cast<T*>(w.a[0]), B, c->S,
                   c->dch[1], s);
    }
    for (int i = 1; i < c->ND; ++i) {
        const int ih = d_spatial(c, i - 1), Cin = c->dch[i], Cout = c->dch[i + 1];
        const std::string nm = "d.c" + std::to_string(i);
        PROF(nm.c_str(), 2.0 * B * (ih / 2) * (ih / 2) * 16.0 * Cin * Cout,
             es * B * ((double)ih * ih * Cin + (double)(ih / 2) * (ih / 2) * Cout));
        const float* mk = masks ? masks + mask_offset(c, B, i) : nullptr;
        const float* bias = params + c->dt[c->d_conv_b[i]].offset;
        if (kTC) {
            sg::ConvGemmArgs e = epi_args(w.a[i], Cout);
            e.bias = bias;
            e.act = sg::kActLeaky;
            e.slope = slope;
            e.mask = mk;
            e.ldmask = Cout;
            SG_UMMA(sg::launch_conv_gemm(sg::kConvS2, reinterpret_cast<const bf16*>(w.a[i - 1]), c->d_packF[i], B, ih, ih,
                                         Cin, Cout, e, s));
        } else {
            sg::Epi e;
            e.bias = bias;
            e.act = 2;
            e.slope = slope;
            e.mask = mk;
            e.ldmask = Cout;
            sg::conv_s2_direct<T>(reinterpret_cast<const T*>(w.a[i - 1]), params + c->dt[c->d_conv_w[i]].offset,
                                  static_cast<long>(Cin) * 16, 16, e, reinterpret_cast<T*>(w.a[i]), B, ih, ih, Cin, Cout,
                                  s);
        }
    }
    const int Cl = c->dch[c->ND];
    PROF("d.cls", 2.0 * B * Cl * 16, es * B * Cl * 16.0);
    sg::classifier_sigmoid<T>(reinterpret_cast<const T*>(w.a[c->ND - 1]), c->cls_wp, params + c->dt[c->d_cls_b].offset,
                              w.prob, B, Cl * 16, s);
    if (prob_out) {
        cudaError_t e = cudaMemcpyAsync(prob_out, w.prob, static_cast<size_t>(B) * 4, cudaMemcpyDeviceToDevice, s);
        if (e != cudaSuccess) return fail("d_forward: copy of probabilities failed: %s", cudaGetErrorString(e));
    }
    if (feat_out) sg::features_nchw<T>(reinterpret_cast<const T*>(w.a[c->ND - 1]), feat_out, B, Cl, s);
    SG_KCHECK("d_forward");
    return 0;
}

// dlogit: d(loss)/d(logit) per sample (already includes sigmoid').
// stage 0: whole backward; 1: classifier + last conv block only (the tail of the gradient bucket is final afterwards);
// 2: the remaining blocks (continues from the scratch buffers stage 1 left).
// only_layer >= 0 (sg_d_backward_layer, parity tests): ONE unit of the backward — ND = the classifier (input dlogit), 1..ND-1
// = that conv block, 0 = the first block (weight gradient + image gradient); `inject` = the gradient w.r.t. the block's
// convolution output (after LeakyReLU' and the dropout mask), NHWC, T; the same quantity of the block below is copied to
// dz_prev_out.
template <typename T>
int d_backward_t(sg_ctx* c, const float* params, const float* x, const void* ws_ptr, const float* masks,
                 const float* dlogit, int B, float* grads, float* dx, cudaStream_t s, int stage = 0, int only_layer = -1,
                 const void* inject = nullptr, void* dz_prev_out = nullptr) {
    constexpr bool kTC = std::is_same<T, bf16>::value;
    DWs w = carve_d(c, const_cast<void*>(ws_ptr), B);
    SideReduce side_reduce(c, s);
    const float slope = c->cfg.leaky_slope;
    float* cpart = static_cast<float*>(c->cpart.p);
    float* wpart = static_cast<float*>(c->wpart.p);
    char* cur = static_cast<char*>(c->bufA.p);
    char* nxt = static_cast<char*>(c->bufB.p);
    const int last = c->ND - 1, Cl = c->dch[c->ND];
    const double es = c->es;
    const bool single = only_layer >= 0;
    auto act_bytes = [&](int layer) {
        return static_cast<size_t>(B) * d_spatial(c, layer) * d_spatial(c, layer) * c->dch[layer + 1] * sizeof(T);
    };
    if (single && only_layer < c->ND)
        cudaMemcpyAsync(cur, inject, act_bytes(only_layer), cudaMemcpyDeviceToDevice, s);
    if (stage == 2) {  // stage 1 processed block `last` and swapped the scratch buffers once
        char* t = cur;
        cur = nxt;
        nxt = t;
    }
    if (single ? only_layer == c->ND : stage != 2) {
    PROF("d.cls_bwd", 4.0 * B * Cl * 16, 3.0 * es * B * Cl * 16.0);
    if (grads) {
        // the classifier's weight / bias gradients only feed the bucket: on the reduction branch, next to the data path
        cudaStream_t sb = sg::wgrad_side_fork(s);
        const int chunks = sg::col_reduce<T>(2, reinterpret_cast<const T*>(w.a[last]), nullptr, nullptr, nullptr, dlogit,
                                             B, Cl * 16, cpart, sb);
        sg::col_finalize(cpart, chunks, Cl * 16, Cl, grads + c->dt[c->d_cls_w].offset, nullptr, sb);
        sg::sum_vector(dlogit, B, grads + c->dt[c->d_cls_b].offset, sb);
        sg::wgrad_side_mark();
    }
    sg::classifier_bwd_dy<T>(dlogit, c->cls_wp, masks ? masks + mask_offset(c, B, last) : nullptr,
                             reinterpret_cast<const T*>(w.a[last]), slope, reinterpret_cast<T*>(cur), B, Cl, s);
    }
    if (single && only_layer == c->ND) {
        if (dz_prev_out) cudaMemcpyAsync(dz_prev_out, cur, act_bytes(last), cudaMemcpyDeviceToDevice, s);
        SG_KCHECK("d_backward");
        return 0;
    }
    const int i_hi = single ? only_layer : (stage == 2 ? last - 1 : last);
    const int i_lo = single ? (only_layer > 0 ? only_layer : 1 << 30) : (stage == 1 ? last : 1);
    for (int i = i_hi; i >= i_lo; --i) {
        const int oh = d_spatial(c, i), Cin = c->dch[i], Cout = c->dch[i + 1];
        const long rows = static_cast<long>(B) * oh * oh;
        const std::string nm = "d.c" + std::to_string(i);
        const double cflops = 2.0 * (double)rows * 16.0 * Cin * Cout;
        if (grads) {
            float* dW = grads + c->dt[c->d_conv_w[i]].offset;
            {
            PROF((nm + ".wgrad").c_str(), cflops, es * ((double)rows * Cout + 4.0 * rows * Cin));
            // bf16: the bias gradient (column sums of dy) comes out of the weight-gradient kernel (all-ones operand)
            if (kTC)
                SG_UMMA(sg::launch_wgrad(reinterpret_cast<const bf16*>(cur), reinterpret_cast<const bf16*>(w.a[i - 1]), B,
                                         oh, oh, Cout, Cin, wpart, c->wpart.cap / 4, dW, 0, s,
                                         grads + c->dt[c->d_conv_b[i]].offset));
            else
                sg::wgrad_direct<T>(reinterpret_cast<const T*>(cur), reinterpret_cast<const T*>(w.a[i - 1]), dW, B, oh,
                                    oh, Cout, Cin, s);
            }
            if (!kTC) {
                PROF((nm + ".dbias").c_str(), 0, es * (double)rows * Cout);
                const int chunks = sg::col_reduce<T>(0, reinterpret_cast<const T*>(cur), nullptr, nullptr, nullptr,
                                                     nullptr, rows, Cout, cpart, s);
                sg::col_finalize(cpart, chunks, Cout, 0, grads + c->dt[c->d_conv_b[i]].offset, nullptr, s);
            }
        }
        // data gradient = transposed conv of dy, gated by the previous block's LeakyReLU' and dropout mask
        const float* mk = masks ? masks + mask_offset(c, B, i - 1) : nullptr;
        PROF((nm + ".dgrad").c_str(), cflops, es * ((double)rows * Cout + 8.0 * rows * Cin));
        if (kTC) {
            sg::ConvGemmArgs e = epi_args(nxt, Cin);
            e.gate = reinterpret_cast<const bf16*>(w.a[i - 1]);
            e.slope = slope;
            e.mask = mk;
            e.ldmask = Cin;
            SG_UMMA(sg::launch_conv_gemm(sg::kConvT, reinterpret_cast<const bf16*>(cur), c->d_packB[i], B, oh, oh, Cout,
                                         Cin, e, s));
        } else {
            sg::Epi e;
            e.gate = w.a[i - 1];
            e.slope = slope;
            e.mask = mk;
            e.ldmask = Cin;
            sg::convT_direct<T>(reinterpret_cast<const T*>(cur), params + c->dt[c->d_conv_w[i]].offset, 16,
                                static_cast<long>(Cin) * 16, e, reinterpret_cast<T*>(nxt), B, oh, oh, Cout, Cin, s);
        }
        char* t = cur;
        cur = nxt;
        nxt = t;
    }
    const int o0 = c->S / 2;
    if (single && only_layer > 0) {
        if (dz_prev_out) cudaMemcpyAsync(dz_prev_out, cur, act_bytes(only_layer - 1), cudaMemcpyDeviceToDevice, s);
        SG_KCHECK("d_backward");
        return 0;
    }
    if (stage == 1) {
        SG_KCHECK("d_backward");
        return 0;
    }
    if (grads) {
        PROF("d.c0.wgrad", 2.0 * B * o0 * o0 * 16.0 * c->dch[1], (double)B * (4.0 * c->S * c->S + es * o0 * o0 * c->dch[1]));
        sg::d_conv0_wgrad<T>(x, reinterpret_cast<const T*>(cur), grads + c->dt[c->d_conv_w[0]].offset, cpart, B, c->S,
                             c->dch[1], s);
    }
    if (dx) {
        PROF("d.c0.dgrad", 2.0 * B * o0 * o0 * 16.0 * c->dch[1], (double)B * (4.0 * c->S * c->S + es * o0 * o0 * c->dch[1]));
        sg::d_conv0_dgrad<T>(reinterpret_cast<const T*>(cur), params + c->dt[c->d_conv_w[0]].offset, dx, B, c->S,
                             c->dch[1], s);
    }
    SG_KCHECK("d_backward");
    return 0;
}

// Every entry point runs on the context's device, whatever the caller's current device is (kernel launches, scratch
// allocation and tensor-map encoding all go to the current device).
struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != dev) cudaSetDevice(dev); else prev = -1;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

// Context-less entry points run on the device that owns their first device pointer.
struct PtrDeviceGuard : DeviceGuard {
    static int device_of(const void* p) {
        cudaPointerAttributes a;
        int cur = 0;
        cudaGetDevice(&cur);
        if (p && cudaPointerGetAttributes(&a, p) == cudaSuccess && a.type == cudaMemoryTypeDevice) return a.device;
        cudaGetLastError();
        return cur;
    }
    explicit PtrDeviceGuard(const void* p) : DeviceGuard(device_of(p)) {}
};

#define DISPATCH_T(ctx, fn, ...) \
    ((ctx)->cfg.precision == SG_PREC_BF16 ? fn<bf16>(__VA_ARGS__) : fn<float>(__VA_ARGS__))

}  // namespace

// ---- fused training step ------------------------------------------------------------------------
// Everything that changes from step to step lives in device memory (Adam step counts, dropout counter) or is staged
// by a few launches OUTSIDE the captured region (real batch -> x2, latents -> workspace, injected masks), so that each
// phase is ONE cudaGraphLaunch: ~110 kernel launches per step otherwise leave 0.7-0.85 ms of gaps between kernels.
static bool graphs_enabled() {
    static int on = -1;
    if (on < 0) {
        const char* e = getenv("SIGGAN_GRAPH");
        on = (e && e[0] == '0') ? 0 : 1;
    }
    return on != 0;
}

static bool whole_step_graph_enabled() {
    static int on = -1;
    if (on < 0) {
        const char* e = getenv("SIGGAN_STEP_GRAPH");
        on = (e && e[0] == '0') ? 0 : 1;
    }
    return on != 0;
}

template <typename T>
static int cast_latents_t(sg_ctx* c, const float* z, void* ws, int B, cudaStream_t s) {
    GWs w = carve_g(c, ws, B);
    PROF("g.cast_z", 0, B * (4.0 * c->cfg.latent_dim + (double)c->es * c->Kp));
    sg::cast_pad_z<T>(z, reinterpret_cast<T*>(w.zp), B, c->cfg.latent_dim, c->Kp, s);
    return 0;
}

static int train_phase_body(sg_ctx* c, const sg_train_state* st, int B, float* d_grads, float* g_grads, float* metrics,
                            int phase, bool masks_injected, cudaStream_t s) {
    const size_t img = (size_t)B * c->S * c->S;
    float* x2 = static_cast<float*>(c->x2.p);
    const bool dropout = st->dropout_p > 0.f;
    const float gscale = st->grad_scale > 0.f ? st->grad_scale : 1.f;
    if (phase == 12) {
        const float* masks = dropout ? static_cast<const float*>(c->masks2.p) : nullptr;
        SG_TRY(DISPATCH_T(c, d_backward_t, c, st->d_params, x2, c->d_ws.p, masks, c->dlogit, 2 * B, d_grads, nullptr, s, 2));
    }
    if (phase == 1 || phase == 11) {
        // ---- D step (train…:281-337): D.train(), G.eval(); one 2B batch [real | G(noise)] since D has no batch coupling
        // The Discriminator's weight packs and dropout masks do not depend on the generated half of the batch: they run
        // on the side branch while the Generator produces it
        const float* masks = nullptr;
        bool packed_aside = false;
        {
            SideReduce branch(c, s);  // (joins the branch back at the end of this scope)
            cudaStream_t sb = sg::wgrad_side_fork(s);
            if (dropout) {
                float* m2 = static_cast<float*>(c->masks2.p);
                if (!masks_injected)
                    sg::dropout_masks_dev(st->seed, &c->counters->dropout_offset, sg_d_mask_count(c, 2 * B), st->dropout_p, m2,
                                          sb);
                masks = m2;
            }
            if (sb != s && !(c->skip_d_pack && c->d_pack_src == st->d_params)) {
                SG_TRY(pack_discriminator(c, st->d_params, sb));
                packed_aside = true;
            }
            sg::wgrad_side_mark();
            SG_TRY(DISPATCH_T(c, g_forward_t, c, st->g_params, st->g_running_stats, nullptr, B, 0, c->gws_tmp.p, x2 + img,
                              nullptr, false, s));
        }
        const bool keep = c->skip_d_pack;
        if (packed_aside) c->skip_d_pack = true;  // d_forward_t finds the packs it needs
        const int rc_fwd = DISPATCH_T(c, d_forward_t, c, st->d_params, x2, 2 * B, masks, c->d_ws.p, nullptr, nullptr, s);
        c->skip_d_pack = keep;
        SG_TRY(rc_fwd);
        DWs w = carve_d(c, c->d_ws.p, 2 * B);
        sg::d_loss_metrics(w.prob, B, st->label_smoothing, metrics, c->dlogit, s);
        SG_TRY(DISPATCH_T(c, d_backward_t, c, st->d_params, x2, c->d_ws.p, masks, c->dlogit, 2 * B, d_grads, nullptr, s,
                          phase == 11 ? 1 : 0));
    }
    if (phase == 2 || phase == 22) {
        PROF("adam.d", 0, 28.0 * c->d_count);
        sg::step_prep(c->counters, 1, st->d_lr, st->beta1, st->beta2,
                      dropout ? (unsigned long long)sg_d_mask_count(c, 2 * B) : 0ull, s);
        bool done = false;
        if (phase == 22) {  // the optimizer emits the packed copies the G step's D forward / backward will read (K10)
            sg::PackPlan plan;
            done = discriminator_pack_plan(c, st->d_params, plan) == 0 &&
                   sg::adam_pack_step_dev(plan, st->d_params, d_grads, st->d_exp_avg, st->d_exp_avg_sq, c->d_count, st->beta1,
                                          st->beta2, st->eps, c->counters->adam_d, gscale, s) == 0;
            if (done) c->d_pack_src = st->d_params;
        }
        if (!done) {
            if (phase == 22) return fail("sg_train_step: the Discriminator's pack plan does not tile its parameter buffer");
            sg::adam_step_dev(st->d_params, d_grads, st->d_exp_avg, st->d_exp_avg_sq, c->d_count, st->beta1, st->beta2,
                              st->eps, c->counters->adam_d, gscale, s);
        }
    }
    if (phase == 32) {
        SG_TRY(DISPATCH_T(c, g_backward_t, c, st->g_params, c->g_ws.p, static_cast<const float*>(c->dximg.p), B, 1, g_grads,
                          nullptr, s, kAllLevels, nullptr, nullptr, 2));
    }
    if (phase == 3 || phase == 31) {
        // ---- G step (train…:339-376): G.train() (batch-stat BN), D.eval() (no dropout), labels = 1
        SG_TRY(DISPATCH_T(c, g_forward_t, c, st->g_params, st->g_running_stats, nullptr, B, 1, c->g_ws.p, nullptr, nullptr,
                          true, s));
        const float* fake = carve_g(c, c->g_ws.p, B).out;  // D reads the generated images where G left them
        SG_TRY(DISPATCH_T(c, d_forward_t, c, st->d_params, fake, B, nullptr, c->d_ws.p, nullptr, nullptr, s));
        DWs w = carve_d(c, c->d_ws.p, B);
        sg::g_loss_metrics(w.prob, B, metrics, c->dlogit, s);
        float* dximg = static_cast<float*>(c->dximg.p);
        // weight gradients of D are not needed here (the reference's autograd computes and discards them)
        SG_TRY(DISPATCH_T(c, d_backward_t, c, st->d_params, fake, c->d_ws.p, nullptr, c->dlogit, B, nullptr, dximg, s));
        SG_TRY(DISPATCH_T(c, g_backward_t, c, st->g_params, c->g_ws.p, dximg, B, 1, g_grads, nullptr, s, kAllLevels, nullptr,
                          nullptr, phase == 31 ? 1 : 0));
    }
    if (phase == 4) {
        PROF("adam.g", 0, 28.0 * c->g_count);
        sg::step_prep(c->counters, 0, st->g_lr, st->beta1, st->beta2, 0ull, s);
        sg::adam_step_dev(st->g_params, g_grads, st->g_exp_avg, st->g_exp_avg_sq, c->g_count, st->beta1, st->beta2, st->eps,
                          c->counters->adam_g, gscale, s);
    }
    SG_KCHECK("sg_train_step");
    return 0;
}

// Per-step inputs and counters of one phase: plain launches on the caller's stream, outside any graph.
static int stage_phase_inputs(sg_ctx* c, sg_train_state* st, const float* real, const float* noise_d, const float* noise_g,
                              int B, float* d_grads, float* g_grads, int phase, cudaStream_t s) {
    const size_t img = (size_t)B * c->S * c->S;
    const bool dropout = st->dropout_p > 0.f;
    const bool masks_injected = dropout && st->masks_real && st->masks_fake;
    if (phase == 1 || phase == 11) {
        if (!real || !noise_d || !d_grads) return fail("sg_train_step: D phase needs real, noise_d and d_grads");
        float* x2 = static_cast<float*>(c->x2.p);
        cudaError_t e = cudaMemcpyAsync(x2, real, img * 4, cudaMemcpyDeviceToDevice, s);
        if (e != cudaSuccess) return fail("sg_train_step: copy of real batch failed: %s", cudaGetErrorString(e));
        SG_TRY(DISPATCH_T(c, cast_latents_t, c, noise_d, c->gws_tmp.p, B, s));
        if (dropout) {
            float* m2 = static_cast<float*>(c->masks2.p);
            if (masks_injected) {
                for (int i = 0; i < c->ND; ++i) {
                    const size_t n = (size_t)B * c->dch[i + 1];
                    cudaMemcpyAsync(m2 + mask_offset(c, 2 * B, i), st->masks_real + mask_offset(c, B, i), n * 4,
                                    cudaMemcpyDeviceToDevice, s);
                    cudaMemcpyAsync(m2 + mask_offset(c, 2 * B, i) + n, st->masks_fake + mask_offset(c, B, i), n * 4,
                                    cudaMemcpyDeviceToDevice, s);
                }
            } else if (!c->mirror_drop_known || c->mirror_drop != st->offset) {
                sg::step_set(c->counters, 2, (long long)st->offset, s);
                c->mirror_drop = st->offset;
                c->mirror_drop_known = true;
            }
        }
    } else if (phase == 12) {
        if (!d_grads) return fail("sg_train_step: D phase needs d_grads");
    } else if (phase == 2 || phase == 22) {
        if (!d_grads) return fail("sg_train_step: D update needs d_grads");
        if (c->mirror_d_step != st->d_step) sg::step_set(c->counters, 1, st->d_step, s);
        c->mirror_d_step = st->d_step + 1;
        if (dropout && c->mirror_drop_known) c->mirror_drop += (unsigned long long)sg_d_mask_count(c, 2 * B);
    } else if (phase == 3 || phase == 31) {
        if (!noise_g || !g_grads) return fail("sg_train_step: G phase needs noise_g and g_grads");
        SG_TRY(DISPATCH_T(c, cast_latents_t, c, noise_g, c->g_ws.p, B, s));
    } else if (phase == 32) {
        if (!g_grads) return fail("sg_train_step: G phase needs g_grads");
    } else if (phase == 4) {
        if (!g_grads) return fail("sg_train_step: G update needs g_grads");
        if (c->mirror_g_step != st->g_step) sg::step_set(c->counters, 0, st->g_step, s);
        c->mirror_g_step = st->g_step + 1;
    } else {
        return fail("sg_train_step: unknown phase %d", phase);
    }
    SG_KCHECK("sg_train_step(inputs)");
    return 0;
}

// kWholeStep: phases 1, 22, 3, 4 of one single-process step as ONE body (one graph): the staged inputs of all four are
// in place before it starts (the G step's latents go to their own workspace), and the G-step part reuses the packs.
constexpr int kWholeStep = 100;
static int phase_bodies(sg_ctx* c, const sg_train_state* st, int B, float* d_grads, float* g_grads, float* metrics,
                        int phase, bool masks_injected, cudaStream_t s) {
    if (phase != kWholeStep) return train_phase_body(c, st, B, d_grads, g_grads, metrics, phase, masks_injected, s);
    SG_TRY(train_phase_body(c, st, B, d_grads, g_grads, metrics, 1, masks_injected, s));
    SG_TRY(train_phase_body(c, st, B, d_grads, g_grads, metrics, 22, masks_injected, s));
    c->skip_g_pack = c->skip_d_pack = true;
    const int rc = train_phase_body(c, st, B, d_grads, g_grads, metrics, 3, masks_injected, s);
    c->skip_g_pack = c->skip_d_pack = false;
    if (rc != 0) return rc;
    return train_phase_body(c, st, B, d_grads, g_grads, metrics, 4, masks_injected, s);
}

static int train_phase(sg_ctx* c, sg_train_state* st, const float* real, const float* noise_d, const float* noise_g,
                       int B, float* d_grads, float* g_grads, float* metrics, int phase, cudaStream_t s) {
    const bool dropout = st->dropout_p > 0.f;
    const bool masks_injected = dropout && st->masks_real && st->masks_fake;
    if (phase == kWholeStep) {
        static const int seq[4] = {1, 22, 3, 4};
        for (int ph : seq) SG_TRY(stage_phase_inputs(c, st, real, noise_d, noise_g, B, d_grads, g_grads, ph, s));
    } else {
        SG_TRY(stage_phase_inputs(c, st, real, noise_d, noise_g, B, d_grads, g_grads, phase, s));
    }
    const bool use_graph = graphs_enabled() && !c->prof.on && !sync_bn_on(c);
    if (!use_graph) return phase_bodies(c, st, B, d_grads, g_grads, metrics, phase, masks_injected, s);

    GraphKey key;
    memset(&key, 0, sizeof(key));
    key.phase = phase;
    key.B = B;
    key.flags = (masks_injected ? 1 : 0) | (dropout ? 2 : 0) | (c->skip_g_pack ? 4 : 0) | (c->skip_d_pack ? 8 : 0);
    const void* ptrs[10] = {st->g_params, st->g_running_stats, st->g_exp_avg, st->g_exp_avg_sq, st->d_params,
                            st->d_exp_avg, st->d_exp_avg_sq, d_grads, g_grads, metrics};
    memcpy(key.ptr, ptrs, sizeof(ptrs));
    const float fl[8] = {st->g_lr, st->d_lr, st->beta1, st->beta2, st->eps, st->label_smoothing, st->dropout_p, st->grad_scale};
    memcpy(key.f, fl, sizeof(fl));
    key.seed = st->seed;
    GraphEntry* ent = nullptr;
    for (GraphEntry& g : c->graphs)
        if (g.key == key) ent = &g;
    if (!ent) {
        if (c->graphs.size() >= 24) {  // evict the least recently used entry
            size_t lru = 0;
            for (size_t i = 1; i < c->graphs.size(); ++i)
                if (c->graphs[i].last < c->graphs[lru].last) lru = i;
            if (c->graphs[lru].exec) cudaGraphExecDestroy(c->graphs[lru].exec);
            c->graphs.erase(c->graphs.begin() + lru);
        }
        c->graphs.emplace_back();
        ent = &c->graphs.back();
        ent->key = key;
    }
    ent->last = ++c->graph_tick;
    if (ent->exec && ent->epoch != g_alloc_epoch) {  // a library buffer moved since the capture
        cudaGraphExecDestroy(ent->exec);
        ent->exec = nullptr;
        ent->seen = 0;
    }
    if (ent->bad || ent->seen == 0) {
        // first sight of this configuration: run it eagerly (loads modules, sets kernel attributes, sizes buffers)
        ent->seen = 1;
        return phase_bodies(c, st, B, d_grads, g_grads, metrics, phase, masks_injected, s);
    }
    if (!ent->exec) {
        if (!c->cap_stream && cudaStreamCreateWithFlags(&c->cap_stream, cudaStreamNonBlocking) != cudaSuccess)
            return fail("sg_train_step: cannot create the capture stream");
        cudaError_t e = cudaStreamBeginCapture(c->cap_stream, cudaStreamCaptureModeThreadLocal);
        if (e != cudaSuccess) {
            cudaGetLastError();
            ent->bad = true;
            return phase_bodies(c, st, B, d_grads, g_grads, metrics, phase, masks_injected, s);
        }
        const unsigned long long launches0 = sg::g_launches;
        const int rc = phase_bodies(c, st, B, d_grads, g_grads, metrics, phase, masks_injected, c->cap_stream);
        ent->n_launches = static_cast<int>(sg::g_launches - launches0);
        sg::g_launches = launches0;  // counted per graph launch below
        cudaGraph_t graph = nullptr;
        e = cudaStreamEndCapture(c->cap_stream, &graph);
        if (rc != 0 || e != cudaSuccess || !graph) {
            if (graph) cudaGraphDestroy(graph);
            cudaGetLastError();
            ent->bad = true;
            if (rc != 0) return rc;
            return phase_bodies(c, st, B, d_grads, g_grads, metrics, phase, masks_injected, s);
        }
        e = cudaGraphInstantiate(&ent->exec, graph, 0);
        cudaGraphDestroy(graph);
        if (e != cudaSuccess) {
            cudaGetLastError();
            ent->exec = nullptr;
            ent->bad = true;
            return phase_bodies(c, st, B, d_grads, g_grads, metrics, phase, masks_injected, s);
        }
        ent->epoch = g_alloc_epoch;
    }
    cudaError_t e = cudaGraphLaunch(ent->exec, s);
    if (e != cudaSuccess) return fail("sg_train_step: cudaGraphLaunch failed: %s", cudaGetErrorString(e));
    sg::g_launches += static_cast<unsigned long long>(ent->n_launches);
    return 0;
}

// Scope of the G-step phase inside one phase-0 call: the packs written earlier in the same call are current.
struct StepPackReuse {
    sg_ctx* c;
    explicit StepPackReuse(sg_ctx* ctx) : c(ctx) { c->skip_g_pack = c->skip_d_pack = true; }
    ~StepPackReuse() { c->skip_g_pack = c->skip_d_pack = false; }
};

// ================================================================================================
// extern "C"
// ================================================================================================
extern "C" {

int sg_abi_version(void) { return SG_ABI_VERSION; }
unsigned long long sg_launch_count(void) { return sg::g_launches; }

int sg_profile_enable(sg_ctx* c, int on) {
    if (!c) return fail("sg_profile_enable: null ctx");
    for (ProfRec& r : c->prof.recs) {
        c->prof.pool.push_back(r.a);
        c->prof.pool.push_back(r.b);
    }
    c->prof.recs.clear();
    c->prof.on = on != 0;
    return 0;
}

// Writes one line per recorded op: "name\tmilliseconds\tflops\tbytes\n". Synchronises the device.
long long sg_profile_dump(sg_ctx* c, char* buf, size_t cap) {
    if (!c || !buf || cap == 0) return fail("sg_profile_dump: bad argument");
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) return fail("sg_profile_dump: %s", cudaGetErrorString(e));
    size_t off = 0;
    buf[0] = 0;
    for (const ProfRec& r : c->prof.recs) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, r.a, r.b);
        int n = snprintf(buf + off, cap - off, "%s\t%.6f\t%.6e\t%.6e\n", r.name.c_str(), ms, r.flops, r.bytes);
        if (n < 0 || static_cast<size_t>(n) >= cap - off) break;
        off += n;
    }
    return static_cast<long long>(off);
}
const char* sg_last_error(void) { return t_err; }

int sg_create(const sg_config* cfg, sg_ctx** out) {
    if (!cfg || !out) return fail("sg_create: null argument");
    if (cfg->image_size != 64 && cfg->image_size != 128)
        return fail("sg_create: image_size must be 64 or 128, got %d", cfg->image_size);
    if (cfg->latent_dim < 1 || cfg->latent_dim > 512) return fail("sg_create: latent_dim %d unsupported", cfg->latent_dim);
    if (cfg->precision != SG_PREC_BF16 && cfg->precision != SG_PREC_FP32) return fail("sg_create: bad precision");
    if (!(cfg->g_act_slope >= 0.f && cfg->g_act_slope < 1.f)) return fail("sg_create: g_act_slope must be in [0, 1)");
    const int wm = cfg->width_mult == 0 ? 1 : cfg->width_mult;
    if (wm != 1 && wm != 2) return fail("sg_create: width_mult must be 1 or 2, got %d", cfg->width_mult);
    if (wm == 2 && cfg->precision != SG_PREC_BF16)
        return fail("sg_create: the 2x-width variant runs in the bf16 tensor-core mode only (its first conv / final conv "
                    "kernels in the fp32 validation mode are specialised to the reference's widths)");
    if (wm == 2 && cfg->g_act_slope != 0.f) return fail("sg_create: the 2x-width variant supports the ReLU generator only");
    int dev_count = 0;
    if (cudaGetDeviceCount(&dev_count) != cudaSuccess || dev_count == 0) {
        cudaGetLastError();
        return fail("sg_create: no CUDA device (this library has no CPU path)");
    }
    sg_ctx* c = new sg_ctx();
    c->cfg = *cfg;
    cudaGetDevice(&c->device);
    c->S = cfg->image_size;
    c->es = cfg->precision == SG_PREC_BF16 ? 2 : 4;
    c->Kp = (cfg->latent_dim + 63) / 64 * 64;
    if (c->S == 64) {
        const int g[] = {256, 128, 64, 32, 32}, d[] = {1, 64, 128, 256, 512};
        c->L = 4;
        c->ND = 4;
        memcpy(c->gch, g, sizeof(g));
        memcpy(c->dch, d, sizeof(d));
    } else {
        const int g[] = {512, 256, 128, 64, 32, 32}, d[] = {1, 64, 128, 256, 512, 512};
        c->L = 5;
        c->ND = 5;
        memcpy(c->gch, g, sizeof(g));
        memcpy(c->dch, d, sizeof(d));
    }
    // 2x hidden width (BASELINE configs[4]; not a configuration of the reference — the same blocks, doubled ladders)
    for (int i = 0; i <= c->L; ++i) c->gch[i] *= wm;
    for (int i = 1; i <= c->ND; ++i) c->dch[i] *= wm;
    // ---- parameter tables in reference named_parameters() order (gen…:124-163, disc…:131-207)
    long long off = 0, soff = 0;
    const int F0 = c->gch[0] * 16;
    c->g_fc_w = (int)c->gt.size();
    add_tensor(c->gt, off, "fc.0.weight", F0, cfg->latent_dim);
    c->g_fc_b = (int)c->gt.size();
    add_tensor(c->gt, off, "fc.0.bias", F0);
    c->bn[0].C = F0;
    c->bn[0].gamma_off = off;
    add_tensor(c->gt, off, "fc.1.weight", F0);
    c->bn[0].beta_off = off;
    add_tensor(c->gt, off, "fc.1.bias", F0);
    c->bn[0].mean_off = soff;
    c->bn[0].var_off = soff + F0;
    soff += 2 * F0;
    for (int i = 0; i < c->L; ++i) {
        const std::string p = "upsample_blocks." + std::to_string(i) + ".block.";
        c->g_up_w[i] = (int)c->gt.size();
        add_tensor(c->gt, off, p + "0.weight", c->gch[i], c->gch[i + 1], 4, 4);
        c->bn[i + 1].C = c->gch[i + 1];
        c->bn[i + 1].gamma_off = off;
        add_tensor(c->gt, off, p + "1.weight", c->gch[i + 1]);
        c->bn[i + 1].beta_off = off;
        add_tensor(c->gt, off, p + "1.bias", c->gch[i + 1]);
        c->bn[i + 1].mean_off = soff;
        c->bn[i + 1].var_off = soff + c->gch[i + 1];
        soff += 2 * c->gch[i + 1];
    }
    c->g_final_w = (int)c->gt.size();
    add_tensor(c->gt, off, "final_conv.0.weight", 1, c->gch[c->L], 3, 3);
    c->g_final_b = (int)c->gt.size();
    add_tensor(c->gt, off, "final_conv.0.bias", 1);
    c->g_count = off;
    c->g_stats = soff;
    off = 0;
    for (int i = 0; i < c->ND; ++i) {
        const std::string p = "conv_blocks." + std::to_string(i) + ".block.0.";
        c->d_conv_w[i] = (int)c->dt.size();
        add_tensor(c->dt, off, p + "weight", c->dch[i + 1], c->dch[i], 4, 4);
        c->d_conv_b[i] = (int)c->dt.size();
        add_tensor(c->dt, off, p + "bias", c->dch[i + 1]);
    }
    c->d_cls_w = (int)c->dt.size();
    add_tensor(c->dt, off, "classifier.0.weight", 1, c->dch[c->ND] * 16);
    c->d_cls_b = (int)c->dt.size();
    add_tensor(c->dt, off, "classifier.0.bias", 1);
    c->d_count = off;
    // ---- weight packs
    size_t pbytes = 0;
    auto reserve = [&](size_t bytes) {
        size_t o = pbytes;
        pbytes += align_up(bytes);
        return o;
    };
    size_t o_fcW = reserve((size_t)F0 * c->Kp * 2), o_fcb = reserve((size_t)F0 * 4);
    size_t o_gF[6], o_gB[6], o_dF[6], o_dB[6];
    for (int i = 0; i < c->L; ++i) {
        o_gF[i] = reserve((size_t)c->gch[i] * c->gch[i + 1] * 32);
        o_gB[i] = reserve((size_t)c->gch[i] * c->gch[i + 1] * 32);
    }
    for (int i = 1; i < c->ND; ++i) {
        o_dF[i] = reserve((size_t)c->dch[i] * c->dch[i + 1] * 32);
        o_dB[i] = reserve((size_t)c->dch[i] * c->dch[i + 1] * 32);
    }
    size_t o_cls = reserve((size_t)c->dch[c->ND] * 16 * 4);
    if (c->packs.ensure(pbytes) != 0) {
        delete c;
        return -1;
    }
    char* pb = static_cast<char*>(c->packs.p);
    c->fc_Wp = reinterpret_cast<bf16*>(pb + o_fcW);
    c->fc_biasp = reinterpret_cast<float*>(pb + o_fcb);
    for (int i = 0; i < c->L; ++i) {
        c->g_packF[i] = reinterpret_cast<bf16*>(pb + o_gF[i]);
        c->g_packB[i] = reinterpret_cast<bf16*>(pb + o_gB[i]);
    }
    for (int i = 1; i < c->ND; ++i) {
        c->d_packF[i] = reinterpret_cast<bf16*>(pb + o_dF[i]);
        c->d_packB[i] = reinterpret_cast<bf16*>(pb + o_dB[i]);
    }
    c->cls_wp = reinterpret_cast<float*>(pb + o_cls);
    *out = c;
    return 0;
}

void sg_destroy(sg_ctx* c) {
    if (!c) return;
    DeviceGuard guard(c->device);
    for (GraphEntry& g : c->graphs)
        if (g.exec) cudaGraphExecDestroy(g.exec);
    if (c->cap_stream) cudaStreamDestroy(c->cap_stream);
    if (c->red_stream) cudaStreamDestroy(c->red_stream);
    if (c->red_fork) cudaEventDestroy(c->red_fork);
    if (c->red_join) cudaEventDestroy(c->red_join);
    sg::comm_release(&c->comm);
    c->counters_buf.release();
    DevBuf* bufs[] = {&c->packs, &c->bufA, &c->bufB, &c->dpre, &c->wpart, &c->cpart, &c->small,
                      &c->g_ws,  &c->d_ws, &c->x2,   &c->masks2, &c->dximg, &c->gws_tmp};
    for (DevBuf* b : bufs) b->release();
    delete c;
}

int sg_num_tensors(const sg_ctx* c, int net) { return (int)(net == SG_NET_G ? c->gt.size() : c->dt.size()); }
int sg_tensor_info(const sg_ctx* c, int net, int index, const char** name, long long* offset, int shape[4]) {
    const std::vector<TensorInfo>& v = net == SG_NET_G ? c->gt : c->dt;
    if (index < 0 || index >= (int)v.size()) return fail("sg_tensor_info: index %d out of range", index);
    if (name) *name = v[index].name.c_str();
    if (offset) *offset = v[index].offset;
    if (shape) memcpy(shape, v[index].shape, sizeof(int) * 4);
    return 0;
}
long long sg_param_count(const sg_ctx* c, int net) { return net == SG_NET_G ? c->g_count : c->d_count; }
long long sg_g_stat_count(const sg_ctx* c) { return c->g_stats; }
int sg_g_num_bn(const sg_ctx* c) { return c->L + 1; }
int sg_g_bn_info(const sg_ctx* c, int index, long long* mean_offset, long long* var_offset, int* channels) {
    if (index < 0 || index > c->L) return fail("sg_g_bn_info: index %d out of range", index);
    *mean_offset = c->bn[index].mean_off;
    *var_offset = c->bn[index].var_off;
    *channels = c->bn[index].C;
    return 0;
}
size_t sg_g_workspace_bytes(const sg_ctx* c, int batch) { return carve_g(c, nullptr, batch).bytes; }
size_t sg_d_workspace_bytes(const sg_ctx* c, int batch) { return carve_d(c, nullptr, batch).bytes; }
long long sg_d_mask_count(const sg_ctx* c, int batch) { return mask_offset(c, batch, c->ND); }
long long sg_d_grad_tail_offset(const sg_ctx* c) { return c->dt[c->d_conv_w[c->ND - 1]].offset; }
long long sg_g_grad_tail_offset(const sg_ctx* c) { return c->gt[c->g_up_w[0]].offset; }
long long sg_d_feature_count(const sg_ctx* c) { return (long long)c->dch[c->ND] * 16; }

// ---- library-owned NCCL communicator ------------------------------------------------------------
int sg_comm_nccl_version(void) { return sg::comm_version(); }
int sg_comm_unique_id(void* host_id_out, size_t cap) {
    if (sg::comm_unique_id(host_id_out, cap) != 0) return fail("%s", sg::comm_last_error());
    return 0;
}
int sg_comm_init(sg_ctx* c, const void* host_id, size_t id_bytes, int rank, int world_size) {
    if (!c) return fail("sg_comm_init: null ctx");
    DeviceGuard guard(c->device);
    if (c->comm.nccl) sg::comm_release(&c->comm);
    if (sg::comm_create(&c->comm, host_id, id_bytes, rank, world_size) != 0) {
        sg::comm_release(&c->comm);
        return fail("%s", sg::comm_last_error());
    }
    return 0;
}
int sg_comm_destroy(sg_ctx* c) {
    if (!c) return fail("sg_comm_destroy: null ctx");
    DeviceGuard guard(c->device);
    sg::comm_release(&c->comm);
    return 0;
}
int sg_comm_world_size(const sg_ctx* c) { return (c && c->comm.nccl) ? c->comm.world : 0; }
int sg_allreduce_grads(sg_ctx* c, int which, float* grads, long long offset, long long count, int async, void* stream) {
    if (!c || !grads || offset < 0 || (which != SG_NET_G && which != SG_NET_D))
        return fail("sg_allreduce_grads: bad argument");
    const long long total = which == SG_NET_G ? c->g_count : c->d_count;
    if (count < 0) count = total - offset;
    if (offset + count > total) return fail("sg_allreduce_grads: [%lld, %lld) exceeds the bucket (%lld)", offset, offset + count, total);
    DeviceGuard guard(c->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (async) SG_COMM(sg::comm_all_reduce_mean_start(&c->comm, grads + offset, count, s));
    else SG_COMM(sg::comm_all_reduce_mean(&c->comm, grads + offset, count, s));
    return 0;
}
int sg_allreduce_join(sg_ctx* c, void* stream) {
    if (!c) return fail("sg_allreduce_join: null ctx");
    DeviceGuard guard(c->device);
    SG_COMM(sg::comm_join(&c->comm, static_cast<cudaStream_t>(stream)));
    return 0;
}

long long sg_ws_offset(const sg_ctx* c, int net, int batch, int kind, int index) {
    if (!c || batch < 1) return -1;
    char base[1];  // carve relative to a dummy base: only differences are used
    if (net == SG_NET_G) {
        GWs w = carve_g(c, base, batch);
        if (index < 0 || index > c->L) return -1;
        const char* p = nullptr;
        switch (kind) {
            case 0: p = w.zp; break;
            case 1: p = w.fc_y; break;
            case 2: p = w.fc_a; break;
            case 3: p = index < c->L ? w.y[index] : nullptr; break;
            case 4: p = index < c->L ? w.a[index] : nullptr; break;
            case 5: p = reinterpret_cast<const char*>(w.out); break;
            case 6: p = reinterpret_cast<const char*>(w.mean[index]); break;
            case 7: p = reinterpret_cast<const char*>(w.rstd[index]); break;
            case 8: p = reinterpret_cast<const char*>(w.scale[index]); break;
            case 9: p = reinterpret_cast<const char*>(w.shift[index]); break;
            default: break;
        }
        return p ? static_cast<long long>(p - base) : -1;
    }
    DWs w = carve_d(c, base, batch);
    if (kind == 0 && index >= 0 && index < c->ND) return static_cast<long long>(w.a[index] - base);
    if (kind == 1) return static_cast<long long>(reinterpret_cast<const char*>(w.prob) - base);
    return -1;
}

int sg_g_forward(sg_ctx* c, const float* params, float* stats, const float* z, int batch, int bn_batch_stats, void* ws,
                 float* out_image, uint8_t* out_u8, void* stream) {
    if (!c || !params || !stats || !z || batch < 1) return fail("sg_g_forward: bad argument");
    if (!ws && !out_image && !out_u8) return fail("sg_g_forward: no output requested");
    if (bn_batch_stats && batch < 2)
        return fail("sg_g_forward: training-mode BatchNorm needs more than 1 value per channel (batch=%d)", batch);
    DeviceGuard guard(c->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    SG_TRY(ensure_scratch(c, batch));
    bool save = ws != nullptr;
    c->u8_only = !ws && !out_image && out_u8;
    if (!ws) {
        SG_TRY(c->gws_tmp.ensure(sg_g_workspace_bytes(c, batch)));
        ws = c->gws_tmp.p;
        if (!out_image) {  // u8-only sampling still needs somewhere to put the fp32 image
            SG_TRY(c->dximg.ensure((size_t)batch * c->S * c->S * 4));
            out_image = static_cast<float*>(c->dximg.p);
        }
    }
    return DISPATCH_T(c, g_forward_t, c, params, stats, z, batch, bn_batch_stats, ws, out_image, out_u8, save, s);
}

int sg_g_backward(sg_ctx* c, const float* params, const void* ws, const float* grad_image, int batch,
                  int bn_batch_stats, float* grads_out, float* dz_out, void* stream) {
    if (!c || !params || !ws || !grad_image || !grads_out || batch < 1) return fail("sg_g_backward: bad argument");
    DeviceGuard guard(c->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    SG_TRY(ensure_scratch(c, batch));
    // the packs are context-global and belong to the latest forward: another Generator sharing this context (a second
    // VanillaGAN, an EMA copy, an ablation model) may have run in between
    if (c->g_pack_src != params) SG_TRY(pack_generator(c, params, s));
    return DISPATCH_T(c, g_backward_t, c, params, ws, grad_image, batch, bn_batch_stats, grads_out, dz_out, s);
}

int sg_d_forward(sg_ctx* c, const float* params, const float* x, int batch, const float* masks, void* ws,
                 float* prob_out, float* features_out, void* stream) {
    if (!c || !params || !x || batch < 1) return fail("sg_d_forward: bad argument");
    DeviceGuard guard(c->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    SG_TRY(ensure_scratch(c, batch));
    if (!ws) {
        SG_TRY(c->d_ws.ensure(sg_d_workspace_bytes(c, batch)));
        ws = c->d_ws.p;
    }
    return DISPATCH_T(c, d_forward_t, c, params, x, batch, masks, ws, prob_out, features_out, s);
}

int sg_d_backward(sg_ctx* c, const float* params, const float* x, const void* ws, const float* masks,
                  const float* grad_prob, int batch, float* grads_out, float* dx_out, void* stream) {
    if (!c || !params || !x || !ws || !grad_prob || batch < 1) return fail("sg_d_backward: bad argument");
    DeviceGuard guard(c->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    SG_TRY(ensure_scratch(c, batch));
    DWs w = carve_d(c, const_cast<void*>(ws), batch);
    sg::sigmoid_bwd(w.prob, grad_prob, c->dlogit, batch, s);
    // the packs are those of the latest forward: a backward for an earlier forward that ran on OTHER weights (the
    // spectral-norm variant hands every forward its own weight_orig / sigma buffer) packs its own again
    if (c->d_pack_src != params) SG_TRY(pack_discriminator(c, params, s));
    return DISPATCH_T(c, d_backward_t, c, params, x, ws, masks, c->dlogit, batch, grads_out, dx_out, s);
}

int sg_d_backward_layer(sg_ctx* c, const float* params, const float* x, const void* ws, const float* masks, int layer,
                        const void* dz_in, int batch, float* grads_out, void* dz_prev_out, float* dx_out, void* stream) {
    if (!c || !params || !x || !ws || !dz_in || !grads_out || batch < 1 || layer < 0 || layer > c->ND)
        return fail("sg_d_backward_layer: bad argument");
    DeviceGuard guard(c->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    SG_TRY(ensure_scratch(c, batch));
    const float* dlogit = nullptr;
    if (layer == c->ND) {  // dz_in = d(loss)/d(probability), as for sg_d_backward
        DWs w = carve_d(c, const_cast<void*>(ws), batch);
        sg::sigmoid_bwd(w.prob, static_cast<const float*>(dz_in), c->dlogit, batch, s);
        dlogit = c->dlogit;
    }
    if (c->d_pack_src != params) SG_TRY(pack_discriminator(c, params, s));
    return DISPATCH_T(c, d_backward_t, c, params, x, ws, masks, dlogit, batch, grads_out, dx_out, s, 0, layer, dz_in,
                      dz_prev_out);
}

int sg_g_backward_layer(sg_ctx* c, const float* params, const void* ws, int level, const void* d_in, int batch,
                        int bn_batch_stats, float* grads_out, void* d_prev_out, void* stream) {
    if (!c || !params || !ws || !d_in || !grads_out || batch < 1 || level < -1 || level >= c->L)
        return fail("sg_g_backward_layer: bad argument");
    DeviceGuard guard(c->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    SG_TRY(ensure_scratch(c, batch));
    if (c->g_pack_src != params) SG_TRY(pack_generator(c, params, s));
    const bool top = level == c->L - 1;
    return DISPATCH_T(c, g_backward_t, c, params, ws, top ? static_cast<const float*>(d_in) : nullptr, batch,
                      bn_batch_stats, grads_out, nullptr, s, level, top ? nullptr : d_in, d_prev_out);
}

int sg_dropout_masks(sg_ctx* c, uint64_t seed, uint64_t offset, int batch, float p, float* masks_out, void* stream) {
    if (!c || !masks_out || batch < 1 || !(p >= 0.f && p < 1.f)) return fail("sg_dropout_masks: bad argument");
    DeviceGuard guard(c->device);
    sg::dropout_masks(seed, offset, sg_d_mask_count(c, batch), p, masks_out, static_cast<cudaStream_t>(stream));
    SG_KCHECK("sg_dropout_masks");
    return 0;
}

int sg_bce_forward(const float* prob, const float* target, int n, float* loss_out, void* stream) {
    PtrDeviceGuard guard(prob);
    if (!prob || !target || !loss_out || n < 1) return fail("sg_bce_forward: bad argument");
    sg::bce_forward(prob, target, n, loss_out, static_cast<cudaStream_t>(stream));
    SG_KCHECK("sg_bce_forward");
    return 0;
}
int sg_bce_backward(const float* prob, const float* target, int n, const float* grad_loss, float* dprob_out,
                    void* stream) {
    PtrDeviceGuard guard(prob);
    if (!prob || !target || !grad_loss || !dprob_out || n < 1) return fail("sg_bce_backward: bad argument");
    sg::bce_backward(prob, target, n, grad_loss, dprob_out, static_cast<cudaStream_t>(stream));
    SG_KCHECK("sg_bce_backward");
    return 0;
}

int sg_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, long long n, float lr,
                 float beta1, float beta2, float eps, long long step, void* stream) {
    PtrDeviceGuard guard(params);
    if (!params || !grads || !exp_avg || !exp_avg_sq || n < 1 || step < 1) return fail("sg_adam_step: bad argument");
    sg::adam_step(params, grads, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, step, static_cast<cudaStream_t>(stream));
    SG_KCHECK("sg_adam_step");
    return 0;
}

int sg_augment_params(const double* host_angles, const double* host_scales, int n, int image_size, int* host_rot_fixed,
                      double* host_scale_affine) {
    if (n < 0 || !host_rot_fixed || !host_scale_affine || (image_size != 64 && image_size != 128))
        return fail("sg_augment_params: bad argument");
    sg::augment_params(host_angles, host_scales, n, image_size, host_rot_fixed, host_scale_affine);
    return 0;
}
int sg_augment_batch(const uint8_t* pool, const int* index, const int* rot_fixed, const double* scale_affine,
                     const uint8_t* flip, int batch, int image_size, float* out, void* stream) {
    PtrDeviceGuard guard(pool);
    if (!pool || !rot_fixed || !scale_affine || !out || batch < 1) return fail("sg_augment_batch: bad argument");
    if (image_size != 64 && image_size != 128)
        return fail("sg_augment_batch: image_size must be 64 or 128, got %d", image_size);
    if ((reinterpret_cast<uintptr_t>(pool) & 15) || (reinterpret_cast<uintptr_t>(out) & 15))
        return fail("sg_augment_batch: pool and out must be 16-byte aligned");
    sg::augment_batch(pool, index, rot_fixed, scale_affine, flip, batch, image_size, out, static_cast<cudaStream_t>(stream));
    SG_KCHECK("sg_augment_batch");
    return 0;
}

int sg_ink_stats(const float* images, int n_images, int pixels_per_image, float threshold, int* count_raw,
                 int* count_rescaled, float* minimum, void* stream) {
    PtrDeviceGuard guard(images);
    if (!images || !count_raw || !count_rescaled || !minimum || n_images < 1 || pixels_per_image < 4 ||
        (pixels_per_image & 3) || (reinterpret_cast<uintptr_t>(images) & 15))
        return fail("sg_ink_stats: bad argument (pixels per image must be a multiple of 4, images 16-byte aligned)");
    sg::ink_stats(images, n_images, pixels_per_image, threshold, count_raw, count_rescaled, minimum,
                  static_cast<cudaStream_t>(stream));
    SG_KCHECK("sg_ink_stats");
    return 0;
}

int sg_set_sync_batchnorm(sg_ctx* c, sg_allreduce_fn fn, void* user, int world_size, float* buf, long long buf_floats) {
    if (!c) return fail("sg_set_sync_batchnorm: null ctx");
    if (fn && (world_size < 1 || !buf || buf_floats < 4)) return fail("sg_set_sync_batchnorm: bad argument");
    c->sync_fn = fn;
    c->sync_user = user;
    c->sync_world = fn ? world_size : 1;
    c->sync_buf = fn ? buf : nullptr;
    c->sync_cap = fn ? buf_floats : 0;
    return 0;
}

int sg_spectral_norm_weight(const float* w_orig, float* u, float* v, int rows, int cols, int power_iterations,
                            float eps, float* w_out, float* sigma_out, float* scratch, void* stream) {
    PtrDeviceGuard guard(w_orig);
    if (!w_orig || !u || !v || !w_out || !sigma_out || !scratch || rows < 1 || cols < 1 || power_iterations < 0 ||
        !(eps > 0.f))
        return fail("sg_spectral_norm_weight: bad argument");
    sg::spectral_norm_weight(w_orig, u, v, rows, cols, power_iterations, eps, w_out, sigma_out, scratch,
                             static_cast<cudaStream_t>(stream));
    SG_KCHECK("sg_spectral_norm_weight");
    return 0;
}
int sg_spectral_norm_backward(const float* w_eff, const float* u, const float* v, const float* sigma, int rows,
                              int cols, float* grad, float* scratch, void* stream) {
    PtrDeviceGuard guard(w_eff);
    if (!w_eff || !u || !v || !sigma || !grad || !scratch || rows < 1 || cols < 1)
        return fail("sg_spectral_norm_backward: bad argument");
    sg::spectral_norm_backward(w_eff, u, v, sigma, rows, cols, grad, scratch, static_cast<cudaStream_t>(stream));
    SG_KCHECK("sg_spectral_norm_backward");
    return 0;
}

int sg_train_step(sg_ctx* c, sg_train_state* st, const float* real, const float* noise_d, const float* noise_g,
                  int B, float* d_grads, float* g_grads, float* metrics, int phase, void* stream) {
    if (!c || !st || !metrics || B < 2) return fail("sg_train_step: bad argument");
    DeviceGuard guard(c->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    // ---- library-owned buffers (never grown inside a captured region) --------------------------------------------
    SG_TRY(ensure_scratch(c, 2 * B));
    const size_t img = (size_t)B * c->S * c->S;
    SG_TRY(c->x2.ensure(2 * img * 4));
    SG_TRY(c->d_ws.ensure(sg_d_workspace_bytes(c, 2 * B)));
    SG_TRY(c->g_ws.ensure(sg_g_workspace_bytes(c, B)));
    SG_TRY(c->gws_tmp.ensure(sg_g_workspace_bytes(c, B)));
    SG_TRY(c->dximg.ensure(img * 4));
    if (st->dropout_p > 0.f) SG_TRY(c->masks2.ensure((size_t)sg_d_mask_count(c, 2 * B) * 4));
    if (!c->counters) {
        SG_TRY(c->counters_buf.ensure(sizeof(sg::StepCounters)));
        c->counters = static_cast<sg::StepCounters*>(c->counters_buf.p);
        cudaMemsetAsync(c->counters, 0, sizeof(sg::StepCounters), s);
    }
    if (phase == 0 && c->comm.nccl && c->comm.world > 1) {
        // data-parallel replica: each gradient group is averaged over the ranks on the communication stream as soon as
        // it is final, while the rest of that backward pass runs (D: classifier + last conv block first, 76 % of the
        // bucket; G: the upsample blocks and the final conv first, the fc stage last), and the update waits for both
        const long long d_tail = sg_d_grad_tail_offset(c), g_tail = sg_g_grad_tail_offset(c);
        if (st->grad_scale != 0.f && st->grad_scale != 1.f)
            return fail("sg_train_step: the library's all-reduce averages; grad_scale must be 0 or 1");
        SG_TRY(train_phase(c, st, real, noise_d, noise_g, B, d_grads, g_grads, metrics, 11, s));
        SG_COMM(sg::comm_all_reduce_mean_start(&c->comm, d_grads + d_tail, c->d_count - d_tail, s));
        SG_TRY(train_phase(c, st, real, noise_d, noise_g, B, d_grads, g_grads, metrics, 12, s));
        SG_COMM(sg::comm_all_reduce_mean_start(&c->comm, d_grads, d_tail, s));
        SG_COMM(sg::comm_join(&c->comm, s));
        SG_TRY(train_phase(c, st, real, noise_d, noise_g, B, d_grads, g_grads, metrics, 22, s));
        {
            StepPackReuse reuse(c);
            SG_TRY(train_phase(c, st, real, noise_d, noise_g, B, d_grads, g_grads, metrics, 31, s));
        }
        SG_COMM(sg::comm_all_reduce_mean_start(&c->comm, g_grads + g_tail, c->g_count - g_tail, s));
        SG_TRY(train_phase(c, st, real, noise_d, noise_g, B, d_grads, g_grads, metrics, 32, s));
        SG_COMM(sg::comm_all_reduce_mean_start(&c->comm, g_grads, g_tail, s));
        SG_COMM(sg::comm_join(&c->comm, s));
        return train_phase(c, st, real, noise_d, noise_g, B, d_grads, g_grads, metrics, 4, s);
    }
    if (phase == 0) {
        if (!whole_step_graph_enabled()) {  // SIGGAN_STEP_GRAPH=0: one graph per phase (A/B comparison)
            SG_TRY(train_phase(c, st, real, noise_d, noise_g, B, d_grads, g_grads, metrics, 1, s));
            SG_TRY(train_phase(c, st, real, noise_d, noise_g, B, d_grads, g_grads, metrics, 22, s));
            {
                StepPackReuse reuse(c);
                SG_TRY(train_phase(c, st, real, noise_d, noise_g, B, d_grads, g_grads, metrics, 3, s));
            }
            return train_phase(c, st, real, noise_d, noise_g, B, d_grads, g_grads, metrics, 4, s);
        }
        return train_phase(c, st, real, noise_d, noise_g, B, d_grads, g_grads, metrics, kWholeStep, s);
    }
    return train_phase(c, st, real, noise_d, noise_g, B, d_grads, g_grads, metrics, phase, s);
}

}  // extern "C"
