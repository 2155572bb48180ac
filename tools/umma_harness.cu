// umma_harness.cu — standalone GPU check of the tcgen05 kernels against straightforward CPU loops.
// Build: make -C signature-gan_b200/csrc harness     Run (GPU box): build/umma_harness [perf]
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../signature-gan_b200/csrc/sg_conv_umma.cuh"
#include "../signature-gan_b200/csrc/sg_kernels.cuh"

namespace sg {
template <typename T>
void final_conv_tanh_stencil(const T* in, const float* scale, const float* shift, const float* w, const float* bias,
                             float* out, uint8_t* out_u8, int B, int S, int C, float act_slope, cudaStream_t s);
template <typename T>
int final_conv_bwd_stencil(const float* dout, const float* out, const T* y, const float* scale, const float* shift,
                           const float* w, T* dbn, float* dW, float* dbias, float* part_w, float* part_bn, int B, int S,
                           int C, float act_slope, cudaStream_t s);
void gfinal_fwd_mma(const bf16* in, const float* scale, const float* shift, const float* w, const float* bias, float* out,
                    uint8_t* out_u8, int B, int S, cudaStream_t s, int C = 32);
int gfinal_bwd_mma(const float* dout, const float* out, const bf16* y, const float* scale, const float* shift,
                   const float* w, bf16* dbn, float* part_w, float* part_bn, int B, int S, int mode, const float* mean,
                   const float* rstd, const float* k1, const float* k2, const float* k3, cudaStream_t s, int ld = 32);
}  // namespace sg

using bf16 = __nv_bfloat16;

#define CK(x)                                                                             \
    do {                                                                                  \
        cudaError_t e_ = (x);                                                             \
        if (e_ != cudaSuccess) {                                                          \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
            exit(2);                                                                      \
        }                                                                                 \
    } while (0)

static uint32_t rng_state = 12345u;
static float frand() {
    rng_state = rng_state * 1664525u + 1013904223u;
    return ((rng_state >> 8) & 0xFFFF) / 32768.0f - 1.0f;
}
static float rbf(float x) { return __bfloat162float(__float2bfloat16(x)); }

struct Dev {
    std::vector<float> h;  // bf16-rounded values as float
    bf16* d = nullptr;
    void init(size_t n, float scale) {
        h.resize(n);
        std::vector<bf16> t(n);
        for (size_t i = 0; i < n; ++i) {
            t[i] = __float2bfloat16(frand() * scale);
            h[i] = __bfloat162float(t[i]);
        }
        CK(cudaMalloc(&d, n * 2));
        CK(cudaMemcpy(d, t.data(), n * 2, cudaMemcpyHostToDevice));
    }
    ~Dev() {
        if (d) cudaFree(d);
    }
};

static int report(const char* name, const std::vector<float>& got, const std::vector<float>& ref, float tol) {
    double max_err = 0, max_ref = 0;
    size_t bad = 0, worst = 0;
    for (size_t i = 0; i < ref.size(); ++i) {
        double e = fabs((double)got[i] - ref[i]);
        if (!(e <= max_err)) {
            if (e > max_err || std::isnan(e)) {
                max_err = e;
                worst = i;
            }
        }
        if (fabs(ref[i]) > max_ref) max_ref = fabs(ref[i]);
        if (!(e <= tol * (1.0 + fabs(ref[i])))) ++bad;
    }
    printf("%-34s max_err=%.4g max_ref=%.4g bad=%zu/%zu (worst idx %zu got %.5g ref %.5g) %s\n", name, max_err, max_ref,
           bad, ref.size(), worst, got[worst], ref[worst], bad == 0 ? "PASS" : "FAIL");
    return bad == 0 ? 0 : 1;
}

static std::vector<float> fetch_bf16(const bf16* d, size_t n) {
    std::vector<bf16> t(n);
    CK(cudaMemcpy(t.data(), d, n * 2, cudaMemcpyDeviceToHost));
    std::vector<float> f(n);
    for (size_t i = 0; i < n; ++i) f[i] = __bfloat162float(t[i]);
    return f;
}

// ---- CPU references (NHWC, weights packed [Cout][16][Cin]) ----
static void cpu_conv_s2(const std::vector<float>& x, const std::vector<float>& w, int N, int H, int W, int Cin,
                        int Cout, std::vector<float>& y) {
    const int OH = H / 2, OW = W / 2;
    y.assign((size_t)N * OH * OW * Cout, 0.f);
    for (int n = 0; n < N; ++n)
        for (int oy = 0; oy < OH; ++oy)
            for (int ox = 0; ox < OW; ++ox)
                for (int co = 0; co < Cout; ++co) {
                    float acc = 0;
                    for (int ky = 0; ky < 4; ++ky) {
                        const int iy = 2 * oy - 1 + ky;
                        if (iy < 0 || iy >= H) continue;
                        for (int kx = 0; kx < 4; ++kx) {
                            const int ix = 2 * ox - 1 + kx;
                            if (ix < 0 || ix >= W) continue;
                            const float* xp = &x[(((size_t)n * H + iy) * W + ix) * Cin];
                            const float* wp = &w[((size_t)co * 16 + ky * 4 + kx) * Cin];
                            for (int ci = 0; ci < Cin; ++ci) acc += xp[ci] * wp[ci];
                        }
                    }
                    y[(((size_t)n * OH + oy) * OW + ox) * Cout + co] = acc;
                }
}
// out[n, oy, ox, co] = sum_{iy,ky: 2iy-1+ky = oy} x[n,iy,ix,ci] * w[co][ky*4+kx][ci]
static void cpu_convT(const std::vector<float>& x, const std::vector<float>& w, int N, int H, int W, int Cin, int Cout,
                      std::vector<float>& y) {
    const int OH = 2 * H, OW = 2 * W;
    y.assign((size_t)N * OH * OW * Cout, 0.f);
    for (int n = 0; n < N; ++n)
        for (int iy = 0; iy < H; ++iy)
            for (int ix = 0; ix < W; ++ix) {
                const float* xp = &x[(((size_t)n * H + iy) * W + ix) * Cin];
                for (int ky = 0; ky < 4; ++ky) {
                    const int oy = 2 * iy - 1 + ky;
                    if (oy < 0 || oy >= OH) continue;
                    for (int kx = 0; kx < 4; ++kx) {
                        const int ox = 2 * ix - 1 + kx;
                        if (ox < 0 || ox >= OW) continue;
                        float* yp = &y[(((size_t)n * OH + oy) * OW + ox) * Cout];
                        for (int co = 0; co < Cout; ++co) {
                            const float* wp = &w[((size_t)co * 16 + ky * 4 + kx) * Cin];
                            float acc = 0;
                            for (int ci = 0; ci < Cin; ++ci) acc += xp[ci] * wp[ci];
                            yp[co] += acc;
                        }
                    }
                }
            }
}
// dW[m][n][tap] = sum coarse[pix][m] * fine[2*pix-1+tap][n]
static void cpu_wgrad(const std::vector<float>& c, const std::vector<float>& f, int N, int cH, int cW, int Mc, int Nf,
                      std::vector<float>& dW) {
    dW.assign((size_t)Mc * Nf * 16, 0.f);
    const int fH = 2 * cH, fW = 2 * cW;
    for (int n = 0; n < N; ++n)
        for (int y = 0; y < cH; ++y)
            for (int x = 0; x < cW; ++x) {
                const float* cp = &c[(((size_t)n * cH + y) * cW + x) * Mc];
                for (int ky = 0; ky < 4; ++ky) {
                    const int fy = 2 * y - 1 + ky;
                    if (fy < 0 || fy >= fH) continue;
                    for (int kx = 0; kx < 4; ++kx) {
                        const int fx = 2 * x - 1 + kx;
                        if (fx < 0 || fx >= fW) continue;
                        const float* fp = &f[(((size_t)n * fH + fy) * fW + fx) * Nf];
                        for (int m = 0; m < Mc; ++m)
                            for (int j = 0; j < Nf; ++j) dW[((size_t)m * Nf + j) * 16 + ky * 4 + kx] += cp[m] * fp[j];
                    }
                }
            }
}

// epi: 0 none, 1 bias + LeakyReLU + dropout mask + gate, 2 mask + gate only (the data-gradient epilogue)
static int test_conv(const char* name, sg::ConvMode mode, int N, int H, int W, int Cin, int Cout, int epi) {
    Dev x, w;
    const int taps = mode == sg::kPlain ? 1 : 16;
    x.init(mode == sg::kPlain ? (size_t)N * Cin : (size_t)N * H * W * Cin, 1.0f);
    w.init((size_t)Cout * taps * Cin, 0.1f);
    std::vector<float> ref;
    size_t out_pix;
    if (mode == sg::kConvS2) {
        cpu_conv_s2(x.h, w.h, N, H, W, Cin, Cout, ref);
        out_pix = (size_t)N * (H / 2) * (W / 2);
    } else if (mode == sg::kConvT) {
        cpu_convT(x.h, w.h, N, H, W, Cin, Cout, ref);
        out_pix = (size_t)N * 4 * H * W;
    } else {
        ref.assign((size_t)N * Cout, 0.f);
        for (int m = 0; m < N; ++m)
            for (int co = 0; co < Cout; ++co) {
                float acc = 0;
                for (int k = 0; k < Cin; ++k) acc += x.h[(size_t)m * Cin + k] * w.h[(size_t)co * Cin + k];
                ref[(size_t)m * Cout + co] = acc;
            }
        out_pix = N;
    }
    const size_t out_n = out_pix * Cout;
    const int opi = (int)(out_pix / N);  // output pixels per image
    sg::ConvGemmArgs a;
    memset(&a, 0, sizeof(a));
    bf16* out;
    CK(cudaMalloc(&out, out_n * 2));
    CK(cudaMemset(out, 0xFF, out_n * 2));
    a.out = out;
    a.ldo = Cout;
    float *bias = nullptr, *mask = nullptr;
    Dev gate;
    if (epi == 3) {  // ReLU gate only (Generator data gradients)
        gate.init(out_n, 1.0f);
        a.gate = gate.d;
        a.slope = 0.f;
        for (size_t i = 0; i < out_n; ++i) ref[i] *= gate.h[i] > 0 ? 1.f : 0.f;
    } else if (epi) {
        std::vector<float> hb(Cout), hm((size_t)N * Cout);
        for (auto& v : hb) v = frand() * 0.5f;
        for (auto& v : hm) v = frand() > -0.5f ? 4.f / 3.f : 0.f;
        if (epi == 2) {
            for (auto& v : hb) v = 0.f;
            a.slope = 0.2f;
        } else {
            CK(cudaMalloc(&bias, Cout * 4));
            CK(cudaMemcpy(bias, hb.data(), Cout * 4, cudaMemcpyHostToDevice));
            a.bias = bias;
            a.act = sg::kActLeaky;
            a.slope = 0.2f;
        }
        if (mode != sg::kPlain) {
            CK(cudaMalloc(&mask, hm.size() * 4));
            CK(cudaMemcpy(mask, hm.data(), hm.size() * 4, cudaMemcpyHostToDevice));
            a.mask = mask;
            a.ldmask = Cout;
            gate.init(out_n, 1.0f);
            a.gate = gate.d;
        }
        for (size_t i = 0; i < out_n; ++i) {
            const int co = (int)(i % Cout);
            const size_t pix = i / Cout;
            float v = ref[i] + hb[co];
            if (epi != 2) v = v > 0 ? v : 0.2f * v;
            if (mode != sg::kPlain) {
                v *= hm[(pix / opi) * Cout + co];
                v *= gate.h[i] > 0 ? 1.f : 0.2f;
            }
            ref[i] = v;
        }
    }
    int rc = sg::launch_conv_gemm(mode, x.d, w.d, N, H, W, Cin, Cout, a, 0);
    if (rc) {
        printf("%-34s LAUNCH ERROR: %s\n", name, sg::umma_last_error());
        return 1;
    }
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        printf("%-34s KERNEL ERROR: %s\n", name, cudaGetErrorString(e));
        exit(3);
    }
    std::vector<float> got = fetch_bf16(out, out_n);
    for (auto& v : ref) v = rbf(v);
    int bad = report(name, got, ref, 0.02f);
    cudaFree(out);
    if (bias) cudaFree(bias);
    if (mask) cudaFree(mask);
    return bad;
}

static int test_wgrad(const char* name, int N, int cH, int cW, int Mc, int Nf) {
    Dev c, f;
    c.init((size_t)N * cH * cW * Mc, 1.0f);
    f.init((size_t)N * 4 * cH * cW * Nf, 1.0f);
    std::vector<float> ref;
    cpu_wgrad(c.h, f.h, N, cH, cW, Mc, Nf, ref);
    const size_t pf = sg::wgrad_partial_floats(N, cH, cW, Mc, Nf);
    float *partial, *dW;
    CK(cudaMalloc(&partial, pf * 4));
    CK(cudaMalloc(&dW, ref.size() * 4));
    CK(cudaMemset(dW, 0xFF, ref.size() * 4));
    int rc = sg::launch_wgrad(c.d, f.d, N, cH, cW, Mc, Nf, partial, pf, dW, 0, 0);
    if (rc) {
        printf("%-34s LAUNCH ERROR: %s\n", name, sg::umma_last_error());
        return 1;
    }
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        printf("%-34s KERNEL ERROR: %s\n", name, cudaGetErrorString(e));
        exit(3);
    }
    std::vector<float> got(ref.size());
    CK(cudaMemcpy(got.data(), dW, ref.size() * 4, cudaMemcpyDeviceToHost));
    double scale = sqrt((double)N * cH * cW);
    int bad = report(name, got, ref, 2e-3f * (float)scale / 10.f + 1e-3f);
    cudaFree(partial);
    cudaFree(dW);
    return bad;
}

// ---- Generator tail (Conv3x3 32->1 + tanh) forward / backward: mma.sync kernels against the streaming stencils
static float* dev_f32(const std::vector<float>& h) {
    float* d;
    CK(cudaMalloc(&d, h.size() * 4));
    CK(cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
    return d;
}
static std::vector<float> fetch_f32(const float* d, size_t n) {
    std::vector<float> h(n);
    CK(cudaMemcpy(h.data(), d, n * 4, cudaMemcpyDeviceToHost));
    return h;
}
static int test_gfinal(const char* name, int B, int S, bool affine, bool perf) {
    const size_t px = (size_t)B * S * S;
    Dev y;
    y.init(px * 32, 2.0f);
    std::vector<float> sc(32), sh(32), w(288), bias(1, 0.05f), dout(px);
    for (int i = 0; i < 32; ++i) {
        sc[i] = 0.6f + 0.4f * frand();
        sh[i] = 0.3f * frand();
    }
    for (auto& v : w) v = 0.08f * frand();
    for (auto& v : dout) v = frand();
    float *d_sc = dev_f32(sc), *d_sh = dev_f32(sh), *d_w = dev_f32(w), *d_b = dev_f32(bias), *d_dout = dev_f32(dout);
    float *out_a, *out_b, *pw, *pbn_a, *pbn_b, *dW_a, *dW_b;
    uint8_t *u8_a, *u8_b;
    bf16 *dbn_a, *dbn_b;
    CK(cudaMalloc(&out_a, px * 4));
    CK(cudaMalloc(&out_b, px * 4));
    CK(cudaMalloc(&u8_a, px));
    CK(cudaMalloc(&u8_b, px));
    CK(cudaMalloc(&dbn_a, px * 64));
    CK(cudaMalloc(&dbn_b, px * 64));
    CK(cudaMalloc(&pw, (size_t)sg::kMaxChunks * 289 * 4));
    CK(cudaMalloc(&pbn_a, (size_t)sg::kMaxChunks * 64 * 4));
    CK(cudaMalloc(&pbn_b, (size_t)sg::kMaxChunks * 64 * 4));
    CK(cudaMalloc(&dW_a, 289 * 4));
    CK(cudaMalloc(&dW_b, 289 * 4));
    const float* scp = affine ? d_sc : nullptr;
    const float* shp = affine ? d_sh : nullptr;
    sg::final_conv_tanh_stencil<bf16>(y.d, scp, shp, d_w, d_b, out_a, u8_a, B, S, 32, 0.f, 0);
    sg::gfinal_fwd_mma(y.d, scp, shp, d_w, d_b, out_b, u8_b, B, S, 0);
    CK(cudaDeviceSynchronize());
    int bad = 0;
    char nm[128];
    snprintf(nm, sizeof(nm), "%s fwd", name);
    bad += report(nm, fetch_f32(out_b, px), fetch_f32(out_a, px), 8e-3f);
    {
        std::vector<uint8_t> a(px), b(px);
        CK(cudaMemcpy(a.data(), u8_a, px, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(b.data(), u8_b, px, cudaMemcpyDeviceToHost));
        size_t off = 0;
        for (size_t i = 0; i < px; ++i) off += abs((int)a[i] - (int)b[i]) > 2;
        printf("%-34s u8 mismatches(>2)=%zu %s\n", nm, off, off ? "FAIL" : "PASS");
        bad += off != 0;
    }
    if (affine) {
        const int ca = sg::final_conv_bwd_stencil<bf16>(d_dout, out_a, y.d, d_sc, d_sh, d_w, dbn_a, dW_a, dW_a + 288, pw,
                                                        pbn_a, B, S, 32, 0.f, 0);
        const int cb = sg::gfinal_bwd_mma(d_dout, out_a, y.d, d_sc, d_sh, d_w, dbn_b, pw, pbn_b, B, S, 0, nullptr, nullptr,
                                          nullptr, nullptr, nullptr, 0);
        sg::vec_finalize(pw, cb, 289, dW_b, 288, dW_b + 288, 0);
        CK(cudaDeviceSynchronize());
        snprintf(nm, sizeof(nm), "%s bwd dbn", name);
        bad += report(nm, fetch_bf16(dbn_b, px * 32), fetch_bf16(dbn_a, px * 32), 1.5e-2f);
        const double nscale = sqrt((double)px);
        snprintf(nm, sizeof(nm), "%s bwd dW,dbias", name);
        bad += report(nm, fetch_f32(dW_b, 289), fetch_f32(dW_a, 289), (float)(2e-3 * nscale + 1e-3));
        auto fold = [&](const float* p, int chunks) {
            std::vector<float> h = fetch_f32(p, (size_t)chunks * 64), r(64, 0.f);
            for (int c = 0; c < chunks; ++c)
                for (int i = 0; i < 64; ++i) r[i] += h[(size_t)c * 64 + i];
            return r;
        };
        snprintf(nm, sizeof(nm), "%s bwd bn sums", name);
        bad += report(nm, fold(pbn_b, cb), fold(pbn_a, ca), (float)(4e-3 * nscale + 1e-3));
        // two-pass form: mode 1 (reductions only) must reproduce mode 0's sums; mode 2 = mode 0 + bn_bwd_apply
        std::vector<float> r0 = fold(pbn_b, cb), w0 = fetch_f32(dW_b, 289);
        CK(cudaMemset(pbn_b, 0, (size_t)sg::kMaxChunks * 64 * 4));
        const int c1 = sg::gfinal_bwd_mma(d_dout, out_a, y.d, d_sc, d_sh, d_w, nullptr, pw, pbn_b, B, S, 1, nullptr, nullptr,
                                          nullptr, nullptr, nullptr, 0);
        sg::vec_finalize(pw, c1, 289, dW_b, 288, dW_b + 288, 0);
        CK(cudaDeviceSynchronize());
        snprintf(nm, sizeof(nm), "%s bwd pass1 sums", name);
        bad += report(nm, fold(pbn_b, c1), r0, 1e-6f);
        snprintf(nm, sizeof(nm), "%s bwd pass1 dW", name);
        bad += report(nm, fetch_f32(dW_b, 289), w0, 1e-6f);
        std::vector<float> mean(32), rstd(32), k1(32), k2(32), k3(32);
        for (int i = 0; i < 32; ++i) {
            mean[i] = 0.2f * frand();
            rstd[i] = 0.9f + 0.3f * frand();
            k1[i] = 1.0f + 0.2f * frand();
            k2[i] = 0.01f * frand();
            k3[i] = 0.02f * frand();
        }
        float *d_mean = dev_f32(mean), *d_rstd = dev_f32(rstd), *d_k1 = dev_f32(k1), *d_k2 = dev_f32(k2), *d_k3 = dev_f32(k3);
        sg::bn_bwd_apply<bf16>(dbn_a, y.d, d_mean, d_rstd, d_k1, d_k2, d_k3, dbn_a, (long)px, 32, 0);
        sg::gfinal_bwd_mma(d_dout, out_a, y.d, d_sc, d_sh, d_w, dbn_b, nullptr, nullptr, B, S, 2, d_mean, d_rstd, d_k1, d_k2,
                           d_k3, 0);
        CK(cudaDeviceSynchronize());
        snprintf(nm, sizeof(nm), "%s bwd pass2 dy", name);
        bad += report(nm, fetch_bf16(dbn_b, px * 32), fetch_bf16(dbn_a, px * 32), 1.5e-2f);
        if (perf) {
            cudaEvent_t e0, e1;
            cudaEventCreate(&e0);
            cudaEventCreate(&e1);
            float m1, m2;
            for (int k = 1; k <= 2; ++k) {
                for (int i = 0; i < 13; ++i) {
                    if (i == 3) cudaEventRecord(e0);
                    sg::gfinal_bwd_mma(d_dout, out_a, y.d, d_sc, d_sh, d_w, k == 1 ? nullptr : dbn_b, pw, pbn_b, B, S, k, d_mean,
                                       d_rstd, d_k1, d_k2, d_k3, 0);
                }
                cudaEventRecord(e1);
                CK(cudaDeviceSynchronize());
                cudaEventElapsedTime(k == 1 ? &m1 : &m2, e0, e1);
            }
            printf("PERF %-24s bwd pass1 (sums) %.3f ms, pass2 (apply) %.3f ms\n", name, m1 / 10, m2 / 10);
        }
        for (float* p : {d_mean, d_rstd, d_k1, d_k2, d_k3}) cudaFree(p);
    }
    if (perf) {
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        float ms[4];
        for (int k = 0; k < 4; ++k) {
            for (int i = 0; i < 13; ++i) {
                if (i == 3) cudaEventRecord(e0);
                if (k == 0) sg::final_conv_tanh_stencil<bf16>(y.d, scp, shp, d_w, d_b, out_a, nullptr, B, S, 32, 0.f, 0);
                if (k == 1) sg::gfinal_fwd_mma(y.d, scp, shp, d_w, d_b, out_b, nullptr, B, S, 0);
                if (k == 2 && affine)
                    sg::final_conv_bwd_stencil<bf16>(d_dout, out_a, y.d, d_sc, d_sh, d_w, dbn_a, dW_a, dW_a + 288, pw, pbn_a,
                                                     B, S, 32, 0.f, 0);
                if (k == 3 && affine)
                    sg::gfinal_bwd_mma(d_dout, out_a, y.d, d_sc, d_sh, d_w, dbn_b, pw, pbn_b, B, S, 0, nullptr, nullptr, nullptr,
                                       nullptr, nullptr, 0);
            }
            cudaEventRecord(e1);
            CK(cudaDeviceSynchronize());
            cudaEventElapsedTime(&ms[k], e0, e1);
            ms[k] /= 10;
        }
        const double gb_f = px * (64.0 + 4.0) * 1e-9, gb_b = px * (128.0 + 8.0) * 1e-9;
        printf("PERF %-24s fwd stencil %.3f ms (%.0f GB/s) mma %.3f ms (%.0f GB/s) | bwd stencil %.3f ms (%.0f GB/s) mma %.3f ms (%.0f GB/s)\n",
               name, ms[0], gb_f / ms[0] * 1e3, ms[1], gb_f / ms[1] * 1e3, ms[2], gb_b / ms[2] * 1e3, ms[3],
               gb_b / ms[3] * 1e3);
    }
    for (void* p : {(void*)d_sc, (void*)d_sh, (void*)d_w, (void*)d_b, (void*)d_dout, (void*)out_a, (void*)out_b, (void*)pw,
                    (void*)pbn_a, (void*)pbn_b, (void*)dW_a, (void*)dW_b, (void*)u8_a, (void*)u8_b, (void*)dbn_a,
                    (void*)dbn_b})
        cudaFree(p);
    return bad;
}

// Operands of the timing runs: random bf16 (a 4 MB random block replicated) — all-zero operands draw less power,
// run at higher clocks and made the MMA-bound kernels look 15-25 % faster than inside the training step.
static void fill_random_bf16(bf16* d, size_t n, float scale) {
    const size_t blk = 2u << 20;
    std::vector<bf16> h(blk);
    for (auto& v : h) v = __float2bfloat16(frand() * scale);
    for (size_t off = 0; off < n; off += blk)
        CK(cudaMemcpy(d + off, h.data(), std::min(blk, n - off) * 2, cudaMemcpyHostToDevice));
}

static bool want(const char* name);
static void perf_conv(const char* name, sg::ConvMode mode, int N, int H, int W, int Cin, int Cout, int epi = 0) {
    if (!want(name)) return;
    bf16 *x, *w, *out;
    const size_t xn = (size_t)N * H * W * Cin, wn = (size_t)Cout * 16 * Cin;
    const size_t on = mode == sg::kConvS2 ? (size_t)N * H / 2 * W / 2 * Cout : (size_t)N * H * 2 * W * 2 * Cout;
    CK(cudaMalloc(&x, xn * 2));
    CK(cudaMalloc(&w, wn * 2));
    CK(cudaMalloc(&out, on * 2));
    fill_random_bf16(x, xn, 1.0f);
    fill_random_bf16(w, wn, 0.05f);
    sg::ConvGemmArgs a;
    memset(&a, 0, sizeof(a));
    a.out = out;
    a.ldo = Cout;
    bf16* gate = nullptr;
    float *mask = nullptr, *bias = nullptr;
    if (epi == 1 || epi == 3 || epi == 4) {  // data-gradient epilogue: gate (saved activation) + dropout mask
        a.slope = 0.2f;
        if (epi != 4) {
            CK(cudaMalloc(&gate, on * 2));
            CK(cudaMemset(gate, 0x3f, on * 2));
            a.gate = gate;
        }
        if (epi != 3) {
            CK(cudaMalloc(&mask, (size_t)N * Cout * 4));
            CK(cudaMemset(mask, 0, (size_t)N * Cout * 4));
            a.mask = mask;
            a.ldmask = Cout;
        }
    } else if (epi == 2) {  // forward epilogue: bias + LeakyReLU + dropout mask
        CK(cudaMalloc(&bias, Cout * 4));
        CK(cudaMemset(bias, 0, Cout * 4));
        CK(cudaMalloc(&mask, (size_t)N * Cout * 4));
        CK(cudaMemset(mask, 0, (size_t)N * Cout * 4));
        a.bias = bias;
        a.act = sg::kActLeaky;
        a.slope = 0.2f;
        a.mask = mask;
        a.ldmask = Cout;
    }
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) sg::launch_conv_gemm(mode, x, w, N, H, W, Cin, Cout, a, 0);
    CK(cudaDeviceSynchronize());
    const int iters = 10;
    cudaEventRecord(e0);
    for (int i = 0; i < iters; ++i) sg::launch_conv_gemm(mode, x, w, N, H, W, Cin, Cout, a, 0);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    ms /= iters;
    const double flops = 2.0 * (mode == sg::kConvS2 ? (double)N * H / 2 * W / 2 * 16 : (double)N * H * W * 16) * Cin * Cout;
    printf("PERF %-28s %.3f ms  %.1f TFLOP/s  (in %.0f MB, out %.0f MB)\n", name, ms, flops / ms * 1e-9, xn * 2e-6,
           on * 2e-6);
    cudaFree(x);
    cudaFree(w);
    cudaFree(out);
    if (gate) cudaFree(gate);
    if (mask) cudaFree(mask);
    if (bias) cudaFree(bias);
}

static void perf_wgrad(const char* name, int N, int cH, int cW, int Mc, int Nf) {
    if (!want(name)) return;
    bf16 *c, *f;
    const size_t cn = (size_t)N * cH * cW * Mc, fn = (size_t)N * 4 * cH * cW * Nf;
    CK(cudaMalloc(&c, cn * 2));
    CK(cudaMalloc(&f, fn * 2));
    fill_random_bf16(c, cn, 1.0f);
    fill_random_bf16(f, fn, 1.0f);
    const size_t pf = sg::wgrad_partial_floats(N, cH, cW, Mc, Nf);
    float *partial, *dW;
    CK(cudaMalloc(&partial, pf * 4));
    CK(cudaMalloc(&dW, (size_t)Mc * Nf * 64));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) sg::launch_wgrad(c, f, N, cH, cW, Mc, Nf, partial, pf, dW, 0, 0);
    CK(cudaDeviceSynchronize());
    const int iters = 10;
    cudaEventRecord(e0);
    for (int i = 0; i < iters; ++i) sg::launch_wgrad(c, f, N, cH, cW, Mc, Nf, partial, pf, dW, 0, 0);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    ms /= iters;
    const double flops = 2.0 * (double)N * cH * cW * 16 * Mc * Nf;
    printf("PERF %-28s %.3f ms  %.1f TFLOP/s (partial %.0f MB)\n", name, ms, flops / ms * 1e-9, pf * 4e-6);
    cudaFree(c);
    cudaFree(f);
    cudaFree(partial);
    cudaFree(dW);
}

static const char* g_filter = nullptr;
static bool want(const char* name) { return !g_filter || strstr(name, g_filter); }

int main(int argc, char** argv) {
    const bool perfonly = argc > 1 && !strcmp(argv[1], "perfonly");  // perfonly [name-substring]: skip the checks
    if (perfonly && argc > 2) g_filter = argv[2];
    int only = (!perfonly && argc > 2) ? atoi(argv[2]) : (perfonly ? 1 << 30 : -1);
    int fails = 0, t = 0;
#define RUN(expr)                          \
    do {                                   \
        if (only < 0 || only == t) fails += (expr); \
        ++t;                               \
    } while (0)
    RUN(test_conv("plain 256x128x256", sg::kPlain, 256, 1, 1, 256, 128, false));
    RUN(test_conv("plain 200x4096x128 +epi", sg::kPlain, 200, 1, 1, 128, 4096, true));
    RUN(test_conv("convS2 8x32x32x64->128", sg::kConvS2, 8, 32, 32, 64, 128, false));
    RUN(test_conv("convS2 8x32x32x64->128 +epi", sg::kConvS2, 8, 32, 32, 64, 128, true));
    RUN(test_conv("convS2 16x8x8x256->512", sg::kConvS2, 16, 8, 8, 256, 512, true));
    RUN(test_conv("convS2 5x8x8x128->256 (tail)", sg::kConvS2, 5, 8, 8, 128, 256, true));
    RUN(test_conv("convS2 4x64x64x32->32 (thin)", sg::kConvS2, 4, 64, 64, 32, 32, true));
    RUN(test_conv("convS2 4x32x32x32->64 (thin)", sg::kConvS2, 4, 32, 32, 32, 64, false));
    RUN(test_conv("convT 8x4x4x256->128", sg::kConvT, 8, 4, 4, 256, 128, false));
    RUN(test_conv("convT 8x4x4x256->128 +epi", sg::kConvT, 8, 4, 4, 256, 128, true));
    RUN(test_conv("convT 4x16x16x64->32", sg::kConvT, 4, 16, 16, 64, 32, true));
    RUN(test_conv("convT 2x32x32x32->32 (thin)", sg::kConvT, 2, 32, 32, 32, 32, true));
    RUN(test_conv("convT 3x8x8x512->256", sg::kConvT, 3, 8, 8, 512, 256, false));
    // more tiles than SMs: every persistent CTA walks several tiles and both TMEM accumulators wrap
    RUN(test_conv("convS2 48x64x64x32->32 (384 tiles)", sg::kConvS2, 48, 64, 64, 32, 32, true));
    RUN(test_conv("convT 24x32x32x32->32 (768 tiles)", sg::kConvT, 24, 32, 32, 32, 32, true));
    RUN(test_conv("convS2 1280x8x8x64->128 (160 t)", sg::kConvS2, 1280, 8, 8, 64, 128, true));
    RUN(test_conv("convT 700x4x4x64->64 (352 tiles)", sg::kConvT, 700, 4, 4, 64, 64, true));
    RUN(test_conv("plain 40000x64x256 (626 tiles)", sg::kPlain, 40000, 1, 1, 64, 256, true));
    // phase-fused thin transposed convolutions (sg_convt4.cu): no bias/mask/gate epilogue
    RUN(test_conv("convT4 6x32x32x32->32", sg::kConvT, 6, 32, 32, 32, 32, false));
    RUN(test_conv("convT4 40x32x32x32->32 (320 t)", sg::kConvT, 40, 32, 32, 32, 32, false));
    RUN(test_conv("convT4 3x64x64x32->32 (128x128 tail)", sg::kConvT, 3, 64, 64, 32, 32, false));
    RUN(test_conv("convT4 5x16x16x64->32", sg::kConvT, 5, 16, 16, 64, 32, false));
    RUN(test_conv("convT4 200x16x16x64->32 (400 t)", sg::kConvT, 200, 16, 16, 64, 32, false));
    RUN(test_wgrad("wgrad 8x16x16 128|64", 8, 16, 16, 128, 64));
    RUN(test_wgrad("wgrad 16x4x4 512|256", 16, 4, 4, 512, 256));
    RUN(test_wgrad("wgrad 8x8x8 256|128", 8, 8, 8, 256, 128));
    RUN(test_wgrad("wgrad 2x32x32 32|32 (thin)", 2, 32, 32, 32, 32));
    RUN(test_wgrad("wgrad 3x16x16 64|32 (thin)", 3, 16, 16, 64, 32));
    RUN(test_wgrad("wgrad 1x32x32 32|32 (pair)", 1, 32, 32, 32, 32));
    RUN(test_wgrad("wgrad 70x32x32 32|32 (pair, 560 t)", 70, 32, 32, 32, 32));
    RUN(test_wgrad("wgrad 5x4x4 256|128 (tail)", 5, 4, 4, 256, 128));
    // CTA-pair kernels (sg_conv2_umma.cu): D conv1 forward and the D conv1 / conv2 data-gradient shapes, with tails
    RUN(test_conv("conv2 S2 3x32x32x64->128 +epi", sg::kConvS2, 3, 32, 32, 64, 128, true));
    RUN(test_conv("conv2 S2 161x32x32x64->128", sg::kConvS2, 161, 32, 32, 64, 128, true));
    RUN(test_conv("conv2 T4 5x16x16x128->64 +epi", sg::kConvT, 5, 16, 16, 128, 64, true));
    RUN(test_conv("conv2 T4 171x16x16x128->64", sg::kConvT, 171, 16, 16, 128, 64, false));
    RUN(test_conv("conv2 T4 171x16x16x128->64 +gate,mask", sg::kConvT, 171, 16, 16, 128, 64, 2));
    RUN(test_conv("conv2 T4 3x32x32x128->64 +gate,mask", sg::kConvT, 3, 32, 32, 128, 64, 2));
    RUN(test_conv("conv2 S2 3x64x64x64->128 +epi", sg::kConvS2, 3, 64, 64, 64, 128, true));
    RUN(test_conv("conv2 T4 3x32x32x128->64 +epi", sg::kConvT, 3, 32, 32, 128, 64, true));
    RUN(test_conv("conv2 T2 7x8x8x256->128 +epi", sg::kConvT, 7, 8, 8, 256, 128, true));
    RUN(test_conv("conv2 T2 341x8x8x256->128", sg::kConvT, 341, 8, 8, 256, 128, true));
    // pixel-pair stride-2 kernel with col2im epilogue (sg_convs2_thin.cu): generator's last block, data gradient
    RUN(test_conv("convS2 thin 1x64x64x32->32", sg::kConvS2, 1, 64, 64, 32, 32, false));
    RUN(test_conv("convS2 thin 5x64x64x32->32 +gate", sg::kConvS2, 5, 64, 64, 32, 32, 3));
    RUN(test_conv("convS2 thin 100x64x64x32->32 +gate (800 t)", sg::kConvS2, 100, 64, 64, 32, 32, 3));
    RUN(test_gfinal("gfinal 5x64 train", 5, 64, true, false));
    RUN(test_gfinal("gfinal 3x64 eval", 3, 64, false, false));
    RUN(test_gfinal("gfinal 3x128 train", 3, 128, true, false));
    RUN(test_gfinal("gfinal 700x64 train (multi-image CTAs)", 700, 64, true, false));
    printf("harness: %d failing tests\n", fails);
    if (argc > 1 && (!strcmp(argv[1], "perf") || perfonly) && fails == 0) {
        const int B = 4096;
        perf_conv("D c1 64->128 @32x32", sg::kConvS2, B, 32, 32, 64, 128);
        perf_conv("D c2 128->256 @16x16", sg::kConvS2, B, 16, 16, 128, 256);
        perf_conv("D c3 256->512 @8x8", sg::kConvS2, B, 8, 8, 256, 512);
        perf_conv("G up0 256->128 @4x4", sg::kConvT, B, 4, 4, 256, 128);
        perf_conv("G up1 128->64 @8x8", sg::kConvT, B, 8, 8, 128, 64);
        perf_conv("G up2 64->32 @16x16", sg::kConvT, B, 16, 16, 64, 32);
        perf_conv("G up3 32->32 @32x32", sg::kConvT, B, 32, 32, 32, 32);
        perf_conv("G up3 dgrad 32->32 @64x64 +gate", sg::kConvS2, B, 64, 64, 32, 32, 3);
        perf_conv("G up2 dgrad 32->64 @32x32 +gate", sg::kConvS2, B, 32, 32, 32, 64, 3);
        perf_conv("D dgrad c3 512->256 @4x4", sg::kConvT, B, 4, 4, 512, 256);
        perf_conv("D dgrad c2 256->128 @8x8", sg::kConvT, B, 8, 8, 256, 128);
        perf_conv("D dgrad c1 128->64 @16x16", sg::kConvT, B, 16, 16, 128, 64);
        perf_conv("D c1 +bias,act,mask", sg::kConvS2, B, 32, 32, 64, 128, 2);
        perf_conv("D c2 +bias,act,mask", sg::kConvS2, B, 16, 16, 128, 256, 2);
        perf_conv("D dgrad c3 +gate,mask", sg::kConvT, B, 4, 4, 512, 256, 1);
        perf_conv("D dgrad c2 +gate,mask", sg::kConvT, B, 8, 8, 256, 128, 1);
        perf_conv("D dgrad c1 +gate,mask", sg::kConvT, B, 16, 16, 128, 64, 1);
        perf_conv("D dgrad c1 +gate only", sg::kConvT, B, 16, 16, 128, 64, 3);
        perf_conv("D dgrad c1 +mask only", sg::kConvT, B, 16, 16, 128, 64, 4);
        perf_conv("D dgrad c2 +gate only", sg::kConvT, B, 8, 8, 256, 128, 3);
        perf_conv("D dgrad c2 +mask only", sg::kConvT, B, 8, 8, 256, 128, 4);
        perf_wgrad("wgrad c3 512|256 @4x4", B, 4, 4, 512, 256);
        perf_wgrad("wgrad c2 256|128 @8x8", B, 8, 8, 256, 128);
        perf_wgrad("wgrad c1 128|64 @16x16", B, 16, 16, 128, 64);
        perf_wgrad("wgrad up3 32|32 @32x32", B, 32, 32, 32, 32);
        if (want("gfinal")) {
            test_gfinal("gfinal 4096x64", B, 64, true, true);
            test_gfinal("gfinal 1024x128", 1024, 128, true, true);
        }
    }
    return fails ? 1 : 0;
}
