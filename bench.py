#!/usr/bin/env python
"""bench.py — training images/sec (full D+G step) and sampling images/sec of the B200 signature-GAN path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...     (one rank per GPU, NCCL)

One "step" = one full adversarial training step (D step + G step, vanilla_gan_model.VanillaGAN.train_step) over a
batch of 4096 synthetic 64x64 signatures per GPU (BASELINE.json configs[2]); weak scaling over data-parallel ranks
with an NCCL all-reduce of the flat D and G gradient buckets. Rank 0 prints ONE JSON line.
`--impl reference` times the reference's own CPU implementation of the same step (the unmodified reference modules
from oracle/_ref — see oracle/make_ref.py — torch CPU fp32, all host threads) on a bounded sample (batch 64 per step,
the reference's own default configuration).
"""
import argparse
import json
import math
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "signature-gan_b200"))

FLOP_PER_IMG_TRAIN = {64: 1.9707e9, 128: 9.2199e9}     # algorithmically necessary conv/linear FLOPs
FLOP_PER_IMG_SAMPLE = {64: 87.06e6, 128: 413.73e6}


def flops_per_image(size, width=1):
    """(training step, sampling) conv / linear FLOPs per image (2 x MAC) of the ladders gen…:131-149 / disc…:131-194 with
    every hidden width multiplied by `width`; width 1 reproduces the SURVEY.md §8d figures above."""
    g = [c * width for c in ((256, 128, 64, 32, 32) if size == 64 else (512, 256, 128, 64, 32, 32))]
    d = [1] + [c * width for c in ((64, 128, 256, 512) if size == 64 else (64, 128, 256, 512, 512))]
    fc = 2.0 * 100 * g[0] * 16
    f_g = fc + sum(2.0 * (4 << i) ** 2 * 16 * g[i] * g[i + 1] for i in range(len(g) - 1)) + 2.0 * size * size * 9 * g[-1]
    c0 = 2.0 * (size // 2) ** 2 * 16 * d[1]
    f_d = sum(2.0 * (size >> (i + 1)) ** 2 * 16 * d[i] * d[i + 1] for i in range(len(d) - 1)) + 2.0 * d[-1] * 16
    return 8 * f_d + 4 * f_g - 2 * c0 - fc, f_g
# sampling, algorithmic HBM bytes per image with the eval tail fused: z (400) + every bf16 level up to the input of the
# last block written and read once (8K + 16K + 32K + 64K, x2) + the fp32 image (16K)
SAMPLE_BYTES_PER_IMG = {64: 400 + 2 * (8192 + 16384 + 32768 + 65536) + 16384}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=4096, help="images per GPU per step")
    ap.add_argument("--size", type=int, default=64)
    ap.add_argument("--width", type=int, default=1, choices=[1, 2],
                    help="hidden-width multiplier (2 = the '2x hidden width' variant of BASELINE configs[4]; bf16 only)")
    ap.add_argument("--sampling-batch", type=int, default=16384)
    ap.add_argument("--cpu-batch", type=int, default=64)
    ap.add_argument("--no-extras", action="store_true", help="skip sampling / cpu baseline / per-op profile")
    ap.add_argument("--cpu-child", action="store_true", help=argparse.SUPPRESS)
    ap.add_argument("--torch-allreduce", action="store_true", help="N > 1: all-reduce the buckets through torch.distributed "
                    "instead of the library-owned NCCL communicator (A/B)")
    ap.add_argument("--sync-bn", action="store_true", help="N > 1: global-batch BatchNorm statistics (sg_set_sync_batchnorm)")
    return ap.parse_args()


def peaks():
    p = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            m = json.load(f)
        p.update({k: m[k] for k in ("hbm_gbs", "bf16_tflops", "bf16_tflops_sustained") if k in m})
        p["source"] = "measured"
    except Exception:
        pass
    return p


# ------------------------------------------------------------------------------------------------
# synthetic signatures (white background +1, dark strokes -1), generated with torch on the target device
# ------------------------------------------------------------------------------------------------
def synthetic_signatures(n, size, device, seed=1234):
    import torch
    g = torch.Generator(device=device).manual_seed(seed)
    t = torch.linspace(0, 1, 192, device=device).view(1, 1, -1)
    u = torch.rand(n, 3, 16, device=device, generator=g)

    def U(k, lo, hi):
        return (lo + (hi - lo) * u[:, :, k]).unsqueeze(-1)

    x = U(0, .4, .6) + U(2, .5, .8) * (t - .5)
    y = U(1, .35, .65) + U(3, -.1, .1) * (t - .5)
    for k in range(1, 4):
        x = x + U(3 + k, -.06, .06) / k * torch.sin(2 * math.pi * (k * t + U(9 + k, 0, 1)))
        y = y + U(6 + k, -.18, .18) / k * torch.sin(2 * math.pi * (k * t + U(12 + k, 0, 1)))
    xi = (x.clamp(0, 1) * (size - 1)).round().long().reshape(n, -1)
    yi = (y.clamp(0, 1) * (size - 1)).round().long().reshape(n, -1)
    ink = torch.zeros(n, size * size, device=device)
    ink.scatter_(1, yi * size + xi, 1.0)
    ink = ink.view(n, 1, size, size)
    ink = torch.nn.functional.avg_pool2d(torch.nn.functional.max_pool2d(ink, 3, 1, 1), 3, 1, 1)
    return (1.0 - 2.0 * ink).clamp(-1, 1).contiguous()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            import torch
            uuid = str(torch.cuda.get_device_properties(device_index).uuid)
            ident = uuid if uuid.startswith("GPU-") else "GPU-" + uuid
            self.p = subprocess.Popen(["nvidia-smi", f"--id={ident}", f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "200"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 7:
                continue
            try:
                sm.append(float(c[0])); mx.append(float(c[1])); pw.append(float(c[2]))
            except ValueError:
                continue
            for nm, v in zip(names, c[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if sm:
            sm.sort()
            out.update({"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "power_w_max": max(pw),
                        "reasons": sorted(reasons), "samples": len(sm)})
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        return out


# ------------------------------------------------------------------------------------------------
# CPU arm: the UNMODIFIED reference (oracle/_ref, built by oracle/make_ref.py) on the host cores — its own
# VanillaGAN(device='cpu').train_step / .generate (vanilla…:308-371), torch CPU fp32, all host threads. Falls back to
# the oracle port of the same algorithm (kind "port") only when oracle/_ref is absent.
# ------------------------------------------------------------------------------------------------
def host_threads():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def _import_reference():
    """The reference's vanilla_gan_model from oracle/_ref/src (first on sys.path so that its own generator_ /
    discriminator_ modules resolve to the reference's, not to this repo's drop-ins). None when the copy is absent."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    try:
        import make_ref
        ref_src = make_ref.src_dir()
    except Exception:
        return None
    if "vanilla_gan_model" in sys.modules:      # this process already holds the drop-in modules: cannot mix
        return None
    sys.path.insert(0, ref_src)
    import vanilla_gan_model as R
    assert os.path.realpath(R.__file__).startswith(os.path.realpath(ref_src)), R.__file__
    return R


def cpu_reference_numbers(size, batch, steps, warmup, sampling=True):
    """Runs in a process of its own (see cpu_baseline_subprocess): times the reference on the host CPU."""
    import torch
    threads = host_threads()
    torch.set_num_threads(threads)        # torchrun exports OMP_NUM_THREADS=1: set the count explicitly
    R = _import_reference()
    out = {"cores": threads, "torch_threads": torch.get_num_threads()}
    real = synthetic_signatures(4 * batch, size, "cpu", seed=1234).view(4, batch, 1, size, size)
    if R is not None:
        torch.manual_seed(0)
        gan = R.VanillaGAN(latent_dim=100, image_size=size, device="cpu")
        for i in range(warmup):
            gan.train_step(real[i % 4])
        t0 = time.perf_counter()
        for i in range(steps):
            m = gan.train_step(real[i % 4])
        dt = time.perf_counter() - t0
        out.update({"kind": "reference", "train_images_per_s": batch * steps / dt, "s_per_step": dt / steps,
                    "last": {k: float(v) for k, v in m.items()},
                    "sample": f"{steps} VanillaGAN(device='cpu').train_step calls of the unmodified reference "
                              f"(oracle/_ref), batch {batch} ({size}x{size}), {warmup} warm-up, torch CPU fp32, "
                              f"{threads} threads"})
        if sampling:
            gan.generate(64)
            t0 = time.perf_counter()
            n_s, reps = 1024, 5
            for _ in range(reps):
                gan.generate(n_s)
            dt = time.perf_counter() - t0
            out["sampling"] = {"value": n_s * reps / dt, "unit": "images/s", "cores": threads, "kind": "reference",
                               "sample": f"{reps} x VanillaGAN(device='cpu').generate({n_s}) of the unmodified "
                                         f"reference, {size}x{size}, torch CPU fp32"}
        return out
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import siggan_oracle as O
    g_sd, d_sd = O.make_state_dicts(size, 100, seed=0)
    g_opt = O.AdamState(g_sd, O.trainable_names(g_sd))
    d_opt = O.AdamState(d_sd, O.trainable_names(d_sd))
    t0 = None
    for i in range(warmup + steps):
        if i == warmup:
            t0 = time.perf_counter()
        masks_r, masks_f = O.make_dropout_masks(batch, size, 2 * i), O.make_dropout_masks(batch, size, 2 * i + 1)
        O.d_step(g_sd, d_sd, d_opt, real[i % 4], torch.randn(batch, 100), size, masks_r, masks_f)
        O.g_step(g_sd, d_sd, g_opt, torch.randn(batch, 100), size)
    dt = time.perf_counter() - t0
    out.update({"kind": "port", "train_images_per_s": batch * steps / dt, "s_per_step": dt / steps,
                "sample": f"{steps} D+G steps of the oracle port (oracle/_ref absent), batch {batch} ({size}x{size}), "
                          f"torch CPU fp32, {threads} threads"})
    return out


def cpu_baseline_subprocess(size, batch, steps, warmup, sampling=True):
    """The GPU arm's process has this repo's drop-in modules loaded under the reference's module names, so the CPU
    baseline runs in a child process (`bench.py --cpu-child`), which prints one JSON object."""
    cmd = [sys.executable, os.path.abspath(__file__), "--cpu-child", "--size", str(size), "--cpu-batch", str(batch),
           "--steps", str(steps), "--warmup", str(warmup)] + ([] if sampling else ["--no-extras"])
    env = {k: v for k, v in os.environ.items() if k not in ("OMP_NUM_THREADS", "MKL_NUM_THREADS")}
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=env, timeout=900)
    if r.returncode != 0:
        raise RuntimeError("cpu baseline child failed: " + r.stderr.decode()[-400:])
    return json.loads(r.stdout.decode().strip().splitlines()[-1])


def run_reference(args, rank):
    if rank != 0:
        return
    r = cpu_reference_numbers(args.size, args.cpu_batch, args.steps, args.warmup, sampling=False)
    ips, spp = r["train_images_per_s"], r["s_per_step"]
    line = {
        "impl": "reference", "metric": "training images/sec (G+D step)", "value": ips, "unit": "images/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": spp * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"train_step_{args.size}x{args.size}_b{args.batch}_per_gpu",
                   "global_batch": args.gpus * args.batch, "image_size": args.size, "latent_dim": 100,
                   "parallelism": f"dp{args.gpus}", "n_critic": 1, "cpu_sample_batch": args.cpu_batch,
                   "note": "the reference's own CPU implementation of the step (unmodified modules from oracle/_ref, "
                           "torch CPU fp32, all host threads) timed on a bounded sample of the workload: batch "
                           f"{args.cpu_batch} per step, the reference's default configuration"},
        "cpu_baseline": {"value": ips, "unit": "images/s", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]},
        "e2e": {"value": ips, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), file=RESULT_OUT, flush=True)


# ------------------------------------------------------------------------------------------------
# the unchanged reference trainer's loop over the drop-in modules (second figure)
# ------------------------------------------------------------------------------------------------
def trainer_style_numbers(gan, pool, B, dev, steps=10, warmup=3):
    """GANTrainer._train_discriminator / _train_generator (train…:281-376) statement for statement — module calls,
    criterion, loss.backward(), optimizer.step() and the 7 `.item()` reads per step — which is what
    train_vanilla_gan_signatures.py executes when it imports this repo's modules (gradient clipping off, its default)."""
    import torch
    D, G, crit = gan.discriminator, gan.generator, gan.criterion
    n_pool = pool.shape[0]

    def one(real):
        D.train()
        G.eval()
        gan.d_optimizer.zero_grad()
        real_labels = torch.full((B, 1), gan.label_smoothing, device=dev)
        real_preds = D(real)
        d_loss_real = crit(real_preds, real_labels)
        noise = torch.randn(B, gan.latent_dim, device=dev)
        with torch.no_grad():
            fake = G(noise)
        fake_preds = D(fake)
        d_loss_fake = crit(fake_preds, torch.zeros(B, 1, device=dev))
        d_loss = d_loss_real + d_loss_fake
        d_loss.backward()
        gan.d_optimizer.step()
        vals = [d_loss.item(), d_loss_real.item(), d_loss_fake.item(), real_preds.mean().item(), fake_preds.mean().item()]
        G.train()
        D.eval()
        gan.g_optimizer.zero_grad()
        noise = torch.randn(B, gan.latent_dim, device=dev)
        fake_preds = D(G(noise))
        g_loss = crit(fake_preds, torch.ones(B, 1, device=dev))
        g_loss.backward()
        gan.g_optimizer.step()
        return vals + [g_loss.item(), fake_preds.mean().item()]

    for i in range(warmup):
        one(pool[i % n_pool])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(steps):
        vals = one(pool[i % n_pool])
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return {"images_per_s": B * steps / dt, "ms_per_step": dt / steps * 1e3, "steps": steps, "item_reads_per_step": 7,
            "api": "module calls + criterion + backward + optimizer.step, as GANTrainer (train…:281-376) drives them",
            "finite": all(v == v for v in vals)}


# ------------------------------------------------------------------------------------------------
# input pipeline: DeviceSignatureLoader (sg_augment_batch) next to the reference's PIL / torchvision transforms
# ------------------------------------------------------------------------------------------------
def input_pipeline_numbers(pool_f32, B, S, dev, pk, measure_cpu, gan=None):
    """Augmented batches of B images from an 8-bit pool in HBM: kernel alone (CUDA events; 1 byte read + 4 bytes
    written per pixel against the measured HBM peak) and through the loader (host parameter sampling +
    sg_augment_params + one 56 B/image upload + kernel). CPU side: the reference's own transform stack
    (data_loader_signatures.py:154-219: torchvision on PIL images), one process, bounded sample."""
    import numpy as np
    import torch
    from device_data_loader import DeviceSignatureLoader
    imgs = ((pool_f32.reshape(-1, S, S) + 1.0) * 127.5).round().clamp(0, 255).to(torch.uint8)
    ld = DeviceSignatureLoader(imgs, batch_size=B, device=dev, seed=5)
    n = imgs.shape[0]
    idx = torch.randperm(n, device=dev)[:B].to(torch.int32)
    angles, scales, _ = ld.sample_params(B)
    for _ in range(3):
        ld.batch(idx, angles, scales)
    torch.cuda.synchronize()
    # kernel alone: same tables every launch, the 80 MB it touches per launch are swept out of L2 by a 256 MB write
    flush = torch.empty(64 << 20, dtype=torch.float32, device=dev)
    lib, cur = ld.lib, torch.cuda.current_stream(dev).cuda_stream
    import ctypes as C
    rot = np.empty((B, 6), np.int32)
    sc = np.empty((B, 4), np.float64)
    lib.sg_augment_params(angles.ctypes.data, scales.ctypes.data, B, S, rot.ctypes.data, sc.ctypes.data)
    rot_d, sc_d = torch.from_numpy(rot).to(dev), torch.from_numpy(sc).to(dev)
    out = torch.empty(B, 1, S, S, device=dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ms, iters = 0.0, 10
    for _ in range(iters):
        flush.zero_()
        e0.record()
        rc = lib.sg_augment_batch(imgs.data_ptr(), idx.data_ptr(), rot_d.data_ptr(), sc_d.data_ptr(), None, B, S,
                                  out.data_ptr(), cur)
        e1.record()
        torch.cuda.synchronize()
        assert rc == 0
        ms += e0.elapsed_time(e1)
    ms /= iters
    gbs = 5.0 * S * S * B / (ms * 1e-3) / 1e9
    # through the loader
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    reps = 20
    for _ in range(reps):
        ld.batch(idx)
    torch.cuda.synchronize()
    loader_ips = B * reps / (time.perf_counter() - t0)
    # training fed by the loader: epochs over the pool (shuffled, augmented), each batch straight into the fused step
    train_ips = None
    if gan is not None:
        ld2 = DeviceSignatureLoader(imgs, batch_size=B, device=dev, seed=6)
        for real in ld2:
            gan.train_step_async(real)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        seen = 0
        for _ in range(3):
            for real in ld2:
                gan.train_step_async(real)
                seen += real.shape[0]
        torch.cuda.synchronize()
        train_ips = seen / (time.perf_counter() - t0)
    res = {"kernel": "sg_augment_batch", "batch": B, "train_images_per_s_fed_by_loader": train_ips, "ms_per_batch": ms, "images_per_s_kernel": B / (ms * 1e-3),
           "roofline": {"bound": "hbm", "achieved": gbs, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": gbs / pk["hbm_gbs"],
                        "bytes_per_image": 5 * S * S, "l2": "256 MB flush between launches"},
           "images_per_s_loader": loader_ips,
           "loader_api": "DeviceSignatureLoader.batch(index): host sampling + sg_augment_params + 56 B/image H2D + kernel"}
    if measure_cpu:
        try:
            from PIL import Image
            from torchvision import transforms
            tf = transforms.Compose([transforms.Resize((S, S)), transforms.RandomRotation(degrees=5.0, fill=255),
                                     transforms.RandomAffine(degrees=0, scale=(0.9, 1.1), fill=255),
                                     transforms.ToTensor(), transforms.Normalize(mean=[0.5], std=[0.5])])
            host = imgs[:512].cpu().numpy()
            pils = [Image.fromarray(h) for h in host]
            t0 = time.perf_counter()
            count = 0
            while time.perf_counter() - t0 < 5.0:
                for pimg in pils:
                    tf(pimg)
                count += len(pils)
            res["cpu_baseline"] = {"value": count / (time.perf_counter() - t0), "unit": "images/s", "cores": 1,
                                   "kind": "reference",
                                   "sample": "torchvision/PIL transform stack of get_train_transforms on decoded 8-bit "
                                             f"{S}x{S} images, one process, ~5 s (the reference runs 4 such workers)"}
        except Exception as e:  # torchvision / PIL missing on the box
            res["cpu_baseline"] = {"unavailable": str(e)[:120]}
    return res


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def lib_nccl_version():
    import _siggan_lib as L
    v = int(L.load_library().sg_comm_nccl_version())
    return "unavailable" if v < 0 else f"{v // 10000}.{v // 100 % 100}.{v % 100}"


def run_ours(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    import _siggan_lib as L
    from vanilla_gan_model import VanillaGAN

    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device (the B200 path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, S = args.batch, args.size
    torch.manual_seed(1234 + rank)
    gan = VanillaGAN(latent_dim=100, image_size=S, device=str(dev), **({"width_mult": args.width} if args.width != 1 else {}))
    gan._fused_ready()
    flop_train, flop_sample = flops_per_image(S, args.width)
    assert args.width != 1 or abs(flop_train / FLOP_PER_IMG_TRAIN[S] - 1) < 1e-3, (flop_train, FLOP_PER_IMG_TRAIN[S])
    from data_parallel import broadcast_replica_
    broadcast_replica_([gan.generator._flat.flat, gan.generator._flat.stats, gan.discriminator._flat.flat])  # identical replicas
    comm = "none"
    if world > 1:
        if args.torch_allreduce:
            comm = "torch.distributed (NCCL process group)"
        else:
            from data_parallel import init_library_comm
            init_library_comm(gan)          # libsiggan's own NCCL communicator: all-reduces issued inside sg_train_step
            comm = f"libsiggan-owned NCCL communicator (nccl {lib_nccl_version()})"
    sync_bn = bool(args.sync_bn and world > 1)
    if sync_bn:
        from data_parallel import enable_sync_batchnorm
        enable_sync_batchnorm(gan.generator)
    lib = L.load_library()
    n_pool = 4     # 4 distinct real batches (268 MB fp32 at B=4096) > 126 MB L2; activations per step are several GB
    pool = synthetic_signatures(n_pool * B, S, dev, seed=1234 + rank).view(n_pool, B, 1, S, S)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timed region --------------------------------------------------------------
    for i in range(args.warmup):
        gan.train_step_async(pool[i % n_pool])
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    launches0 = lib.sg_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        metrics = gan.train_step_async(pool[i % n_pool])
    e1.record()
    barrier()
    launches = lib.sg_launch_count() - launches0
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms)
    last_metrics = metrics.tolist()
    value = world * B * args.steps / (ms_total * 1e-3)
    # ---- replicas must still be bit-identical after the timed steps (every rank applied the same averaged gradients
    # to the same parameters): exact integer checksums of the raw fp32 bit patterns of the flat G / D buffers, gathered
    replicas_identical = None
    if world > 1:
        def bits_checksum(t):
            w = t.detach().contiguous().view(torch.int32).to(torch.int64)
            idx = torch.arange(1, w.numel() + 1, device=w.device, dtype=torch.int64)
            return torch.stack([w.sum(), (w * (idx % 65521)).sum()])
        mine = torch.cat([bits_checksum(gan.generator._flat.flat), bits_checksum(gan.discriminator._flat.flat)])
        allc = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allc, mine)
        replicas_identical = all(bool(torch.equal(c, allc[0])) for c in allc)
        assert replicas_identical, "data-parallel replicas diverged: " + str([c.tolist() for c in allc])

    # ---- end to end: host-resident (pinned) real batches, H2D copy + D2H metric read every step ----
    host_pool = torch.empty(n_pool, B, 1, S, S, dtype=torch.float32).pin_memory()
    host_pool.copy_(pool)
    copy_stream = torch.cuda.Stream(dev)
    bufs = [torch.empty(B, 1, S, S, device=dev), torch.empty(B, 1, S, S, device=dev)]
    ready = [torch.cuda.Event(), torch.cuda.Event()]

    def prefetch(i):
        with torch.cuda.stream(copy_stream):
            bufs[i % 2].copy_(host_pool[i % n_pool], non_blocking=True)
            ready[i % 2].record(copy_stream)

    e2e_steps = max(3, min(args.steps, 20))
    for phase_steps, timed in ((2, False), (e2e_steps, True)):
        barrier()
        copy_stream.wait_stream(torch.cuda.current_stream(dev))
        prefetch(0)
        t0 = time.perf_counter()
        for i in range(phase_steps):
            torch.cuda.current_stream(dev).wait_event(ready[i % 2])
            if i + 1 < phase_steps:   # buffer (i+1)%2 was last read by step i-1, which has completed (train_step syncs)
                prefetch(i + 1)
            out = gan.train_step(bufs[i % 2])        # public API: returns python floats (device->host read)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    t = torch.tensor([dt], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * B * e2e_steps / float(t)
    clocks = sampler.stop() if sampler else None

    line = {
        "metric": "training images/sec (G+D step)", "value": value, "unit": "images/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": f"train_step_{S}x{S}_b{B}_per_gpu" + (f"_width{args.width}x" if args.width != 1 else ""),
                   "global_batch": world * B, "image_size": S, "width_mult": args.width,
                   "latent_dim": 100, "parallelism": f"dp{world}", "n_critic": 1, "sync_bn": sync_bn, "gradient_allreduce": comm,
                   "l2": "4 rotating real batches (268 MB) and multi-GB per-step activations exceed the 126 MB L2"},
        "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": B * S * S * 4, "d2h_bytes_per_step": 48,
                "steps": e2e_steps, "api": "VanillaGAN.train_step(real) from pinned host batches, double-buffered H2D"},
        "gpu_launches": int(launches),
        "replicas_identical": replicas_identical,
        "clocks": clocks,
        "last_metrics": {k: round(v, 5) for k, v in zip(
            ["d_loss", "d_loss_real", "d_loss_fake", "d_real_acc", "d_fake_acc", "d_real_mean", "d_fake_mean", "g_loss",
             "g_fake_mean"], last_metrics)},
    }
    pk = peaks()
    step_tflops = flop_train * value / world / 1e12
    line["step_tensor"] = {"achieved": step_tflops, "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s",
                           "frac": step_tflops / pk["bf16_tflops_sustained"], "peak_source": pk["source"],
                           "flop_per_image": flop_train}

    if not args.no_extras:
        sctx = gan.generator._ctx
        # ---- per-op device times (CUDA events inside the library, on the launch stream) ----------------
        prof_steps = 3
        barrier()
        sctx.profile(True)
        for i in range(prof_steps):
            gan.train_step_async(pool[i % n_pool])
        recs = sctx.profile_records()
        sctx.profile(False)
        agg = {}
        for name, ms_, fl, by in recs:
            a = agg.setdefault(name, [0.0, 0.0, 0.0, 0])
            a[0] += ms_; a[1] += fl; a[2] += by; a[3] += 1
        tot = sum(a[0] for a in agg.values())
        ops = sorted(agg.items(), key=lambda kv: -kv[1][0])
        tensor_ops = [(k, a) for k, a in ops if a[1] > 0 and ("wgrad" in k or "dgrad" in k or k[-1].isdigit() or k == "g.fc")
                      and not k.startswith("d.c0")]
        if tensor_ops:
            k, a = tensor_ops[0]
            ach = a[1] / (a[0] * 1e-3) / 1e12
            # DRAM traffic per launch from the committed ncu --set full capture of this op's kernel (per image x the
            # images of an average launch of the op: the D step runs it on 2B images, the G step on B)
            traffic, traffic_src = None, None
            try:
                with open(os.path.join(ROOT, "profiles", "r02_roofline_traffic.json")) as f:
                    t = json.load(f).get(k)
                if t and S == 64:
                    per_img = (t["dram_read_bytes"] + t["dram_write_bytes"]) / t["images"]
                    flops_per_img = {"d.c1.dgrad": 2.0 * 256 * 16 * 64 * 128}.get(k)
                    imgs_per_launch = (a[1] / a[3]) / flops_per_img if flops_per_img else B
                    traffic, traffic_src = per_img * imgs_per_launch, t["source"]
            except Exception:
                pass
            line["roofline"] = {"bound": "tensor", "kernel": k, "achieved": ach, "peak": pk["bf16_tflops_sustained"],
                                "unit": "TFLOP/s", "frac": ach / pk["bf16_tflops_sustained"], "traffic": traffic,
                                "traffic_source": traffic_src, "ms_per_launch": a[0] / a[3], "share_of_step": a[0] / tot,
                                "peak_source": pk["source"]}
        line["ops_ms_per_step"] = {k: round(a[0] / prof_steps, 4) for k, a in ops}
        # per op: achieved TFLOP/s on its algorithmic flops and GB/s on its algorithmic bytes
        line["ops_detail"] = {k: {"tflops": round(a[1] / (a[0] * 1e-3) / 1e12, 1) if a[1] else None,
                                  "gbs": round(a[2] / (a[0] * 1e-3) / 1e9, 0) if a[2] else None,
                                  "launches_per_step": a[3] / prof_steps} for k, a in ops if a[0] > 0}
        line["ops_total_ms_per_step"] = tot / prof_steps
        tens = sum(a[0] for k, a in tensor_ops)
        tfl = sum(a[1] for k, a in tensor_ops)
        line["tensor_ops"] = {"share_of_step": tens / tot, "achieved_tflops": tfl / (tens * 1e-3) / 1e12 if tens else None}

        # ---- sampling (BASELINE.json configs[1]): generator only, eval mode ----------------------------
        if rank == 0:
            SB = args.sampling_batch
            G = gan.generator
            G.eval()
            gz = torch.Generator(device=dev).manual_seed(0)
            z = torch.randn(SB, 100, device=dev, generator=gz)
            with torch.no_grad():
                for _ in range(3):
                    G(z)
                torch.cuda.synchronize()
                e0.record()
                iters = 10
                for _ in range(iters):
                    G(z)
                e1.record()
                torch.cuda.synchronize()
                samp_ms = e0.elapsed_time(e1) / iters
                for _ in range(2):
                    G.sample_uint8(z)
                torch.cuda.synchronize()
                e0.record()
                for _ in range(iters):
                    G.sample_uint8(z)
                e1.record()
                torch.cuda.synchronize()
                samp8_ms = e0.elapsed_time(e1) / iters
                # end to end: latents from pinned host memory, uint8 images back in pinned host memory, through the
                # public bulk sampler (device->host copy of chunk i overlapped with the generator on chunk i+1)
                zh = torch.randn(8 * SB, 100).pin_memory()
                out_h = torch.empty(8 * SB, 1, S, S, dtype=torch.uint8, pin_memory=True)
                G.sample_uint8_to_host(2 * SB, batch=SB, latents=zh[:2 * SB], out=out_h[:2 * SB])
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                G.sample_uint8_to_host(8 * SB, batch=SB, latents=zh, out=out_h)
                samp_e2e = 8 * SB / (time.perf_counter() - t0)
            sf = flop_sample * SB / (samp_ms * 1e-3) / 1e12
            line["sampling"] = {"batch": SB, "images_per_s_fp32_out": SB / (samp_ms * 1e-3),
                                "images_per_s_uint8_out": SB / (samp8_ms * 1e-3), "ms_fp32_out": samp_ms,
                                "achieved_tflops": sf, "tensor_frac": sf / pk["bf16_tflops"],
                                # layered bf16 activations with the tail fused: z + fc + 3 ConvT levels
                                # read and written once + the fp32 image
                                "bytes_per_image": SAMPLE_BYTES_PER_IMG.get(S) if args.width == 1 else None,
                                "hbm_frac": (SAMPLE_BYTES_PER_IMG[S] * SB / (samp_ms * 1e-3) / 1e9 / pk["hbm_gbs"])
                                if S in SAMPLE_BYTES_PER_IMG and args.width == 1 else None,
                                "e2e_images_per_s_uint8_host": samp_e2e,
                                "e2e_bytes": {"h2d": SB * 400, "d2h": SB * S * S},
                                "e2e_api": "Generator.sample_uint8_to_host (8 chunks, double-buffered D2H)"}
        # ---- the reference trainer's own loop over the modules (7 .item() per step) ------------------------
        if rank == 0 and world == 1:
            line["trainer_style"] = trainer_style_numbers(gan, pool, B, dev)
        # ---- input pipeline: augmentation kernel over a device-resident uint8 pool ------
        if rank == 0:
            line["input_pipeline"] = input_pipeline_numbers(pool, B, S, dev, pk, measure_cpu=(world == 1),
                                                             gan=gan if world == 1 else None)
        # ---- CPU baseline (rank 0, N = 1 only): the unmodified reference on the host cores, bounded sample ----
        if rank == 0 and world == 1 and args.width != 1:
            line["cpu_baseline"] = {"unavailable": "the 2x-width variant is not a configuration of the reference "
                                                   "(its base_features argument is inert)"}
        elif rank == 0 and world == 1:
            try:
                r = cpu_baseline_subprocess(S, args.cpu_batch, 12, 2, sampling=True)
                line["cpu_baseline"] = {"value": r["train_images_per_s"], "unit": "images/s", "cores": r["cores"],
                                        "kind": r["kind"], "sample": r["sample"]}
                if "sampling" in r and "sampling" in line:
                    line["sampling"]["cpu_baseline"] = r["sampling"]
            except Exception as e:
                line["cpu_baseline"] = {"unavailable": str(e)[:200]}
    if rank == 0:
        print(json.dumps(line), file=RESULT_OUT, flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


RESULT_OUT = sys.stdout


def claim_stdout():
    """stdout must carry exactly ONE JSON line. Libraries write to the C-level stdout behind Python's back (NCCL prints
    its version banner there under NCCL_DEBUG=VERSION — setting NCCL_DEBUG_FILE did not stop it on the box), so the
    process keeps a private duplicate of the original stdout for the result line and points fd 1 at stderr for
    everything else."""
    global RESULT_OUT
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    RESULT_OUT = os.fdopen(saved, "w")


def main():
    claim_stdout()
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.cpu_child:
        print(json.dumps(cpu_reference_numbers(args.size, args.cpu_batch, args.steps, args.warmup,
                                               sampling=not args.no_extras)), file=RESULT_OUT, flush=True)
        return
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world == 1 and args.gpus > 1:
        # launched without torchrun: re-exec under the launcher so that there is one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29511", os.path.abspath(__file__)] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd, stdout=RESULT_OUT))   # the ranks write their line to OUR original stdout
    run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
