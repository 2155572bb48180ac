"""Asynchronous checkpoint writer in the reference trainer's format.

The reference's `GANTrainer._save_checkpoint` (train…:402-444) builds four state dicts and calls `torch.save` three times
(epoch file, `checkpoint_latest.pt`, optionally `checkpoint_best.pt`) on the training thread: ~45 MB of device->host
copies through pageable memory plus pickling, with the GPU idle. Here the training thread only ENQUEUES a
device-to-device snapshot of the flat buffers (G parameters, G BatchNorm statistics, D parameters, and the two Adam
moment pairs: seven copies, tens of microseconds) on its own stream and their device->pinned-host copies on a side
stream, and returns; a worker thread waits for the copies, rebuilds the four state dicts as views into the host snapshot (same keys, shapes, dtypes and key order as the
modules' own `state_dict()` / `torch.optim.Adam.state_dict()`), and writes the same files. Training continues — and may
already be changing the parameters — while the snapshot is copied and pickled; two host snapshots alternate, so a save
only waits when the save before the previous one is still being written.

    writer = AsyncCheckpointWriter(gan, checkpoint_dir)
    writer.save(epoch, global_step, config=cfg.to_dict(), fixed_noise=fixed_noise, best_g_loss=best, is_best=False)
    ...
    writer.wait()          # before reading the files / at the end of training

The files load with the reference's own `GANTrainer.load_checkpoint` (train…:446-486), with `utils/inference.load_generator`
and with `VanillaGAN.load`-style code, on either implementation.
"""
from __future__ import annotations

import queue
import threading
from collections import OrderedDict
from pathlib import Path
from typing import Any, Dict, List, Optional

import torch


class _HostSnapshot:
    """Pinned host copies of the flat device buffers of one VanillaGAN."""

    def __init__(self, sizes: Dict[str, int], n_bn: int) -> None:
        self.buf = {k: torch.empty(n, dtype=torch.float32).pin_memory() for k, n in sizes.items()}
        self.nbt = torch.empty(n_bn, dtype=torch.int64).pin_memory()
        self.ready = torch.cuda.Event()
        self.free = threading.Event()
        self.free.set()


class AsyncCheckpointWriter:
    def __init__(self, gan, checkpoint_dir, keep_epoch_files: bool = True) -> None:
        if gan.use_spectral_norm:
            raise NotImplementedError("AsyncCheckpointWriter snapshots the flat parameter buffers of the standard "
                                      "Discriminator; save the spectral-norm variant with VanillaGAN.save")
        self.gan = gan
        self.dir = Path(checkpoint_dir)
        self.dir.mkdir(parents=True, exist_ok=True)
        self.keep_epoch_files = keep_epoch_files
        gan._fused_ready()                      # flat buffers + Adam moments exist
        g, d = gan.generator, gan.discriminator
        self.device = g._flat.flat.device
        self._sizes = {"g": g._flat.flat.numel(), "g_stats": g._flat.stats.numel(), "d": d._flat.flat.numel(),
                       "g_m": g._flat.flat.numel(), "g_v": g._flat.flat.numel(), "d_m": d._flat.flat.numel(),
                       "d_v": d._flat.flat.numel()}
        self._bns = g._bn_modules()
        self._snaps = [_HostSnapshot(self._sizes, len(self._bns)) for _ in range(2)]
        self._stage = {k: torch.empty(n, dtype=torch.float32, device=self.device) for k, n in self._sizes.items()}
        self._stage_nbt = torch.empty(len(self._bns), dtype=torch.int64, device=self.device)
        self._staged, self._stage_drained = torch.cuda.Event(), torch.cuda.Event()
        self._stage_drained.record(torch.cuda.current_stream(self.device))
        self._g_keys = list(g.state_dict().keys())
        self._d_keys = list(d.state_dict().keys())
        self._turn = 0
        self._stream = torch.cuda.Stream(self.device)
        self._jobs: "queue.Queue" = queue.Queue()
        self._errors: List[BaseException] = []
        self._worker = threading.Thread(target=self._run, name="siggan-checkpoint-writer", daemon=True)
        self._worker.start()
        self.written: List[Path] = []

    # -- training thread -------------------------------------------------------------------------------------------
    def save(self, epoch: int, global_step: int, config: Optional[Dict[str, Any]] = None,
             fixed_noise: Optional[torch.Tensor] = None, best_g_loss: float = float("inf"), is_best: bool = False) -> Path:
        """Snapshot now (stream-ordered after everything enqueued so far), write in the background. Returns the path of
        the epoch file the worker will write."""
        self._raise_pending()
        gan = self.gan
        gan._fused_ready()
        g, d = gan.generator, gan.discriminator
        snap = self._snaps[self._turn]
        self._turn ^= 1
        snap.free.wait()                        # the save before the previous one must have left this snapshot
        snap.free.clear()
        src = {"g": g._flat.flat, "g_stats": g._flat.stats, "d": d._flat.flat, "g_m": gan.g_optimizer._m,
               "g_v": gan.g_optimizer._v, "d_m": gan.d_optimizer._m, "d_v": gan.d_optimizer._v}
        cur = torch.cuda.current_stream(self.device)
        # 1. device-side snapshot on the training stream (45 MB of device-to-device copies, tens of microseconds): it
        #    sees every update enqueued before this call and none enqueued after it, and training goes on at once
        cur.wait_event(self._stage_drained)     # the previous save's host copies have left the staging buffers
        for k, t in src.items():
            self._stage[k].copy_(t, non_blocking=True)
        self._stage_nbt.copy_(torch.stack([bn.num_batches_tracked for bn in self._bns]), non_blocking=True)
        self._staged.record(cur)
        # 2. device -> pinned host on the side stream, overlapping the training kernels
        with torch.cuda.stream(self._stream):
            self._stream.wait_event(self._staged)
            for k in src:
                snap.buf[k].copy_(self._stage[k], non_blocking=True)
            snap.nbt.copy_(self._stage_nbt, non_blocking=True)
            snap.ready.record(self._stream)
            self._stage_drained.record(self._stream)
        meta = {
            "epoch": int(epoch), "global_step": int(global_step), "config": dict(config or gan.get_config()),
            "fixed_noise": self._noise_on_host(fixed_noise),
            "best_g_loss": float(best_g_loss), "is_best": bool(is_best),
            "g_opt": self._opt_meta(gan.g_optimizer), "d_opt": self._opt_meta(gan.d_optimizer),
        }
        path = self.dir / f"checkpoint_epoch_{int(epoch):04d}.pt"
        self._jobs.put((snap, meta, path))
        return path

    def wait(self) -> None:
        """Block until every enqueued checkpoint is on disk; re-raises a worker failure."""
        self._jobs.join()
        self._raise_pending()

    def close(self) -> None:
        self.wait()
        self._jobs.put(None)
        self._worker.join(timeout=30)

    # -- helpers ---------------------------------------------------------------------------------------------------
    def _noise_on_host(self, fixed_noise: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
        """The trainer's fixed noise never changes (train…:197-201): one synchronising copy the first time it is seen."""
        if fixed_noise is None:
            return None
        key = (fixed_noise.data_ptr(), tuple(fixed_noise.shape), fixed_noise._version)
        if getattr(self, "_noise_key", None) != key:
            self._noise_key, self._noise_cpu = key, fixed_noise.detach().cpu()
        return self._noise_cpu

    @staticmethod
    def _opt_meta(opt) -> Dict[str, Any]:
        sd = torch.optim.Optimizer.state_dict(opt)          # param_groups (+ ids); the state tensors come from the snapshot
        return {"param_groups": sd["param_groups"], "step": float(opt._steps)}

    def _raise_pending(self) -> None:
        if self._errors:
            raise RuntimeError("checkpoint writer failed") from self._errors.pop(0)

    def _module_state(self, module, flat: torch.Tensor, stats: Optional[torch.Tensor], nbt: Optional[torch.Tensor]):
        """The module's state_dict (keys in its own order) as clones of slices of the host snapshot."""
        fp = module._flat
        where: Dict[str, torch.Tensor] = {}
        for name, (off, n, shape) in zip(fp._names, fp.layout):
            where[name] = flat[off:off + n].view(shape)
        if stats is not None:
            prefixes = ["fc.1"] + [f"upsample_blocks.{i}.block.1" for i in range(len(self._bns) - 1)]
            for k, (prefix, (mo, vo, ch)) in enumerate(zip(prefixes, module._ctx.bn_table())):
                where[prefix + ".running_mean"] = stats[mo:mo + ch]
                where[prefix + ".running_var"] = stats[vo:vo + ch]
                where[prefix + ".num_batches_tracked"] = nbt[k]
        out = OrderedDict()
        for key in (self._g_keys if stats is not None else self._d_keys):      # key order of the module's own state_dict
            out[key] = where[key].clone()
        return out

    def _opt_state(self, module, m: torch.Tensor, v: torch.Tensor, meta: Dict[str, Any]) -> Dict[str, Any]:
        fp = module._flat
        state = {}
        step = torch.tensor(meta["step"])
        if meta["step"] > 0:
            for i, (off, n, shape) in enumerate(fp.layout):
                state[i] = {"step": step.clone(), "exp_avg": m[off:off + n].view(shape).clone(),
                            "exp_avg_sq": v[off:off + n].view(shape).clone()}
        return {"state": state, "param_groups": meta["param_groups"]}

    # -- worker thread ---------------------------------------------------------------------------------------------
    def _run(self) -> None:
        while True:
            job = self._jobs.get()
            if job is None:
                self._jobs.task_done()
                return
            snap, meta, path = job
            try:
                snap.ready.synchronize()
                gan = self.gan
                ckpt = {
                    "epoch": meta["epoch"], "global_step": meta["global_step"],
                    "generator_state_dict": self._module_state(gan.generator, snap.buf["g"], snap.buf["g_stats"], snap.nbt),
                    "discriminator_state_dict": self._module_state(gan.discriminator, snap.buf["d"], None, None),
                    "g_optimizer_state_dict": self._opt_state(gan.generator, snap.buf["g_m"], snap.buf["g_v"], meta["g_opt"]),
                    "d_optimizer_state_dict": self._opt_state(gan.discriminator, snap.buf["d_m"], snap.buf["d_v"], meta["d_opt"]),
                    "config": meta["config"], "best_g_loss": meta["best_g_loss"],
                }
                if meta["fixed_noise"] is not None:
                    ckpt["fixed_noise"] = meta["fixed_noise"]
                snap.free.set()                      # everything was cloned out of the snapshot
                targets = ([path] if self.keep_epoch_files else []) + [self.dir / "checkpoint_latest.pt"]
                if meta["is_best"]:
                    targets.append(self.dir / "checkpoint_best.pt")
                for t in targets:                   # write-then-rename: a reader never sees a partial file
                    tmp = t.with_suffix(t.suffix + ".tmp")
                    torch.save(ckpt, tmp)
                    tmp.replace(t)
                    self.written.append(t)
            except BaseException as e:  # noqa: BLE001 — surfaced on the training thread by save() / wait()
                self._errors.append(e)
                snap.free.set()
            finally:
                self._jobs.task_done()
