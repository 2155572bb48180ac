"""Data-parallel plumbing of the training step: one process per GPU, full replicas, the step shards by
batch. Losses are batch means, so the global-batch gradient of everything except BatchNorm statistics is the mean of
the per-rank gradients: each rank runs its local D / G backward into the flat fp32 gradient bucket of the network
(`_siggan_lib.FlatParams.grad_staging()`), the bucket is summed over ranks with ONE all-reduce (NCCL over
NVLink / NVSwitch on the GPU box, gloo in the CPU tests) and scaled by 1 / world_size before the fused Adam update.
BatchNorm statistics stay local by default (the DDP convention; at 4096 images per replica they match the reference
run at that batch); `enable_sync_batchnorm` switches the Generator's BatchNorm layers to global-batch statistics, so
that W ranks x B images reproduce ONE process at W*B images. The reference scripts are single-process (train…:494-502); multi-GPU runs are driven by bench.py / a launcher
using these helpers with the same modules.
"""
from __future__ import annotations

from typing import Iterable, Optional, Tuple

import torch
import torch.distributed as dist


def world() -> Tuple[int, int]:
    """(rank, world_size); (0, 1) outside a process group."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def _mean_op(group=None):
    """NCCL averages inside the collective (ncclAvg: no extra kernel); gloo has no AVG, there it is SUM + one scale."""
    try:
        if dist.get_backend(group) == "nccl":
            return dist.ReduceOp.AVG, False
    except Exception:
        pass
    return dist.ReduceOp.SUM, True


def average_gradients_(flat_grad: torch.Tensor, group: Optional["dist.ProcessGroup"] = None) -> torch.Tensor:
    """In place: flat_grad <- mean over ranks of flat_grad. A no-op for a single process."""
    if not (dist.is_available() and dist.is_initialized()):
        return flat_grad
    n = dist.get_world_size(group)
    if n > 1:
        op, scale = _mean_op(group)
        dist.all_reduce(flat_grad, op=op, group=group)
        if scale:
            flat_grad.mul_(1.0 / n)
    return flat_grad


def all_reduce_start(part: torch.Tensor, group: Optional["dist.ProcessGroup"] = None):
    """Start averaging `part` (a contiguous slice of a gradient bucket) over ranks without blocking the caller's stream:
    NCCL runs the collective on its own stream after the work already enqueued on the current stream, so kernels
    enqueued afterwards (the rest of the backward pass) overlap it. Returns a work handle for all_reduce_finish."""
    return dist.all_reduce(part, op=_mean_op(group)[0], group=group, async_op=True)


def all_reduce_finish(works, flat_grad: torch.Tensor, group: Optional["dist.ProcessGroup"] = None) -> torch.Tensor:
    """Make the current stream wait for the started reductions; the bucket then holds the mean over ranks."""
    for w in works:
        if w is not None:
            w.wait()
    if _mean_op(group)[1]:
        flat_grad.mul_(1.0 / dist.get_world_size(group))
    return flat_grad


def broadcast_replica_(tensors: Iterable[torch.Tensor], src: int = 0) -> None:
    """Make every rank start from rank `src`'s parameters / BatchNorm statistics (flat buffers, one broadcast each)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return
    for t in tensors:
        dist.broadcast(t, src)


def init_library_comm(module, group: Optional["dist.ProcessGroup"] = None) -> int:
    """Give the library context behind `module` (a siggan_b200 Generator / Discriminator / VanillaGAN on its CUDA
    device) its own NCCL communicator over `group` (sg_comm_init). From then on VanillaGAN's data-parallel steps let
    libsiggan issue the bucket all-reduces itself, on its communication stream, overlapped with the backward passes
    (sg_train_step phase 0 / sg_allreduce_grads) instead of going through torch.distributed. Returns the world size."""
    gen = getattr(module, "generator", module)
    dev = next(gen.parameters()).device
    gen._prepare(dev)
    return gen._ctx.init_comm(group)


def shard_range(n_items: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous [begin, end) share of `n_items` for `rank`; shares differ by at most one item."""
    base, extra = divmod(n_items, world_size)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def enable_sync_batchnorm(generator, group: Optional["dist.ProcessGroup"] = None, enable: bool = True) -> None:
    """Global-batch BatchNorm statistics for `generator` (a siggan_b200 Generator already on its CUDA device): the
    per-channel sums of every training-mode BatchNorm forward / backward are all-reduced over `group` from inside
    libsiggan (sg_set_sync_batchnorm; ten 2*C-float latency-bound collectives per G step). Affects every module
    that shares the generator's library context (same device / image size / precision)."""
    dev = next(generator.parameters()).device
    generator._prepare(dev)
    generator._ctx.set_sync_batchnorm(group, enable)
