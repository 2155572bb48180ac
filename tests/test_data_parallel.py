"""World-size-2 `gloo` tests of the data-parallel host logic on CPU: gradient-bucket averaging, replica
broadcast, batch sharding — and the property the whole scheme rests on, checked with the oracle: the mean of the
per-rank Discriminator-step gradients IS the global-batch gradient, and replicas that apply it stay bit-identical."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, out_dir: str) -> None:
    for sub in ("oracle", "signature-gan_b200"):
        sys.path.insert(0, os.path.join(ROOT, sub))
    import siggan_oracle as O
    from data_parallel import average_gradients_, broadcast_replica_, shard_range, world as dp_world

    torch.set_num_threads(2)
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        assert dp_world() == (rank, world)
        # ---- bucket averaging and replica broadcast
        g = torch.full((1000,), float(rank + 1))
        average_gradients_(g)
        assert torch.allclose(g, torch.full((1000,), (1 + world) / 2))
        p = torch.full((7,), float(rank))
        broadcast_replica_([p])
        assert torch.equal(p, torch.zeros(7))
        # ---- the step shards by batch: mean of per-rank D-step gradients == global-batch gradient
        size, B = 64, 8
        g_sd, d_sd = O.make_state_dicts(size, 100, seed=3)
        real = O.synthetic_signatures(B, size, seed=5)
        noise = O.hash_normal((B, 100), 77)
        masks_r, masks_f = O.make_dropout_masks(B, size, 1), O.make_dropout_masks(B, size, 2)
        names = O.trainable_names(d_sd)
        lo, hi = shard_range(B, rank, world)
        opt = O.AdamState(d_sd, names)
        _, grads, _ = O.d_step(g_sd, d_sd, opt, real[lo:hi], noise[lo:hi], size, [m[lo:hi] for m in masks_r],
                               [m[lo:hi] for m in masks_f], apply_update=False)
        flat = torch.cat([grads[k].reshape(-1) for k in names])
        average_gradients_(flat)
        _, ref, _ = O.d_step(g_sd, d_sd, O.AdamState(d_sd, names), real, noise, size, masks_r, masks_f,
                             apply_update=False)
        ref_flat = torch.cat([ref[k].reshape(-1) for k in names])
        err = (flat - ref_flat).norm() / ref_flat.norm()
        assert err < 1e-5, f"rank {rank}: sharded D gradient differs from the global-batch gradient by {err:.2e}"
        # ---- replicas that apply the averaged bucket stay bit-identical
        off = 0
        avg = {}
        for k in names:
            n = grads[k].numel()
            avg[k] = flat[off:off + n].view_as(grads[k])
            off += n
        opt.apply(d_sd, avg, 2e-4, 0.5, 0.999)
        mine = torch.cat([d_sd[k].reshape(-1) for k in names])
        both = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(both, mine)
        assert torch.equal(both[0], both[1])
        open(os.path.join(out_dir, f"ok{rank}"), "w").close()
    finally:
        dist.destroy_process_group()


def test_two_rank_gradient_averaging(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))


def test_shard_range_covers_everything():
    sys.path.insert(0, os.path.join(ROOT, "signature-gan_b200"))
    from data_parallel import shard_range
    for n, w in ((4096, 8), (10, 4), (3, 8), (0, 2)):
        spans = [shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [e - b for b, e in spans]
        assert max(sizes) - min(sizes) <= 1
