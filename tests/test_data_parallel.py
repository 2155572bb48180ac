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


def _syncbn_worker(rank: int, world: int, port: int, out_dir: str) -> None:
    """The algebra sg_set_sync_batchnorm implements in CUDA, restated with the oracle's BatchNorm on two gloo ranks:
    all-reduce the [2][C] rows (sum x, sum x^2 | sum d, sum d*xhat), finalize with rows x world_size; BatchNorm weight /
    bias gradients stay local and are averaged with the bucket. Each rank's upstream gradient is that of ITS mean loss
    (world_size x the global-loss gradient of its samples), as in the data-parallel step."""
    for sub in ("oracle", "signature-gan_b200"):
        sys.path.insert(0, os.path.join(ROOT, sub))
    import siggan_oracle as O
    from data_parallel import average_gradients_, shard_range

    torch.set_num_threads(2)
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        B, C, H = 12, 16, 8
        y = O.hash_normal((B, C, H, H), 5).double() * 1.7 + 0.3
        d = O.hash_normal((B, C, H, H), 6).double()                 # d(global mean loss) / d(bn output)
        sd = {"bn.weight": O.hash_normal((C,), 7, 1.0, 0.1).double(), "bn.bias": O.hash_normal((C,), 8).double(),
              "bn.running_mean": torch.zeros(C, dtype=torch.float64), "bn.running_var": torch.ones(C, dtype=torch.float64),
              "bn.num_batches_tracked": torch.tensor(0)}
        stats = {}
        out_ref, xhat_ref, rstd_ref = O._bn_forward(y, sd, "bn", True, stats)
        dy_ref, dgamma_ref, dbeta_ref = O._bn_backward(d, xhat_ref, rstd_ref, sd["bn.weight"], True)
        lo, hi = shard_range(B, rank, world)
        yl, dl = y[lo:hi], d[lo:hi] * world
        rows = (hi - lo) * H * H
        # ---- forward: one [2][C] row per rank, summed over ranks, finalized with the global row count
        row = torch.stack([yl.sum(dim=(0, 2, 3)), (yl * yl).sum(dim=(0, 2, 3))])
        dist.all_reduce(row)
        n = rows * world
        mean = row[0] / n
        var = (row[1] / n - mean * mean).clamp_min(0)
        rstd = torch.rsqrt(var + O.BN_EPS)
        xhat = (yl - mean.view(1, -1, 1, 1)) * rstd.view(1, -1, 1, 1)
        out = xhat * sd["bn.weight"].view(1, -1, 1, 1) + sd["bn.bias"].view(1, -1, 1, 1)
        assert torch.allclose(out, out_ref[lo:hi], rtol=1e-10, atol=1e-10)
        run_var = 0.9 * sd["bn.running_var"] + 0.1 * var * n / (n - 1)
        assert torch.allclose(run_var, stats["bn.running_var"], rtol=1e-10, atol=1e-12)
        assert torch.allclose(0.1 * mean, stats["bn.running_mean"], rtol=1e-10, atol=1e-12)
        # ---- backward: local dgamma / dbeta (averaged like the rest of the bucket), global k2 / k3 for the data gradient
        brow = torch.stack([dl.sum(dim=(0, 2, 3)), (dl * xhat).sum(dim=(0, 2, 3))])
        dbeta, dgamma = brow[0].clone(), brow[1].clone()
        dist.all_reduce(brow)
        k1 = (sd["bn.weight"] * rstd).view(1, -1, 1, 1)
        dy = k1 * (dl - (brow[0] / n).view(1, -1, 1, 1) - xhat * (brow[1] / n).view(1, -1, 1, 1))
        assert torch.allclose(dy, world * dy_ref[lo:hi], rtol=1e-9, atol=1e-10)
        bucket = torch.cat([dgamma, dbeta])
        average_gradients_(bucket)
        assert torch.allclose(bucket, torch.cat([dgamma_ref, dbeta_ref]), rtol=1e-9, atol=1e-10)
        # ---- and local statistics would NOT have given the global result (what SyncBN is for)
        lm = yl.mean(dim=(0, 2, 3))
        assert (lm - mean).abs().max() > 1e-3
        open(os.path.join(out_dir, f"sbn{rank}"), "w").close()
    finally:
        dist.destroy_process_group()


def test_two_rank_sync_batchnorm_algebra(tmp_path):
    world = 2
    mp.spawn(_syncbn_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"sbn{r}").exists() for r in range(world))
