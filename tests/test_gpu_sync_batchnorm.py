"""Synchronised BatchNorm: W ranks x B/W images with sg_set_sync_batchnorm reproduce ONE process at
B images — Generator output, running statistics and, after the gradient bucket is averaged, every gradient of the
G step (reference vanilla…:254-306), checked against the CPU oracle run on the GLOBAL batch.

Two processes share cuda:0 and talk over gloo (NCCL refuses two ranks on one device); on the 8-GPU box the same
callback carries an NCCL all-reduce. The library calls back into Python from inside sg_g_forward / sg_g_backward /
sg_train_step, so this also covers the callback plumbing."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, out_dir: str, precision: str, size: int, B: int) -> None:
    for sub in ("oracle", "signature-gan_b200", "tests"):
        sys.path.insert(0, os.path.join(ROOT, sub))
    import siggan_oracle as O
    from _util import check_grads, make_gan, rel_err, to64, tol
    from data_parallel import average_gradients_, enable_sync_batchnorm, shard_range

    torch.set_num_threads(4)
    torch.cuda.set_device(0)
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        gan, g_sd, d_sd = make_gan(size, 1, precision)
        G, D = gan.generator, gan.discriminator
        z = O.hash_normal((B, 100), 11)
        lo, hi = shard_range(B, rank, world)
        # ---- reference: ONE process at the global batch (fp32 and fp64 oracle)
        refs = []
        for gs, ds, zz in ((dict(g_sd), d_sd, z), (to64(g_sd), to64(d_sd), z.double())):
            _, grads, aux = O.g_step(gs, ds, O.AdamState(gs, O.trainable_names(gs)), zz, size, apply_update=False)
            refs.append((grads, aux, gs))
        img_ref = refs[1][1]["fake"]
        # ---- local statistics must NOT reproduce it (the test would be vacuous otherwise) ...
        G.train()
        D.eval()
        with torch.no_grad():
            local = G(z[lo:hi].cuda())
        assert rel_err(local, img_ref[lo:hi]) > 10 * tol(precision)
        G.load_state_dict(g_sd)                        # undo the running-stat update of that forward
        # ---- ... and synchronised statistics must
        enable_sync_batchnorm(G)
        fake = G(z[lo:hi].cuda())
        assert rel_err(fake, img_ref[lo:hi]) <= tol(precision), rel_err(fake, img_ref[lo:hi])
        for k, v in G.state_dict().items():
            if "running" in k:
                assert rel_err(v, refs[1][2][k]) <= tol(precision), (k, rel_err(v, refs[1][2][k]))
        pred = D(fake)
        loss = gan.criterion(pred, torch.ones(hi - lo, 1, device="cuda"))
        loss.backward()
        flat = G._flat.flat_grad_if_contiguous()
        assert flat is not None
        average_gradients_(flat)                       # what the data-parallel step does with the bucket
        got = {k: p.grad for k, p in G.named_parameters()}
        assert G.fc[0].bias.grad.abs().max() <= 1e-2 * G.fc[0].weight.grad.abs().max()
        check_grads(precision, "G", got, refs[0][0], refs[1][0])
        # ---- fused G step (sg_train_step phases 3 / 4) on the shards == oracle step on the global batch
        gan2, _, _ = make_gan(size, 1, precision)      # same library context: SyncBN stays on
        gan2.generator_step_async(hi - lo, noise=z[lo:hi].cuda())
        torch.cuda.synchronize()
        gs = {k: v.clone() for k, v in g_sd.items()}   # the oracle's Adam updates its tensors in place
        O.g_step(gs, d_sd, O.AdamState(gs, O.trainable_names(gs)), z, size)
        for k, v in gan2.generator.state_dict().items():
            if "num_batches" in k:
                continue
            kind = "act" if "running" in k else "param"
            assert rel_err(v, gs[k]) <= tol(precision, kind), (k, rel_err(v, gs[k]))
        # ---- switching it off restores local statistics
        enable_sync_batchnorm(G, enable=False)
        G.load_state_dict(g_sd)
        with torch.no_grad():
            again = G(z[lo:hi].cuda())
        assert torch.equal(again, local)
        open(os.path.join(out_dir, f"ok{rank}"), "w").close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("precision,size,B", [("fp32", 64, 16), ("bf16", 64, 64), ("fp32", 128, 8)])
def test_two_ranks_with_sync_batchnorm_equal_one_process(tmp_path, precision, size, B):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path), precision, size, B), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))
