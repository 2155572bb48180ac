"""Multi-GPU check of the library-owned NCCL communicator (run under torchrun, one rank per GPU; launched by
tests/test_gpu_dp_comm.py when the box has >= 2 GPUs, and by hand: `gpurun --gpus 2 -- python -m torch.distributed.run
--nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/run_dp_library_comm.py`).

 1. sg_allreduce_grads == the mean of the per-rank buckets gathered through torch.distributed (both halves: in order and
    overlapped on the communication stream).
 2. One data-parallel step as ONE library call (sg_train_step phase 0 with a communicator) == the same step with the
    bucket all-reduces done by torch.distributed between the phases: parameters bit-identical after 4 steps.
 3. The replicas stay bit-identical, and differ from a run without gradient averaging (the all-reduce is not a no-op).
Rank 0 prints one line `RESULT {json}`.
"""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "signature-gan_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def main():
    import _siggan_lib as L
    import data_parallel as dp
    import siggan_oracle as O
    from vanilla_gan_model import VanillaGAN

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    out = {"world": world}
    size, B = 64, 64
    g_sd, d_sd = O.make_state_dicts(size, 100, seed=3)

    def make():
        torch.manual_seed(100 + rank)
        gan = VanillaGAN(latent_dim=100, image_size=size, device=str(dev))
        gan.generator.load_state_dict(g_sd)
        gan.discriminator.load_state_dict(d_sd)
        gan._fused_ready()
        return gan

    real = [O.synthetic_signatures(B, size, seed=40 + 7 * rank + i).to(dev) for i in range(4)]

    # ---- 1. the collective itself
    gan = make()
    sctx = gan._fused_ready()
    assert dp.init_library_comm(gan) == world and sctx.comm_world() == world
    n = sctx.param_count(L.SG_NET_D)
    torch.manual_seed(7 + rank)
    bucket = torch.randn(n, device=dev)
    gathered = [torch.empty_like(bucket) for _ in range(world)]
    dist.all_gather(gathered, bucket)
    want = torch.stack(gathered).double().mean(0)
    tail = int(sctx.lib.sg_d_grad_tail_offset(sctx.handle))
    sctx.allreduce_grads(L.SG_NET_D, bucket, tail, -1, overlap=True)
    sctx.allreduce_grads(L.SG_NET_D, bucket, 0, tail, overlap=False)
    sctx.allreduce_join()
    torch.cuda.synchronize()
    out["allreduce_max_err"] = float((bucket.double() - want).abs().max())

    # ---- 2. whole step in the library vs. torch.distributed between the phases
    def run(mode):
        torch.manual_seed(100 + rank)
        L.DROPOUT.offset = 0
        g = make()
        ctx = g._fused_ready()
        if mode == "library":
            dp.init_library_comm(g)
        else:
            L.check(ctx.lib.sg_comm_destroy(ctx.handle), "comm destroy")
        for i in range(4):
            if mode == "none":      # no averaging at all: every rank trains on its own shard only
                saved = dp.world
                dp.world = lambda: (0, 1)
                try:
                    g.train_step_async(real[i])
                finally:
                    dp.world = saved
            else:
                g.train_step_async(real[i])
        torch.cuda.synchronize()
        return torch.cat([g.generator._flat.flat, g.discriminator._flat.flat, g.generator._flat.stats]).clone()

    p_lib, p_torch, p_none = run("library"), run("torch"), run("none")
    out["library_equals_torch_path"] = bool(torch.equal(p_lib, p_torch))
    out["library_vs_torch_max_abs"] = float((p_lib - p_torch).abs().max())
    # ---- 3. replicas
    def identical(p):
        allp = [torch.empty_like(p) for _ in range(world)]
        dist.all_gather(allp, p)
        return all(bool(torch.equal(a, allp[0])) for a in allp)

    n_params = p_lib.numel() - gan.generator._flat.stats.numel()      # BatchNorm running statistics stay local
    out["replicas_identical_library"] = identical(p_lib[:n_params])
    out["replicas_identical_torch"] = identical(p_torch[:n_params])
    out["replicas_identical_without_allreduce"] = identical(p_none[:n_params])
    out["finite"] = bool(torch.isfinite(p_lib).all())
    if rank == 0:
        print("RESULT " + json.dumps(out), flush=True)
    ok = (out["allreduce_max_err"] < 1e-6 and out["library_equals_torch_path"] and out["replicas_identical_library"]
          and out["replicas_identical_torch"] and not out["replicas_identical_without_allreduce"] and out["finite"])
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
