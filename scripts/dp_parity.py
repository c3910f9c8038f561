"""Data-parallel parity on W GPUs (torchrun, NCCL), also run by tests/test_gpu_dp.py:

  1. fp32 validation mode: W ranks x batch 1, gradients all-reduced and scaled by 1/W, equal the single-process batch-W
     step <= 1e-5 (relative L2 per group) -- the product's data-parallel plumbing below the bf16 noise floor;
  2. bf16 product path, overlapped bucketed all-reduce: ranks that were constructed with DIFFERENT seeds start from
     rank 0's weights (broadcast) and hold bit-identical weights after 3 steps (an all-reduce that ran before a
     bucket was final would leave rank-local contributions behind and the ranks would drift apart);
  3. the same gradients by direction against the single-process batch-W step in bf16 (noise-limited: cos > 0.95).
"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import unpaired_image_generation_b200 as cgb  # noqa: E402


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())


def grads_vs_global(world, rank, real_A, real_B, precision):
    mk = lambda: (cgb.Generator(seed=1), cgb.Generator(seed=2), cgb.Discriminator(seed=3), cgb.Discriminator(seed=4))
    tr = cgb.CycleGANTrainer(*mk(), precision=precision)
    sl = tr.sync.shard_batch(world)
    a, b = real_A[sl].cuda(), real_B[sl].cuda()
    eng = tr._ensure_engine(a)
    with torch.cuda.stream(tr.stream):
        eng.set_inputs(a, b)
        eng.phase_generators()
        eng.phase_discriminators()
        tr.sync.all_reduce_(eng.grads[0])
        tr.sync.all_reduce_(eng.grads[1])
    torch.cuda.synchronize()
    gG, gD = eng.grads[0] * tr.sync.grad_scale, eng.grads[1] * tr.sync.grad_scale
    out = None
    if rank == 0:
        full = cgb.CycleGANTrainer(*mk(), precision=precision)
        full.sync.world_size, full.sync.enabled = 1, False  # single-process reference on the concatenated batch
        full.backward_only(real_A.cuda(), real_B.cuda())
        fG, fD = full.engine.grads[0], full.engine.grads[1]
        cos = lambda x, y: float((x.double() * y.double()).sum() / (x.double().norm() * y.double().norm()))
        out = dict(relG=rel(gG, fG), relD=rel(gD, fD), cosG=cos(gG, fG), cosD=cos(gD, fD))
    return out


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
    size = int(os.environ.get("DP_PARITY_SIZE", "64"))
    g = torch.Generator().manual_seed(5)
    real_A = torch.rand(world, 3, size, size, generator=g) * 2 - 1
    real_B = torch.rand(world, 3, size, size, generator=g) * 2 - 1

    r32 = grads_vs_global(world, rank, real_A, real_B, "fp32")
    rbf = grads_vs_global(world, rank, real_A, real_B, "bf16")
    if rank == 0:
        print(f"DP parity fp32 mode ({world} ranks x batch 1 vs batch {world}): rel G {r32['relG']:.3e} D {r32['relD']:.3e}", flush=True)
        print(f"DP parity bf16 mode: cos G {rbf['cosG']:.5f} D {rbf['cosD']:.5f} rel G {rbf['relG']:.3e} D {rbf['relD']:.3e}", flush=True)
        assert r32["relG"] < 1e-5 and r32["relD"] < 1e-5, r32
        assert rbf["cosG"] > 0.95 and rbf["cosD"] > 0.98, rbf

    # different seeds per rank + overlapped bucketed all-reduce: identical weights on every rank after 3 steps
    for precision in ("bf16", "fp32"):
        s = 100 * (rank + 1)
        tr = cgb.CycleGANTrainer(cgb.Generator(seed=s), cgb.Generator(seed=s + 1), cgb.Discriminator(seed=s + 2),
                                 cgb.Discriminator(seed=s + 3), precision=precision)
        sl = tr.sync.shard_batch(world)
        a, b = real_A[sl].cuda(), real_B[sl].cuda()
        for _ in range(3):
            losses = tr.train_step(a, b)
        torch.cuda.synchronize()
        eng = tr.engine
        same = True
        for t in eng.params + eng.exp_avg + eng.exp_avg_sq:
            lo, hi = t.clone(), t.clone()
            dist.all_reduce(lo, op=dist.ReduceOp.MIN)
            dist.all_reduce(hi, op=dist.ReduceOp.MAX)
            same = same and bool(torch.equal(lo, hi))
        ref0 = cgb.Generator(seed=100)  # rank 0's initial G_AB: training must have started from it on every rank
        if rank == 0:
            print(f"[{precision}] buckets {[(b['group'], b['net'], b['layer_lo'], b['layer_hi'], b['numel']) for b in eng.grad_bucket_plan()]}", flush=True)
            print(f"[{precision}] weights and Adam state identical across {world} differently seeded ranks after 3 "
                  f"overlapped DP steps: {same}; losses {losses}", flush=True)
        assert same
        del ref0
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
