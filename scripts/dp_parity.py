"""2-rank check (torchrun, NCCL): W ranks x batch 1 must equal the single-process batch-W step:
gradients after the all-reduce * 1/W vs a single-process engine on the concatenated batch."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import unpaired_image_generation_b200 as cgb  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
    size = 64
    g = torch.Generator().manual_seed(5)
    real_A = torch.rand(world, 3, size, size, generator=g) * 2 - 1
    real_B = torch.rand(world, 3, size, size, generator=g) * 2 - 1
    mk = lambda: (cgb.Generator(seed=1), cgb.Generator(seed=2), cgb.Discriminator(seed=3), cgb.Discriminator(seed=4))
    tr = cgb.CycleGANTrainer(*mk())
    sl = tr.sync.shard_batch(world)
    a, b = real_A[sl].cuda(), real_B[sl].cuda()
    # gradients: phases + all-reduce (no optimiser), compared with the global-batch engine on rank 0
    eng = tr._ensure_engine(a)
    with torch.cuda.stream(tr.stream):
        eng.set_inputs(a, b)
        eng.phase_generators()
        eng.phase_discriminators()
        tr.sync.all_reduce_(eng.grads[0])
        tr.sync.all_reduce_(eng.grads[1])
    torch.cuda.synchronize()
    gG = eng.grads[0].clone() * tr.sync.grad_scale
    gD = eng.grads[1].clone() * tr.sync.grad_scale
    if rank == 0:
        full = cgb.CycleGANTrainer(*mk(), process_group=None)
        full.sync.world_size = 1  # single-process reference on the concatenated batch
        full.sync.enabled = False
        full.backward_only(real_A.cuda(), real_B.cuda())
        fG, fD = full.engine.grads[0], full.engine.grads[1]
        cosG = float((gG * fG).sum() / (gG.norm() * fG.norm()))
        cosD = float((gD * fD).sum() / (gD.norm() * fD.norm()))
        print(f"DP parity ({world} ranks x batch 1 vs batch {world}): cos G {cosG:.5f} D {cosD:.5f}; "
              f"rel G {float((gG - fG).norm() / fG.norm()):.3e} D {float((gD - fD).norm() / fD.norm()):.3e}", flush=True)
        assert cosG > 0.95 and cosD > 0.98  # bf16 rounding noise is amplified to ~0.2 rel on G
    # and three full DP steps run and stay finite / identical across ranks
    for _ in range(3):
        losses = tr.train_step(a, b)
    chk = torch.tensor([float(eng.params[0].double().sum()), float(eng.params[1].double().sum())], device="cuda", dtype=torch.float64)
    lo, hi = chk.clone(), chk.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    if rank == 0:
        print("weights identical across ranks after 3 DP steps:", bool(torch.equal(lo, hi)), losses, flush=True)
        assert torch.equal(lo, hi)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
