"""world_size-2 gloo tests (CPU) of the data-parallel host logic: GradSync all-reduce + 1/world scale
reproduces the single-process step on the concatenated batch (oracle/cyclegan_standin.py:352)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from unpaired_image_generation_b200.parallel import GradSync


def _flat_grads(nets):
    return torch.cat([p.grad.reshape(-1) for n in nets for p in n.parameters()])


def _worker(rank, world, port, tmp):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(2)
        from oracle import cyclegan_standin as ref
        sync = GradSync()
        assert sync.world_size == world and sync.rank == rank and sync.grad_scale == 1.0 / world
        # 1. all-reduce of a flat buffer range by range (the trainer reduces each gradient bucket as the step graph
        #    announces it), and broadcast of rank 0's state to ranks that were initialised differently
        flat = torch.arange(1000, dtype=torch.float32) * (rank + 1)
        for off, n in ((700, 300), (256, 444), (0, 256)):  # buckets arrive head-first, ranges are disjoint
            sync.all_reduce_(flat[off:off + n])
        assert torch.equal(flat, torch.arange(1000, dtype=torch.float32) * 3)
        state = torch.full((17,), float(rank + 5))  # "different seeds": every rank starts from its own values
        sync.broadcast_(state)
        assert torch.equal(state, torch.full((17,), 5.0))
        # 2. each rank: backward on its shard of a global batch of 2
        real_A, real_B = ref.synthetic_pair(2, 32, seed=7)
        sl = sync.shard_batch(2)
        nets = ref.build_models(seed=0, n_blocks=2)
        tr = ref.CycleGANTrainer(*nets)
        tr.backward_only(real_A[sl], real_B[sl])
        gG = _flat_grads(nets[:2])
        gD = _flat_grads(nets[2:])
        sync.all_reduce_(gG)
        sync.all_reduce_(gD)
        gG *= sync.grad_scale
        gD *= sync.grad_scale
        if rank == 0:
            full = ref.build_models(seed=0, n_blocks=2)
            trf = ref.CycleGANTrainer(*full)
            trf.backward_only(real_A, real_B)
            fG, fD = _flat_grads(full[:2]), _flat_grads(full[2:])
            relG = float((gG - fG).norm() / fG.norm())
            relD = float((gD - fD).norm() / fD.norm())
            torch.save({"relG": relG, "relD": relD}, tmp)
    finally:
        dist.destroy_process_group()


def test_gradsync_two_ranks_equals_global_batch(tmp_path):
    port = 29500 + os.getpid() % 2000
    out = str(tmp_path / "res.pt")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    res = torch.load(out)
    assert res["relG"] < 5e-3 and res["relD"] < 5e-3, res


def test_gradsync_single_process_is_identity():
    s = GradSync()
    t = torch.randn(10)
    assert s.world_size == 1 and s.grad_scale == 1.0
    assert s.all_reduce_(t.clone()).equal(t)
    assert s.shard_batch(4) == slice(0, 4)
    with pytest.raises(ValueError):
        GradSync.shard_batch(type("X", (), {"world_size": 3, "rank": 0})(), 4)
