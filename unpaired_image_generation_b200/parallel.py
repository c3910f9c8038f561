"""Data-parallel plumbing for the CycleGAN step.

Samples are independent (InstanceNorm is per-sample), so the only collective on the path is a
sum-allreduce of the two flat gradient buffers (generators: 22.8 M floats, discriminators: 5.5 M),
followed by a 1/world scale folded into the Adam kernel.  The trainer all-reduces them bucket by bucket
as the step graph announces each range final (see CycleGANTrainer._train_step_dp_nosync); this module
only wraps the collectives, so the same code runs over NCCL on B200s and over gloo in the CPU test-suite.
Stand-in counterpart: a single-process step on the concatenated batch
(oracle/cyclegan_standin.py:374 train_step) -- see tests/test_parallel_gloo.py.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist


class GradSync:
    def __init__(self, process_group: Optional["dist.ProcessGroup"] = None):
        self.group = process_group
        self.enabled = dist.is_available() and dist.is_initialized()
        self.world_size = dist.get_world_size(process_group) if self.enabled else 1
        self.rank = dist.get_rank(process_group) if self.enabled else 0

    @property
    def grad_scale(self) -> float:
        """factor the optimiser applies to the summed gradients (mean over ranks)"""
        return 1.0 / self.world_size

    def all_reduce_(self, flat: torch.Tensor) -> torch.Tensor:
        """in-place SUM over ranks of a contiguous gradient range (enqueued on the current stream)"""
        if self.world_size > 1:
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
        return flat

    def broadcast_(self, t: torch.Tensor, src: int = 0) -> torch.Tensor:
        """in-place broadcast from `src` (rank within the group): initial parameters and optimiser state"""
        if self.world_size > 1:
            dist.broadcast(t, src=dist.get_global_rank(self.group, src) if self.group is not None else src,
                           group=self.group)
        return t

    def shard_batch(self, global_batch: int) -> slice:
        """which samples of a global batch this rank owns (pure data parallelism)"""
        if global_batch % self.world_size != 0:
            raise ValueError(f"global batch {global_batch} is not divisible by world size {self.world_size}")
        per = global_batch // self.world_size
        return slice(self.rank * per, (self.rank + 1) * per)
