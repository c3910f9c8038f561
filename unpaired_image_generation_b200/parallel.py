"""Data-parallel gradient exchange for the CycleGAN step.

Samples are independent (InstanceNorm is per-sample), so the only collective on the path is a
sum-allreduce of the two flat gradient buffers (generators: 22.8 M floats, discriminators: 5.5 M),
followed by a 1/world scale folded into the Adam kernel.  This module is device-agnostic on
purpose: the same code runs over NCCL on B200s and over gloo in the CPU test-suite.
Stand-in counterpart: a single-process step on the concatenated batch
(oracle/cyclegan_standin.py:374 train_step) -- see tests/test_parallel_gloo.py.
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.distributed as dist


class GradSync:
    def __init__(self, process_group: Optional["dist.ProcessGroup"] = None, bucket_elems: int = 1 << 25):
        self.group = process_group
        self.enabled = dist.is_available() and dist.is_initialized()
        self.world_size = dist.get_world_size(process_group) if self.enabled else 1
        self.rank = dist.get_rank(process_group) if self.enabled else 0
        self.bucket_elems = int(bucket_elems)

    @property
    def grad_scale(self) -> float:
        """factor the optimiser applies to the summed gradients (mean over ranks)"""
        return 1.0 / self.world_size

    def buckets(self, flat: torch.Tensor) -> List[torch.Tensor]:
        """contiguous views covering `flat`: few, large buckets (NVSwitch: size for launch latency)"""
        n = flat.numel()
        if n <= self.bucket_elems:
            return [flat]
        return [flat[i:min(n, i + self.bucket_elems)] for i in range(0, n, self.bucket_elems)]

    def all_reduce_(self, flat: torch.Tensor) -> torch.Tensor:
        """in-place SUM over ranks of a flat gradient buffer (enqueued on the current stream)"""
        if self.world_size == 1:
            return flat
        for b in self.buckets(flat):
            dist.all_reduce(b, op=dist.ReduceOp.SUM, group=self.group)
        return flat

    def shard_batch(self, global_batch: int) -> slice:
        """which samples of a global batch this rank owns (pure data parallelism)"""
        if global_batch % self.world_size != 0:
            raise ValueError(f"global batch {global_batch} is not divisible by world size {self.world_size}")
        per = global_batch // self.world_size
        return slice(self.rank * per, (self.rank + 1) * per)
