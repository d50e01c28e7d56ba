"""Multi-GPU layout of the batched step: envs are independent, so the batch is sharded by contiguous
env ranges, one process per GPU (`torch.distributed`, backend nccl; gloo in the CPU test-suite), with
no collective inside the physics.  The single exchange of the path is the end-of-step observation
all-gather."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Tuple

import torch
import torch.distributed as dist


@dataclass(frozen=True)
class Shard:
    rank: int
    world: int
    n_total: int

    @property
    def bounds(self) -> Tuple[int, int]:
        """Contiguous env range [lo, hi) of this rank; sizes differ by at most one."""
        base, rem = divmod(self.n_total, self.world)
        lo = self.rank * base + min(self.rank, rem)
        return lo, lo + base + (1 if self.rank < rem else 0)

    @property
    def n_local(self) -> int:
        lo, hi = self.bounds
        return hi - lo

    def slice_actions(self, actions: torch.Tensor) -> torch.Tensor:
        """Every rank holds the replicated action tensor `[n_total, nmotors]`; it reads its own rows."""
        lo, hi = self.bounds
        return actions[lo:hi]


def gather_observations(local_obs: torch.Tensor, shard: Shard) -> torch.Tensor:
    """All-gather of `[n_local, width]` observation shards into `[n_total, width]`, global env order.
    Equal shards use one `all_gather_into_tensor` (a single NCCL all-gather over NVLink); ragged ones
    fall back to padded gathers."""
    if shard.world == 1:
        return local_obs
    width = local_obs.shape[1]
    base, rem = divmod(shard.n_total, shard.world)
    if rem == 0:
        out = torch.empty((shard.n_total, width), dtype=local_obs.dtype, device=local_obs.device)
        dist.all_gather_into_tensor(out, local_obs.contiguous())
        return out
    pad = torch.zeros((base + 1, width), dtype=local_obs.dtype, device=local_obs.device)
    pad[:local_obs.shape[0]] = local_obs
    parts = [torch.empty_like(pad) for _ in range(shard.world)]
    dist.all_gather(parts, pad)
    sizes = [Shard(r, shard.world, shard.n_total).n_local for r in range(shard.world)]
    return torch.cat([p[:n] for p, n in zip(parts, sizes)], dim=0)
