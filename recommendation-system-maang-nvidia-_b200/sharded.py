"""Catalog partitioned over the GPUs of one node (BASELINE config 3).

One process per GPU.  Rank g holds the contiguous catalog rows [bounds[g], bounds[g+1]) in its own
IndexFlatIP; queries are replicated.  A search is: local scan + local top-k (global row numbers via
id_base) -> one all-gather of the [nq, k] (score, row) lists over NCCL/NVLink (G*nq*k*12 bytes, ~10 MB at
G=8, nq=1024, k=100) -> on-device G-way merge (K7).  The tie-break (score desc, row asc) makes the merge
independent of the shard order, so every rank ends with the same lists as a single-GPU search over the
whole catalog -- bit for bit.

torch.distributed is the plumbing (process group, all_gather_into_tensor); the scan and the merge are this
repo's kernels.  `local_search` / `merge` are injectable so the host logic (bounds, id offsets, gather
layout) can be exercised on CPU with gloo, where the tests plug the oracle in as the checker.
"""

from typing import Callable, List, Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist


def shard_bounds(n_total: int, world: int) -> List[int]:
    """Contiguous row ranges: shard g = [g*N//G, (g+1)*N//G)."""
    return [n_total * g // world for g in range(world + 1)]


class ShardedIndexFlatIP:
    def __init__(self, d: int, n_total: int, group: Optional[dist.ProcessGroup] = None,
                 local_index=None, local_search: Optional[Callable] = None, merge: Optional[Callable] = None):
        self.d = int(d)
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.bounds = shard_bounds(int(n_total), self.world)
        self.row_begin, self.row_end = self.bounds[self.rank], self.bounds[self.rank + 1]
        self.n_total = int(n_total)
        if local_search is None:
            from .index import IndexFlatIP
            from . import ops
            self.local = local_index if local_index is not None else IndexFlatIP(d)
            self._local_search = lambda q, k: self.local.search(q, k, id_base=self.row_begin)
            self._merge = ops.topk_merge
        else:
            self.local = local_index
            self._local_search = local_search
            self._merge = merge

    @property
    def ntotal(self) -> int:
        return self.n_total

    def add_local(self, x) -> None:
        """Append rows of this rank's range (callers feed rows row_begin.. in order)."""
        self.local.add(x)
        if self.local.ntotal > self.row_end - self.row_begin:
            raise ValueError("more rows added than this rank's shard holds")

    def search(self, q, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
        """q: [nq, d] replicated on every rank -> (D [nq, k], I [nq, k] global rows), identical on every rank."""
        Dl, Il = self._local_search(q, k)
        if self.world == 1:
            return Dl, Il
        nq, kk = int(Dl.shape[0]), int(Dl.shape[1])
        # rank-major concatenation along dim 0 (the layout both NCCL and gloo accept), viewed as [G, nq, k]
        Dg = torch.empty((self.world * nq, kk), dtype=Dl.dtype, device=Dl.device)
        Ig = torch.empty((self.world * nq, kk), dtype=Il.dtype, device=Il.device)
        dist.all_gather_into_tensor(Dg, Dl.contiguous(), group=self.group)
        dist.all_gather_into_tensor(Ig, Il.contiguous(), group=self.group)
        return self._merge(Dg.view(self.world, nq, kk), Ig.view(self.world, nq, kk))


def gather_bytes(world: int, nq: int, k: int) -> int:
    """Bytes each rank receives in the exchange step (f32 score + i64 row per hit)."""
    return world * nq * k * (4 + 8)
