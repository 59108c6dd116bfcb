"""Job hand-off for multi-GPU runs: compactions of different shards are independent
(src/main.rs:39-64 one executor per core, src/shards.rs:313-317 one directory per shard), so the
data path has no collective at all.  torch.distributed (NCCL on GPUs, gloo in the CPU tests) only
(1) broadcasts the job table from rank 0, (2) closes barriers around timed regions and (3) reduces
per-rank timings / counters for the report.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence

import torch
import torch.distributed as dist


@dataclass(frozen=True)
class ShardJob:
    shard_id: int  # the dbeel shard (core) whose tree is being compacted
    seed: int  # synthetic-input seed (workloads.cfg4_shard)
    n_runs: int
    keys_per_run: int
    doc_bytes: int
    keep_tombstones: bool

    def as_row(self) -> List[int]:
        return [self.shard_id, self.seed, self.n_runs, self.keys_per_run, self.doc_bytes, int(self.keep_tombstones)]

    @staticmethod
    def from_row(row: Sequence[int]) -> "ShardJob":
        return ShardJob(int(row[0]), int(row[1]), int(row[2]), int(row[3]), int(row[4]), bool(row[5]))


def world() -> int:
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def rank() -> int:
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def owner_of(shard_id: int, world_size: int) -> int:
    """Shard i runs on GPU i mod N: 8 shards map one per GPU at N = 8, 8/N per GPU below."""
    return shard_id % world_size


def hand_off(jobs: Optional[Sequence[ShardJob]], device: torch.device, src: int = 0) -> List[ShardJob]:
    """Rank `src` owns the job table; every rank returns the jobs it must run, in table order."""
    w, r = world(), rank()
    if w == 1:
        return list(jobs or [])
    n = torch.tensor([len(jobs) if r == src and jobs is not None else 0], dtype=torch.int64, device=device)
    dist.broadcast(n, src=src)
    table = torch.zeros((int(n.item()), 6), dtype=torch.int64, device=device)
    if r == src:
        table.copy_(torch.tensor([j.as_row() for j in jobs], dtype=torch.int64).reshape(-1, 6))
    dist.broadcast(table, src=src)
    rows = table.cpu().tolist()
    return [ShardJob.from_row(row) for row in rows if owner_of(int(row[0]), w) == r]


def barrier(device: torch.device) -> None:
    if world() > 1:
        dist.barrier()
    if device.type == "cuda":
        torch.cuda.synchronize(device)


def reduce_report(local: Dict[str, float], device: torch.device, max_keys: Sequence[str], sum_keys: Sequence[str]) -> Dict[str, float]:
    """Timings combine as the MAX over ranks, byte / entry / launch counters as the SUM."""
    out = dict(local)
    if world() == 1:
        return out
    if max_keys:
        t = torch.tensor([float(local[k]) for k in max_keys], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out.update({k: float(v) for k, v in zip(max_keys, t.tolist())})
    if sum_keys:
        t = torch.tensor([float(local[k]) for k in sum_keys], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        out.update({k: float(v) for k, v in zip(sum_keys, t.tolist())})
    return out
