"""The N > 1 path on CPU: world_size-2 gloo processes exercise the job hand-off, the barrier and
the max/sum reduction that bench.py uses over NCCL on the GPUs (no data-path collective exists)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dbeel_b200 import shard_jobs as sj


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        dev = torch.device("cpu")
        jobs = [sj.ShardJob(i, 40 + i, 8, 1_000_000, 256, False) for i in range(8)] if rank == 0 else None
        mine = sj.hand_off(jobs, dev)
        sj.barrier(dev)
        rep = sj.reduce_report({"ms": 10.0 + rank, "bytes": 1000.0 * len(mine), "launches": 15.0 * len(mine)}, dev,
                               max_keys=["ms"], sum_keys=["bytes", "launches"])
        out.put((rank, [j.shard_id for j in mine], [j.seed for j in mine], rep))
    finally:
        dist.destroy_process_group()


def test_job_hand_off_world_2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, shards0, seeds0, rep0), (r1, shards1, seeds1, rep1) = got
    assert shards0 == [0, 2, 4, 6] and shards1 == [1, 3, 5, 7]  # shard i -> rank i mod N, every job exactly once
    assert seeds0 == [40, 42, 44, 46] and seeds1 == [41, 43, 45, 47]
    assert rep0 == rep1 == {"ms": 11.0, "bytes": 8000.0, "launches": 120.0}  # max over ranks / sum over ranks


def test_single_process_is_a_no_op():
    jobs = [sj.ShardJob(i, i, 2, 10, 64, True) for i in range(3)]
    assert sj.hand_off(jobs, torch.device("cpu")) == jobs
    assert sj.world() == 1 and sj.rank() == 0 and sj.owner_of(5, 4) == 1
    assert sj.reduce_report({"ms": 1.0}, torch.device("cpu"), ["ms"], []) == {"ms": 1.0}
    assert sj.ShardJob.from_row(jobs[1].as_row()) == jobs[1]


def test_reference_arm_fits_its_steps_into_the_budget():
    """bench.py --impl reference: the number of concurrent full-size jobs per step is cut so that K + W steps end within the
    budget, never below the 8 jobs of one GPU step, never above the host's threads."""
    import bench
    # 8 concurrent jobs took 5 s: 0.625 s per job; 20 + 5 steps in 210 s -> 8.4 s per step -> 13 jobs
    assert bench.ref_jobs_per_step(64, 8, 5.0, 20, 5, 210.0) == 13
    # a short run has room for every thread
    assert bench.ref_jobs_per_step(64, 8, 5.0, 3, 1, 210.0) == 64
    # a slow host never drops below one GPU step's jobs
    assert bench.ref_jobs_per_step(64, 8, 60.0, 20, 5, 210.0) == 8
    assert bench.ref_jobs_per_step(4, 4, 60.0, 20, 5, 210.0) == 4
