#!/usr/bin/env python
"""bench.py -- compaction throughput (MB/s of input bytes) of the B200 engine vs the CPU path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A *step* is one pass of the hot path over one batch of synthetic input: one 8-way compaction
of 1M-key x 256-byte-doc runs (BASELINE.json configs[1], "cfg2": 2.57 GB of .data+.index in,
bloom on, tombstones dropped).  With N > 1 every rank compacts its own independent shard of
that shape (configs[3], weak scaling, no data-path collective; NCCL only hands the job table
out and closes the barriers).

`value`    whole-job MB/s of input bytes with the runs resident in HBM (all kernels of the
           pipeline; timed with CUDA events on the engine's stream, max over ranks)
`e2e`      the same metric through dbeel_compact() with HOST (pinned) buffers: H2D of the runs
           and D2H of the output SSTable inside the timed region
`roofline` the dominant kernel (k_gather): algorithmic bytes / its CUDA-event time vs the
           measured HBM copy peak
`cpu_baseline` the CPU oracle (port of the reference's single-threaded compact()) on one core

--impl reference times the CPU path alone with every host core busy (one shard compaction per
core, the way dbeel's thread-per-core runtime would run them).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "compaction throughput MB/s (input bytes) at 1/2/4/8 B200 vs CPU shards"
UNIT = "MB/s"
SEED32 = bytes(range(32))
FALLBACK_HBM_GBS = 6650.0  # B200_PROFILING.md fallback when MEASURED_PEAKS.json is absent


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def gather_traffic_from_profile():
    """dram read+write bytes per k_gather launch from the committed ncu summary, if any."""
    p = os.path.join(ROOT, "profiles", "gather_traffic.json")
    try:
        with open(p) as f:
            return json.load(f).get("dram_bytes_per_launch")
    except Exception:
        return None


# ------------------------------------------------------------------------------------ clocks

class ClockSampler:
    """Samples SM clock / throttle reasons of one GPU with NVML while the timed region runs."""
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x10: "sync_boost"}

    def __init__(self, device_index: int, period_s: float = 0.01):
        self.samples = []
        self.marks = []
        self._stop = threading.Event()
        self._period = period_s
        self._ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nv = pynvml
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = device_index
            if vis:
                try:
                    phys = int(vis.split(",")[device_index])
                except Exception:
                    phys = device_index
            self._h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
            self._ok = True
        except Exception as ex:  # pragma: no cover
            log(f"[bench] NVML unavailable ({ex}); clocks not sampled")
            self.max_mhz = None
        self._t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        nv = self._nv
        while not self._stop.is_set():
            try:
                mhz = nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)
                try:
                    rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self._h)
                except Exception:
                    rs = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                self.samples.append((time.perf_counter(), mhz, rs))
            except Exception:
                pass
            time.sleep(self._period)

    def start(self):
        if self._ok:
            self._t.start()

    def stop(self):
        self._stop.set()
        if self._ok:
            self._t.join(timeout=2)

    def summary(self, t0: float, t1: float) -> dict:
        if not self._ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        inside = [s for s in self.samples if t0 <= s[0] <= t1]
        if not inside:  # region shorter than the sampling period: take the nearest samples
            inside = sorted(self.samples, key=lambda s: min(abs(s[0] - t0), abs(s[0] - t1)))[:3]
        mhz = sorted(s[1] for s in inside)
        bits = 0
        for s in inside:
            bits |= s[2]
        return {"sm_mhz": mhz[len(mhz) // 2], "sm_max_mhz": self.max_mhz,
                "reasons": sorted(n for b, n in self.REASONS.items() if bits & b), "samples": len(inside)}


# ------------------------------------------------------------------------------------ workload

def shard_config(rank: int, world: int, workload: str = "cfg2"):
    from dbeel_b200 import workloads as W
    if workload == "cfg3":  # BASELINE.json configs[2]: not the headline, kept for cross-checks
        return W.CFG3
    if world == 1:
        return W.CFG2
    return W.cfg4_shard(rank)


def make_runs_parallel(cfg):
    """make_merge_runs, one thread per run (numpy releases the GIL in the heavy parts)."""
    import dataclasses
    from concurrent.futures import ThreadPoolExecutor

    from dbeel_b200 import workloads as W

    def one(r):
        rng = np.random.default_rng([cfg.seed, r])
        ids = W._run_ids(cfg, rng, r)
        keys = W.format_keys(ids)
        dense = W._dense_rows([cfg.seed, r, 7], ids.shape[0], keys.shape[1], cfg.doc_bytes)
        tomb = rng.random(ids.shape[0]) < cfg.tombstone_frac
        ts = W.BASE_TS + (np.int64(r) * ids.shape[0] + rng.permutation(ids.shape[0]).astype(np.int64)) * 1000
        from dbeel_b200.sstable import build_run_dense
        return build_run_dense(dense, keys, ts, tomb)

    with ThreadPoolExecutor(max_workers=min(cfg.n_runs, os.cpu_count() or 1)) as ex:
        return list(ex.map(one, range(cfg.n_runs)))


# ------------------------------------------------------------------------------------ GPU arm

def run_gpu(args):
    import torch
    import torch.distributed as dist

    from dbeel_b200 import capi, sstable

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            log(f"[bench] --gpus {args.gpus} needs torchrun with {args.gpus} ranks; running 1 GPU")
        args.gpus = world
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the compaction engine has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    from dbeel_b200 import shard_jobs as sj
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    # job hand-off (NCCL broadcast when N > 1): rank 0 owns the table, one independent shard compaction per GPU
    table = [sj.ShardJob(r, 40 + r if world > 1 else 2, 8, 1_000_000, 256, False) for r in range(world)] if rank == 0 else None
    mine = sj.hand_off(table, dev)
    assert len(mine) == 1 and mine[0].shard_id == rank, mine

    cfg = shard_config(rank, world, args.workload)
    t = time.time()
    runs = make_runs_parallel(cfg)
    in_bytes = sstable.input_bytes(runs)
    log(f"[bench r{rank}] generated {cfg.name}: {in_bytes / 1e6:.1f} MB in {time.time() - t:.1f}s")

    eng = capi.Engine(local)
    opts = capi.make_opts(cfg.keep_tombstones, seed=SEED32)
    dc, ic, bc = capi.compact_bound([(d.size, i.size) for d, i in runs], opts)

    # device-resident inputs / outputs (torch only owns the memory)
    t_runs = [(torch.from_numpy(d).to(dev), torch.from_numpy(i).to(dev)) for d, i in runs]
    od = torch.empty(dc + 16, dtype=torch.uint8, device=dev)
    oi = torch.empty(ic + 16, dtype=torch.uint8, device=dev)
    ob = torch.empty(bc + 16, dtype=torch.uint8, device=dev)
    d_runs = [(d.data_ptr(), d.numel(), i.data_ptr(), i.numel()) for d, i in t_runs]
    d_out = (od.data_ptr(), dc, oi.data_ptr(), ic, ob.data_ptr(), bc)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        return eng.compact_device(d_runs, d_out, opts)

    for _ in range(max(3, args.warmup)):
        res = step()
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    t0 = time.perf_counter()
    dev_ms = 0.0
    gather_ms = 0.0
    launches = 0
    stage = {"ms_extract": 0.0, "ms_merge": 0.0, "ms_resolve": 0.0, "ms_gather": 0.0}
    for _ in range(args.steps):
        res = step()
        st = eng.stats()
        dev_ms += st["ms_total"]
        gather_ms += st["ms_gather"]
        launches += st["kernel_launches"]
        for k in stage:
            stage[k] += st[k]
    barrier()
    t1 = time.perf_counter()
    sampler.stop()
    wall_ms = (t1 - t0) * 1e3
    clocks = sampler.summary(t0, t1)
    st = eng.stats()
    dl, il, bl, items = res

    # max over ranks of the device time (CUDA events on the engine's stream, summed over K steps)
    tm = torch.tensor([dev_ms, wall_ms, gather_ms], dtype=torch.float64, device=dev)
    tot = torch.tensor([float(in_bytes), float(launches)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    dev_ms_max, wall_ms_max, gather_ms_max = (float(x) for x in tm.tolist())
    total_in = float(tot[0])
    value = total_in * args.steps / 1e6 / (dev_ms_max / 1e3)

    # ---- end to end through the host entry point (pinned host buffers, H2D + D2H timed)
    e2e_steps = max(2, min(5, args.steps))
    pins = []
    h_runs = []
    for d, i in runs:
        pd, pi = capi.PinnedBuffer(d.size), capi.PinnedBuffer(i.size)
        pd.array[:] = d
        pi.array[:] = i
        pins += [pd, pi]
        h_runs.append((pd.array, pi.array))
    pod, poi, pob = capi.PinnedBuffer(max(1, dc)), capi.PinnedBuffer(max(1, ic)), capi.PinnedBuffer(max(1, bc))
    del t_runs, od, oi, ob
    torch.cuda.empty_cache()
    hres = eng.compact(h_runs, cfg.keep_tombstones, seed=SEED32, out_buffers=(pod.array, poi.array, pob.array))
    barrier()
    e0 = time.perf_counter()
    for _ in range(e2e_steps):
        hres = eng.compact(h_runs, cfg.keep_tombstones, seed=SEED32, out_buffers=(pod.array, poi.array, pob.array))
    barrier()
    e1 = time.perf_counter()
    st_e2e = eng.stats()
    e2e_ms = torch.tensor([(e1 - e0) * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_value = total_in * e2e_steps / 1e6 / (float(e2e_ms[0]) / 1e3)
    out_bytes = int(hres[0].size + hres[1].size + (hres[2].size if hres[2] is not None else 0))

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- CPU baseline + byte parity of this very run (rank 0, N = 1 only)
    cpu = None
    parity = None
    if world == 1 and not args.no_cpu:
        import oracle
        t = time.perf_counter()
        cd, ci, cb, cn = oracle.compact(runs, cfg.keep_tombstones, seed=SEED32, emulate_page_cache=True)
        cpu_s = time.perf_counter() - t
        cpu = {"value": round(in_bytes / 1e6 / cpu_s, 2), "unit": UNIT, "cores": 1, "kind": "port",
               "sample": f"the full {cfg.name} workload, 1 pass ({in_bytes / 1e6:.0f} MB in, {cpu_s:.1f} s), "
                         "C oracle port of LSMTree::compact incl. page-cache write-through copies, RAM-resident files"}
        parity = bool(cn == hres[3] and np.array_equal(cd, hres[0]) and np.array_equal(ci, hres[1])
                      and cb is not None and hres[2] is not None and np.array_equal(cb, hres[2]))
        log(f"[bench] cpu oracle {cpu['value']} MB/s; byte parity of the benchmarked output: {parity}")

    peak, peak_src = hbm_peak()
    gbytes = st["gather_bytes"]
    g_ms = gather_ms_max / args.steps
    achieved = gbytes / 1e9 / (g_ms / 1e3)
    algo_total = st["input_bytes"] + st["output_bytes"]
    line = {
        "metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(3, args.warmup), "ms_per_step": round(dev_ms_max / args.steps, 4), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": cfg.name if world == 1 else "cfg4: one cfg2-shaped shard compaction per GPU (seeds 40+rank)",
                   "runs": cfg.n_runs, "keys_per_run": cfg.keys_per_run, "doc_bytes": cfg.doc_bytes,
                   "input_bytes_per_step_per_gpu": in_bytes, "output_bytes_per_step": out_bytes,
                   "entries_out": items, "bloom": bool(bl), "keep_tombstones": cfg.keep_tombstones,
                   "l2_policy": "inputs_larger_than_l2 (2.57 GB per step vs 126 MB L2)",
                   "parallelism": f"{world} independent shard(s), one per GPU, no data-path collective",
                   "timing": "sum over steps of CUDA-event time on the engine stream (first to last kernel), max over ranks"},
        "wall_ms_per_step": round(wall_ms_max / args.steps, 4),
        "stage_ms": {k: round(v / args.steps, 4) for k, v in stage.items()},
        "pipeline_roofline": {"algo_bytes": algo_total, "achieved_gbs": round(algo_total / 1e9 / (dev_ms / args.steps / 1e3), 1),
                              "frac": round(algo_total / 1e9 / (dev_ms / args.steps / 1e3) / peak, 4),
                              "read_only_frac": round(st["input_bytes"] / 1e9 / (dev_ms / args.steps / 1e3) / peak, 4)},
        "roofline": {"kernel": "k_gather", "bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                     "frac": round(achieved / peak, 4), "traffic": gather_traffic_from_profile(),
                     "algo_bytes_per_launch": gbytes, "ms_per_launch": round(g_ms, 4), "peak_source": peak_src},
        "cpu_baseline": cpu,
        "e2e": {"value": round(e2e_value, 1), "unit": UNIT, "h2d_bytes_per_step": in_bytes,
                "d2h_bytes_per_step": out_bytes, "steps": e2e_steps,
                "ms_per_step": round(float(e2e_ms[0]) / e2e_steps, 3),
                "ms_kernels": round(st_e2e["ms_total"], 3), "partitions": st_e2e["partitions"],
                "api": "dbeel_compact (host pinned buffers; key-range partitions pipelined over H2D / kernels / D2H streams)"},
        "gpu_launches": int(tot[1]),
        "clocks": clocks,
        "parity_vs_oracle": parity,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------ reference arm

def run_reference(args):
    """The reference's CPU implementation of the path on this box's host cores: the C oracle
    port of LSMTree::compact (the Rust reference cannot be built here: no cargo/rustc), one
    single-threaded shard compaction per core, all cores busy -- dbeel's thread-per-core model."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle
    from dbeel_b200 import sstable
    from dbeel_b200 import workloads as W

    cores = os.cpu_count() or 1
    try:
        cores = len(os.sched_getaffinity(0))
    except Exception:
        pass
    threads = max(1, min(cores, 64))
    steps, warm = args.steps, args.warmup
    # calibrate the per-thread sample so the whole run stays within ~2.5 minutes
    probe = W.make_merge_runs(W.scaled(W.CFG2, 20_000))
    t = time.perf_counter()
    oracle.compact(probe, False, seed=SEED32, emulate_page_cache=True)
    mbps_1 = sstable.input_bytes(probe) / 1e6 / (time.perf_counter() - t)
    budget_s = min(2.0, 120.0 / max(1, steps + warm))
    # bounded sample: at most 60k keys per run (154 MB per shard) so that `threads` concurrent compactions and
    # their outputs stay within a few tens of GB of RAM whatever the core count
    keys = int(max(10_000, min(60_000, budget_s * mbps_1 * 1e6 / (8 * 321.0) * 0.6)))
    distinct = [W.make_merge_runs(W.scaled(W.cfg4_shard(i), keys)) for i in range(min(threads, 8))]
    shards = [distinct[i % len(distinct)] for i in range(threads)]  # inputs are read-only: threads may share them
    in_bytes = sum(sstable.input_bytes(s) for s in shards)
    log(f"[bench ref] {threads} threads x 8-way x {keys} keys ({in_bytes / 1e6:.0f} MB per step), "
        f"1-thread probe {mbps_1:.0f} MB/s")

    def one(s):
        oracle.compact(s, False, seed=SEED32, emulate_page_cache=True)

    def step():
        ts = [threading.Thread(target=one, args=(s,)) for s in shards]
        for th in ts:
            th.start()
        for th in ts:
            th.join()

    for _ in range(warm):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    value = in_bytes * steps / 1e6 / dt
    sample = (f"{threads} concurrent single-threaded shard compactions per step, each 8 runs x {keys} keys x 256 B docs "
              f"(cfg2 shape scaled; {in_bytes / 1e6:.0f} MB in per step), RAM-resident files, page-cache copies emulated")
    line = {"impl": "reference", "metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": args.gpus,
            "steps": steps, "warmup": warm, "ms_per_step": round(dt / steps * 1e3, 2), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": "cfg2-8way-1M-256B (bounded sample per core)", "runs": 8, "keys_per_run": keys,
                       "doc_bytes": 256, "threads": threads},
            "cpu_baseline": {"value": round(value, 2), "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": round(value, 2), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="dbeel_b200", choices=["dbeel_b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
    ap.add_argument("--workload", default="cfg2", choices=["cfg2", "cfg3"], help="cfg2 is BASELINE.json's headline config")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
