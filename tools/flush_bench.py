#!/usr/bin/env python
"""Memtable-flush throughput on one GPU (BASELINE.json configs[4] shape: Zipf(0.99) writes, 512-byte docs,
memtables of 8192 distinct keys): device-resident dbeel_flush_device per memtable vs the CPU oracle's
red-black-tree replay + flush_memtable_to_disk.  Usage: tools/flush_bench.py [n_writes]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import oracle  # noqa: E402
from dbeel_b200 import capi, sstable, storage_engine as se  # noqa: E402
from dbeel_b200 import workloads as W  # noqa: E402


def main():
    n_writes = int(sys.argv[1]) if len(sys.argv) > 1 else 400_000
    batch = W.make_arrival_batch(n_writes=n_writes, n_ids=n_writes // 4, doc_bytes=512, seed=5)
    nbytes = batch[0].size + batch[1].size
    # memtable boundaries (host logic, dbeel_memtable_cut)
    cuts, pos = [], 0
    while pos < n_writes:
        n = se.memtable_cut(batch, pos, 8192)
        cuts.append((pos, n))
        pos += n
    idx = batch[1].view("<u8").reshape(-1, 2)[:, 0]
    dev = torch.device("cuda:0")
    eng = capi.Engine(0)
    subs = []
    for p0, n in cuts:  # each memtable's arrivals as its own batch (offsets restart at 0)
        lo = int(idx[p0])
        hi = int(idx[p0 + n]) if p0 + n < n_writes else batch[0].size
        d = batch[0][lo:hi]
        ix = batch[1][16 * p0:16 * (p0 + n)].copy()
        ix.view("<u8").reshape(-1, 2)[:, 0] -= lo
        subs.append((torch.from_numpy(d.copy()).to(dev), torch.from_numpy(ix).to(dev)))
    cap_d = max(d.numel() for d, _ in subs) + 16
    cap_i = max(i.numel() for _, i in subs) + 16
    od = torch.empty(cap_d, dtype=torch.uint8, device=dev)
    oi = torch.empty(cap_i, dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()

    def run_all():
        tot_ms, items = 0.0, 0
        for d, i in subs:
            _, _, n = eng.flush_device((d.data_ptr(), d.numel(), i.data_ptr(), i.numel()), (od.data_ptr(), d.numel(), oi.data_ptr(), i.numel()))
            tot_ms += eng.stats()["ms_total"]
            items += n
        return tot_ms, items

    run_all()
    t0 = time.perf_counter()
    dev_ms, items = run_all()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0

    # all memtables in ONE launch sequence (dbeel_flush_many_device)
    tot_d = sum(d.numel() for d, _ in subs)
    tot_i = sum(i.numel() for _, i in subs)
    bd = torch.empty(tot_d + 16, dtype=torch.uint8, device=dev)
    bi = torch.empty(tot_i + 16, dtype=torch.uint8, device=dev)
    args = [(d.data_ptr(), d.numel(), i.data_ptr(), i.numel()) for d, i in subs]
    eng.flush_many_device(args, (bd.data_ptr(), tot_d, bi.data_ptr(), tot_i))
    t1 = time.perf_counter()
    _, _, many_items, _ = eng.flush_many_device(args, (bd.data_ptr(), tot_d, bi.data_ptr(), tot_i))
    torch.cuda.synchronize()
    many_wall = time.perf_counter() - t1
    many_ms = eng.stats()["ms_total"]
    assert many_items == items

    # several engines on the same GPU, one per host thread (= several shards sharing a GPU): small jobs overlap
    import threading
    def many_engines(T):
        engs = [capi.Engine(0) for _ in range(T)]
        outs = [(torch.empty(cap_d, dtype=torch.uint8, device=dev), torch.empty(cap_i, dtype=torch.uint8, device=dev)) for _ in range(T)]
        torch.cuda.synchronize()
        def work(t):
            for k in range(t, len(subs), T):
                d, i = subs[k]
                engs[t].flush_device((d.data_ptr(), d.numel(), i.data_ptr(), i.numel()),
                                     (outs[t][0].data_ptr(), d.numel(), outs[t][1].data_ptr(), i.numel()))
        for rep in range(2):
            ths = [threading.Thread(target=work, args=(t,)) for t in range(T)]
            t1 = time.perf_counter()
            for th in ths: th.start()
            for th in ths: th.join()
            torch.cuda.synchronize()
            el = time.perf_counter() - t1
        for e in engs: e.close()
        return el

    t0 = time.perf_counter()
    flushed = oracle.memtable_flushes(batch, capacity=8192, emulate_page_cache=True)
    cpu = time.perf_counter() - t0
    assert sum(n for _, _, n in flushed) == items
    print(f"{len(cuts)} memtables, {nbytes / 1e6:.0f} MB of arrivals, {items} entries flushed")
    print(f"GPU device-resident, 1 engine: {dev_ms:.2f} ms in kernels ({nbytes / 1e6 / dev_ms * 1e3:.0f} MB/s), {wall * 1e3:.1f} ms wall "
          f"({nbytes / 1e6 / wall:.0f} MB/s incl. launches + control-block read-back per memtable)")
    print(f"GPU device-resident, dbeel_flush_many_device (all {len(cuts)} memtables, one launch sequence): {many_ms:.2f} ms in kernels "
          f"({nbytes / 1e6 / many_ms * 1e3:.0f} MB/s), {many_wall * 1e3:.1f} ms wall ({nbytes / 1e6 / many_wall:.0f} MB/s)")
    for T in (2, 4, 8):
        el = many_engines(T)
        print(f"GPU device-resident, {T} engines / host threads on one GPU: {el * 1e3:.1f} ms wall ({nbytes / 1e6 / el:.0f} MB/s)")
    print(f"CPU oracle (rb-tree inserts + flush, 1 core): {cpu * 1e3:.0f} ms ({nbytes / 1e6 / cpu:.0f} MB/s)")

    # N4: recovery of ONE unflushed memtable from its write-ahead log (every write occupies >= one 4096-byte page)
    p0, n = cuts[0]
    ents = sstable.parse_run(batch[0], batch[1][:16 * (p0 + n)])[p0:p0 + n]
    wal = sstable.build_wal(ents)
    d_wal = torch.from_numpy(wal).to(dev)
    wd = torch.empty(subs[0][0].numel() + 64, dtype=torch.uint8, device=dev)
    wi = torch.empty(subs[0][1].numel() + 64, dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()
    for _ in range(2):
        eng.wal_flush_device(d_wal.data_ptr(), wal.size, (wd.data_ptr(), wd.numel() - 32, wi.data_ptr(), wi.numel() - 32))
    t2 = time.perf_counter()
    reps = 5
    for _ in range(reps):
        _, _, wn = eng.wal_flush_device(d_wal.data_ptr(), wal.size, (wd.data_ptr(), wd.numel() - 32, wi.data_ptr(), wi.numel() - 32))
    torch.cuda.synchronize()
    wal_wall = (time.perf_counter() - t2) / reps
    t3 = time.perf_counter()
    _, _, on, _ = oracle.wal_flush(wal, 8192, emulate_page_cache=True)
    wal_cpu = time.perf_counter() - t3
    assert wn == on == 8192 or wn == on
    print(f"WAL replay + flush of one memtable ({n} logged writes, {wal.size / 1e6:.0f} MB log, {wn} distinct keys): GPU device-resident "
          f"{wal_wall * 1e3:.2f} ms wall ({wal.size / 1e6 / wal_wall:.0f} MB/s of log, {n / wal_wall / 1e6:.1f} M writes/s); "
          f"CPU oracle {wal_cpu * 1e3:.1f} ms ({n / wal_cpu / 1e6:.2f} M writes/s)")


if __name__ == "__main__":
    main()
