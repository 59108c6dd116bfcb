#!/usr/bin/env python
"""Batched point lookups (row N2) on the table a cfg2 compaction leaves in HBM: Mlookups/s of dbeel_get_many_device
(reference-loop and exact modes, present and absent keys) next to the CPU oracle's restatement of get_entry's loop.
Usage: tools/lookup_bench.py [n_queries (default 4000000)] [keys_per_run (default 1000000)]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import oracle  # noqa: E402  (CPU baseline only)
from bench import SEED32, make_runs_parallel  # noqa: E402
from dbeel_b200 import capi  # noqa: E402
from dbeel_b200 import workloads as W  # noqa: E402


def main():
    nq = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
    kpr = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
    cfg = W.CFG2 if kpr == 1_000_000 else W.scaled(W.CFG2, kpr)
    runs = make_runs_parallel(cfg)
    dev = torch.device("cuda:0")
    t_runs = [(torch.from_numpy(d).to(dev), torch.from_numpy(i).to(dev)) for d, i in runs]
    opts = capi.make_opts(cfg.keep_tombstones, seed=SEED32)
    dc, ic, bc = capi.compact_bound([(d.size, i.size) for d, i in runs], opts)
    od = torch.empty(dc + 16, dtype=torch.uint8, device=dev)
    oi = torch.empty(ic + 16, dtype=torch.uint8, device=dev)
    ob = torch.empty(bc + 16, dtype=torch.uint8, device=dev)
    eng = capi.Engine(0)
    dl, il, bl, n = eng.compact_device([(d.data_ptr(), d.numel(), i.data_ptr(), i.numel()) for d, i in t_runs],
                                       (od.data_ptr(), dc, oi.data_ptr(), ic, ob.data_ptr(), bc), opts)
    del t_runs
    print(f"table: {n} entries, {dl / 1e6:.0f} MB .data, {il / 1e6:.0f} MB .index, {bl / 1e6:.1f} MB .bloom", flush=True)
    rng = np.random.default_rng(11)
    # present keys: read them out of the table itself (index record -> entry key); absent keys: ids above the id space
    h_index = oi[:il].cpu().numpy().view("<u8").reshape(-1, 2)
    pick = rng.integers(0, n, nq // 2)
    offs = h_index[pick, 0].astype(np.int64)
    h_data = od[:dl].cpu().numpy()
    klen = 17
    present = np.stack([h_data[offs + 8 + b] for b in range(klen)], axis=1)  # every key of this workload is 17 bytes
    absent_ids = rng.integers(20_000_000, 1 << 40, nq - nq // 2)
    absent = np.frombuffer(b"".join(b"\xb0k%015d" % int(x) for x in absent_ids), dtype=np.uint8).reshape(-1, klen)
    for name, keys in (("present keys", present), ("absent keys", absent), ("50/50 mix", np.concatenate([present, absent])[rng.permutation(nq)])):
        m = keys.shape[0]
        blob = np.ascontiguousarray(keys).reshape(-1)
        off = (np.arange(m + 1, dtype=np.uint64) * klen)
        d_keys, d_off = torch.from_numpy(blob.copy()).to(dev), torch.from_numpy(off.view(np.int64)).to(dev)
        d_res = torch.zeros(2 * m, dtype=torch.int64, device=dev)
        table = [(od.data_ptr(), dl, oi.data_ptr(), il, ob.data_ptr(), bl)]
        for mode, mname in ((capi.LOOKUP_REFERENCE, "reference loop"), (capi.LOOKUP_EXACT, "exact")):
            for _ in range(2):
                eng.get_many_device(table, d_keys.data_ptr(), d_off.data_ptr(), m, d_res.data_ptr(), mode)
            ms = []
            for _ in range(5):
                eng.get_many_device(table, d_keys.data_ptr(), d_off.data_ptr(), m, d_res.data_ptr(), mode)
                ms.append(eng.stats()["ms_total"])
            res = d_res.cpu().numpy().view(capi.LOOKUP_DTYPE)
            found = int((res["table"] >= 0).sum())
            t = float(np.median(ms))
            print(f"GPU {name:13s} {mname:15s}: {m} keys in {t:7.3f} ms = {m / t / 1e3:8.1f} M lookups/s  (found {found}, "
                  f"filter said no for {int((res['bloom_rejects'] & 0x7fffffff).sum())})", flush=True)
        if name == "50/50 mix":
            sample = min(m, 400_000)
            host_table = [(h_data, oi[:il].cpu().numpy(), ob[:bl].cpu().numpy())]
            t0 = time.perf_counter()
            et, er, ej = oracle.get_many(host_table, blob[: sample * klen], off[: sample + 1])
            dt = time.perf_counter() - t0
            eng.get_many_device(table, d_keys.data_ptr(), d_off.data_ptr(), m, d_res.data_ptr(), capi.LOOKUP_REFERENCE)
            res = d_res.cpu().numpy().view(capi.LOOKUP_DTYPE)[:sample]
            same = np.array_equal(res["table"], et) and np.array_equal(res["bloom_rejects"], ej) and \
                np.array_equal(np.where(res["table"] >= 0, res["record"], 0), er)
            print(f"CPU oracle (get_entry's loop restated, 1 core): {sample} keys in {dt * 1e3:.0f} ms = {sample / dt / 1e6:.2f} M lookups/s; "
                  f"GPU rows identical to the oracle's: {same}", flush=True)
    eng.close()


if __name__ == "__main__":
    main()
