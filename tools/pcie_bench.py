#!/usr/bin/env python
"""What the host link can do, next to what the e2e path gets: H2D alone, D2H alone, both at once
(two streams, pinned host memory), for a few chunk sizes and CPU affinities (pinned pages land on the
NUMA node of the allocating thread).  Usage: tools/pcie_bench.py [GB_in] [GB_out]"""
import os
import sys
import time

import torch


def bench(gb_in: float, gb_out: float, chunk_mb: int, label: str):
    dev = torch.device("cuda:0")
    n_in, n_out = int(gb_in * 1e9), int(gb_out * 1e9)
    h_in = torch.empty(n_in, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(n_out, dtype=torch.uint8).pin_memory()
    h_in.fill_(1)
    h_out.fill_(2)
    d_in = torch.empty(n_in, dtype=torch.uint8, device=dev)
    d_out = torch.ones(n_out, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    ch = chunk_mb << 20

    def h2d():
        with torch.cuda.stream(s1):
            for o in range(0, n_in, ch):
                d_in[o:o + ch].copy_(h_in[o:o + ch], non_blocking=True)

    def d2h():
        with torch.cuda.stream(s2):
            for o in range(0, n_out, ch):
                h_out[o:o + ch].copy_(d_out[o:o + ch], non_blocking=True)

    def timed(fns):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for f in fns:
            f()
        torch.cuda.synchronize()
        return time.perf_counter() - t0

    for f in (h2d, d2h):
        timed([f])
    t_in = min(timed([h2d]) for _ in range(3))
    t_out = min(timed([d2h]) for _ in range(3))
    t_both = min(timed([h2d, d2h]) for _ in range(3))
    print(f"{label:28s} chunk {chunk_mb:4d} MB: H2D {gb_in / t_in:6.1f} GB/s ({t_in * 1e3:5.1f} ms)  D2H {gb_out / t_out:6.1f} GB/s "
          f"({t_out * 1e3:5.1f} ms)  both at once {(gb_in + gb_out) / t_both:6.1f} GB/s aggregate ({t_both * 1e3:5.1f} ms)", flush=True)


def main():
    gb_in = float(sys.argv[1]) if len(sys.argv) > 1 else 2.547
    gb_out = float(sys.argv[2]) if len(sys.argv) > 2 else 2.059
    ncpu = os.cpu_count()
    print(f"{ncpu} host CPUs; affinity at start: {len(os.sched_getaffinity(0))} CPUs", flush=True)
    try:
        for node in sorted(os.listdir("/sys/devices/system/node")):
            if node.startswith("node"):
                print(node, open(f"/sys/devices/system/node/{node}/cpulist").read().strip(), flush=True)
    except OSError:
        pass
    torch.cuda.init()
    bench(gb_in, gb_out, 256, "default affinity")
    bench(gb_in, gb_out, 32, "default affinity")
    all_cpus = sorted(os.sched_getaffinity(0))
    half = len(all_cpus) // 2
    for name, cpus in (("first half of the CPUs", all_cpus[:half]), ("second half of the CPUs", all_cpus[half:])):
        if not cpus:
            continue
        os.sched_setaffinity(0, cpus)
        bench(gb_in, gb_out, 256, name)
    os.sched_setaffinity(0, all_cpus)


if __name__ == "__main__":
    main()
