#!/usr/bin/env python
"""Which part of the host link do GPUs share?  Simultaneous H2D + D2H (pinned host memory allocated next to each GPU's
NUMA node, 2.55 GB down / 2.06 GB up per round like one cfg2 job) on several SETS of GPUs at once: one GPU alone, two GPUs
that are neighbours on the PCIe tree, two that are not, four, all eight.  Prints per-set aggregate and per-GPU rates.
Usage (on an 8-GPU box): tools/pcie_topology.py"""
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from dbeel_b200 import capi  # noqa: E402

GB_IN, GB_OUT, ROUNDS = 2.547, 2.059, 4


class Gpu:
    def __init__(self, idx):
        self.idx = idx
        self.node = None

    def alloc(self):  # runs on its own thread: the thread (and its pinned pages) move next to the GPU
        self.node, _ = capi.bind_to_gpu(self.idx)
        dev = torch.device("cuda", self.idx)
        self.h_in = torch.empty(int(GB_IN * 1e9), dtype=torch.uint8).pin_memory()
        self.h_out = torch.empty(int(GB_OUT * 1e9), dtype=torch.uint8).pin_memory()
        self.h_in.fill_(1)
        self.h_out.fill_(2)
        self.d_in = torch.empty_like(self.h_in, device=dev)
        self.d_out = torch.ones_like(self.h_out, device=dev)
        self.s1, self.s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def go(self, barrier, out):
        dev = torch.device("cuda", self.idx)
        torch.cuda.set_device(dev)
        barrier.wait()
        t0 = time.perf_counter()
        for _ in range(ROUNDS):
            with torch.cuda.stream(self.s1):
                self.d_in.copy_(self.h_in, non_blocking=True)
            with torch.cuda.stream(self.s2):
                self.h_out.copy_(self.d_out, non_blocking=True)
        self.s1.synchronize()
        self.s2.synchronize()
        out[self.idx] = time.perf_counter() - t0


def main():
    n = torch.cuda.device_count()
    try:
        print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout, flush=True)
    except OSError:
        pass
    gpus = [Gpu(i) for i in range(n)]
    ths = [threading.Thread(target=g.alloc) for g in gpus]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    print("NUMA node of every GPU:", [g.node for g in gpus], flush=True)
    sets = [[0], [0, 1], [0, 2], [0, 4], [0, 1, 2, 3], [0, 2, 4, 6], [4, 5], [6, 7], list(range(n))]
    for s in sets:
        if max(s) >= n:
            continue
        for rep in range(2):  # first pass warms up
            barrier = threading.Barrier(len(s))
            out = {}
            ths = [threading.Thread(target=gpus[i].go, args=(barrier, out)) for i in s]
            for t in ths:
                t.start()
            for t in ths:
                t.join()
        per = {i: (GB_IN + GB_OUT) * ROUNDS / out[i] for i in s}
        tot = sum(per.values())
        print(f"GPUs {s}: aggregate {tot:7.1f} GB/s (in + out), per GPU " + " ".join(f"{i}:{v:5.1f}" for i, v in per.items())
              + f"   H2D share {tot * GB_IN / (GB_IN + GB_OUT):6.1f} GB/s", flush=True)


if __name__ == "__main__":
    main()
