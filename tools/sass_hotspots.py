#!/usr/bin/env python
"""Attribute an ncu SASS-level profile to source lines: pairs the per-instruction counters of
`ncu --page source --csv` with the line table of the locally built cubin (same build).
Usage: tools/sass_hotspots.py <ncu-rep> <kernel-substring> [top] [cubin-section-substring] [inst|smp]"""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
    sect = sys.argv[4] if len(sys.argv) > 4 else kern  # template instances: the mangled name picks one (e.g. k_merge_finalILb1E)
    order = 1 if len(sys.argv) > 5 and sys.argv[5] == "smp" else 0
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", kern],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    # first kernel instance only
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hdr_i]
    ie, ss, src = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Source")
    prof = []
    for r in rows[hdr_i + 1:]:
        if not r or r[0] in ("Kernel Name", "Address"):
            break
        prof.append((r[src].strip(), int(r[ie] or 0), int(r[ss] or 0)))
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "dbeel_b200", "libdbeel_compact.so")], cwd=tmp,
                   capture_output=True)
    dis = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, "dbeel_compact.sm_100a.cubin")],
                         capture_output=True, text=True).stdout.splitlines()
    start = next(i for i, l in enumerate(dis) if l.startswith(".text.") and sect in l)
    lines = []
    cur = None
    for l in dis[start + 1:]:
        if l.startswith("//--------------------- "):
            break
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
            lines.append(cur)
    if len(lines) != len(prof):
        print(f"warning: {len(lines)} SASS instructions locally vs {len(prof)} in the report", file=sys.stderr)
    by_line = defaultdict(lambda: [0, 0])
    tot_i = tot_s = 0
    for (s, n, smp), ln in zip(prof, lines):
        by_line[ln][0] += n
        by_line[ln][1] += smp
        tot_i += n
        tot_s += smp
    srcs = {}
    print(f"{kern}: {tot_i} warp-instructions, {tot_s} samples")
    for ln, (n, smp) in sorted(by_line.items(), key=lambda kv: -kv[1][order])[:top]:
        text = ""
        if ln:
            f = os.path.join(ROOT, "dbeel_b200", "csrc", ln[0])
            if f not in srcs and os.path.exists(f):
                srcs[f] = open(f).read().splitlines()
            if f in srcs and ln[1] <= len(srcs[f]):
                text = srcs[f][ln[1] - 1].strip()[:90]
        print(f"{100 * n / tot_i:5.1f}% inst {100 * smp / max(1, tot_s):5.1f}% smp  {ln}  {text}")


if __name__ == "__main__":
    main()
