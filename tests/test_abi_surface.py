"""CPU-only checks of the drop-in boundary: the built library loads and exports every function
that include/dbeel_compact.h declares; host-side arithmetic entry points agree with the oracle;
without a GPU the engine refuses to exist (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

import oracle
from dbeel_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(ROOT, "include", "dbeel_compact.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dbeel_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    names = declared_functions()
    assert len(names) >= 16 and "dbeel_compact" in names and "dbeel_flush_device" in names
    lib = capi.lib()
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"declared in the header but not exported: {missing}"
    assert sorted(capi.EXPORTS) == [n for n in names if n in capi.EXPORTS]
    assert lib.dbeel_abi_version() == 1


def test_headers_are_plain_c(tmp_path):
    """The boundary is a C ABI: both headers compile as C99 (no C++-isms, no torch types), and a C translation unit
    that uses every struct links against the built library."""
    import subprocess
    src = tmp_path / "use_abi.c"
    src.write_text("""
#include "dbeel_compact.h"
#include "dbeel_tree.h"
#include <stdio.h>
int main(void) {
    dbeel_run r = {0, 0, 0, 0};
    dbeel_out o = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    dbeel_table t = {0, 0, 0, 0, 0, 0};
    dbeel_lookup_result lr = {-1, 0, 0};
    dbeel_flush_table ft = {0, 0, 0, 0, 0};
    dbeel_stats st;
    dbeel_engine *e = 0;
    int rc = dbeel_engine_create(-1, &e); /* no such device: must fail, never fall back */
    (void)r; (void)o; (void)t; (void)lr; (void)ft; (void)st;
    printf("%d %d %s %llu\\n", dbeel_abi_version(), rc != DBEEL_OK, dbeel_strerror(DBEEL_ERR_TREE_FULL),
           (unsigned long long)dbeel_bloom_file_size(1000, 0.01));
    return sizeof(dbeel_lookup_result) == 16 ? 0 : 1;
}
""")
    inc = os.path.join(ROOT, "include")
    subprocess.check_call(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-fsyntax-only", "-I", inc, str(src)])
    exe = tmp_path / "use_abi"
    libdir = os.path.dirname(capi.LIB_PATH)
    capi.lib()  # builds the library if it is not there yet
    subprocess.check_call(["gcc", "-std=c99", "-I", inc, str(src), "-o", str(exe), "-L", libdir, "-ldbeel_compact",
                           "-Wl,-rpath," + libdir])
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    fields = out.stdout.split()
    assert fields[0] == "1" and fields[1] == "1" and int(fields[-1]) == oracle.bloom_file_size(1000)


def test_bloom_sizing_matches_oracle():
    lib = capi.lib()
    for n in [1, 2, 7, 1000, 65_536, 4_000_000, 8_000_000, 123_456_789]:
        for fp in [0.01, 0.001, 0.3]:
            assert lib.dbeel_bloom_bitmap_bytes(n, fp) == oracle.bloom_bitmap_bytes(n, fp)
            bits = oracle.bloom_bitmap_bytes(n, fp) * 8
            assert lib.dbeel_bloom_k_num(bits, n) == oracle.bloom_k_num(bits, n)
            assert lib.dbeel_bloom_file_size(n, fp) == oracle.bloom_file_size(n, fp)


def test_compact_bound():
    opts = capi.make_opts(False)
    d, i, b = capi.compact_bound([(2_000_000, 160_007), (500_000, 32)], opts)
    assert d == 2_500_000 and i == 16 * (10_000 + 2) and b == oracle.bloom_file_size(10_002)
    d, i, b = capi.compact_bound([(1_048_576, 1600)], opts)  # strict '>' (lsm_tree.rs:1027)
    assert b == 0
    assert capi.compact_bound([], opts) == (0, 0, 0)


def test_compact_many_bound_sizes_every_job_on_its_own_inputs():
    """dbeel_compact_many_bound: data / index caps are the sums over all jobs; every job gets a filter iff ITS input
    .data bytes exceed the threshold (lsm_tree.rs:1026-1034), sized for ITS entry count, at a 16-byte aligned offset."""
    lib = capi.lib()
    shapes = [[(3_000_000, 16 * 9000), (10, 16)], [(500, 16 * 3)], [], [(1_048_577, 16 * 1)], [(1_048_576, 16 * 7)]]
    jobs = (capi.Job * len(shapes))()
    keep = []
    for j, runs in enumerate(shapes):
        arr = (capi.Run * max(1, len(runs)))()
        for k, (dl, il) in enumerate(runs):
            arr[k] = capi.Run(1, dl, 1, il)  # never dereferenced by the bound function
        keep.append(arr)
        jobs[j] = capi.Job(arr, len(runs), 0, None)
    dc, ic, bc = C.c_uint64(), C.c_uint64(), C.c_uint64()
    assert lib.dbeel_compact_many_bound(jobs, len(shapes), 1 << 20, 0.01, C.byref(dc), C.byref(ic), C.byref(bc)) == 0
    assert dc.value == sum(dl for runs in shapes for dl, _ in runs)
    assert ic.value == sum(il for runs in shapes for _, il in runs)
    pad16 = lambda n: (n + 15) // 16 * 16
    assert bc.value == pad16(oracle.bloom_file_size(9001)) + pad16(oracle.bloom_file_size(1))  # jobs 0 and 3 only (strict >)
    assert lib.dbeel_compact_many_bound(None, 0, 1 << 20, 0.01, C.byref(dc), C.byref(ic), C.byref(bc)) == 0 and dc.value == 0
    assert lib.dbeel_compact_many_bound(jobs, 1, 1 << 20, 1.5, None, None, None) == capi.ERR_INVALID_ARG


def test_strerror_and_null_handling():
    lib = capi.lib()
    assert lib.dbeel_strerror(0) == b"ok"
    assert lib.dbeel_strerror(capi.ERR_NO_DEVICE) != lib.dbeel_strerror(12345)
    assert lib.dbeel_engine_create(0, None) == capi.ERR_INVALID_ARG
    lib.dbeel_engine_destroy(None)  # no-op
    lib.dbeel_host_free(None)


def test_no_gpu_means_no_engine():
    """The product path must fail loudly without its device: there is no CPU fallback."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(capi.DbeelError) as ei:
        capi.Engine(0)
    assert ei.value.code == capi.ERR_NO_DEVICE


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "dbeel_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cc", ".h")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "import oracle" not in src and "from oracle" not in src and "liboracle" not in src, f
