"""CPU checks of the scalar building blocks the CUDA kernels are made of.

dbeel_b200/csrc/device_fns.cuh is compiled twice from the same text: by nvcc as device code
and here by g++ into tests/_host_shim.so.  No GPU is involved and nothing here is a product
path -- it pins the arithmetic (merge-record order, SipHash pair, modular reduction, byte
realignment, i128 order) against Python / the oracle before the kernels ever run."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle
from helpers import nasty_keys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.fixture(scope="module")
def shim():
    so = os.path.join(HERE, "_host_shim.so")
    src = os.path.join(HERE, "host_shim.cc")
    hdr = os.path.join(ROOT, "dbeel_b200", "csrc", "device_fns.cuh")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-x", "c++", "-o", so, src])
    L = C.CDLL(so)
    L.shim_fastmod.restype = C.c_uint64
    L.shim_fastmod.argtypes = [C.c_uint64, C.c_uint64]
    L.shim_bloom_hash_i.restype = C.c_uint64
    L.shim_bloom_hash_i.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32]
    L.shim_bloom_probe_all.restype = C.c_uint32
    L.shim_bloom_probe_all.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint64, C.POINTER(C.c_uint64)]
    L.shim_make_rec.argtypes = [C.c_char_p, C.c_uint64, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint32)]
    L.shim_rec_cmp.restype = C.c_int
    L.shim_rec_cmp.argtypes = [C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_int)]
    L.shim_sip_pair.argtypes = [C.POINTER(C.c_uint64), C.c_char_p, C.c_uint64, C.POINTER(C.c_uint64)]
    L.shim_realign16.argtypes = [C.c_char_p, C.c_uint32, C.c_char_p]
    L.shim_window32.argtypes = [C.c_char_p, C.c_uint32, C.c_char_p]
    L.shim_blend32.argtypes = [C.c_char_p, C.c_char_p, C.c_uint32, C.c_char_p]
    L.shim_ts_decodes.restype = C.c_int
    L.shim_ts_decodes.argtypes = [C.c_char_p]
    L.shim_ts_greater.restype = C.c_int
    L.shim_ts_greater.argtypes = [C.c_char_p, C.c_char_p]
    L.shim_murmur3_32.restype = C.c_uint32
    L.shim_murmur3_32.argtypes = [C.c_char_p, C.c_uint64, C.c_uint32]
    L.shim_ring_owner.restype = C.c_uint32
    L.shim_ring_owner.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32]
    return L


def _rec(shim, key: bytes, L: int, gid: int = 0):
    out = (C.c_uint32 * 4)()
    shim.shim_make_rec(key + b"\xEE" * 24, len(key), L, gid, out)  # trailing garbage must be ignored
    return out


def test_merge_record_order_matches_bytes_order(shim):
    rng = np.random.default_rng(3)
    for L, stem in [(0, b""), (5, b"\xb0k000"), (33, b"common-prefix-that-is-quite-long/")]:
        keys = [stem + k for k in nasty_keys(rng, 300, max_len=30)] + [stem, stem + b"\x00", stem + b"\x00" * 11,
                                                                      stem + b"\x00" * 12, stem + b"a" * 11,
                                                                      stem + b"a" * 12, stem + b"a" * 13]
        keys = sorted(set(keys))
        recs = [_rec(shim, k, L) for k in keys]
        und = C.c_int()
        n_und = 0
        for i in range(len(keys)):
            for j in range(max(0, i - 6), min(len(keys), i + 7)):
                c = shim.shim_rec_cmp(recs[i], recs[j], C.byref(und))
                exp = (keys[i] > keys[j]) - (keys[i] < keys[j])
                if und.value:
                    n_und += 1
                    # undecided only when both keys run past the 11-byte window and agree on it
                    assert len(keys[i]) - L > 11 and len(keys[j]) - L > 11
                    assert keys[i][:L + 11] == keys[j][:L + 11]
                else:
                    assert c == exp, (keys[i], keys[j], L)
        assert n_und > 0


def test_sip_pair_matches_oracle_and_cpython_pinned_core(shim):
    rng = np.random.default_rng(4)
    k = (C.c_uint64 * 4)(0x0706050403020100, 0x0F0E0D0C0B0A0908, 0x1716151413121110, 0x1F1E1D1C1B1A1918)
    out = (C.c_uint64 * 2)()
    for klen in list(range(0, 40)) + [63, 64, 65, 255, 256, 1000]:
        key = bytes(rng.integers(0, 256, klen, dtype=np.uint8))
        shim.shim_sip_pair(k, key + b"\x55" * 16, klen, out)
        msg = klen.to_bytes(8, "little") + key  # Hash for Vec<u8>: write_usize(len) then bytes
        assert out[0] == oracle.siphash13(k[0], k[1], msg)
        assert out[1] == oracle.siphash13(k[2], k[3], msg)


def test_fastmod_and_double_hashing(shim):
    rng = np.random.default_rng(5)
    P = 0xFFFFFFFFFFFFFFC5
    for d in [8, 24, 628_168, 38_340_240, 76_680_472, (1 << 40) + 8, (1 << 63) + 8]:
        hs = [0, 1, d - 1, d, d + 1, 2 * d - 1, 2 * d, (1 << 64) - 1, (1 << 64) - d] + \
             [int(x) for x in rng.integers(0, 1 << 63, 200, dtype=np.uint64)]
        for h in hs:
            h &= (1 << 64) - 1
            assert shim.shim_fastmod(h, d) == h % d
    for _ in range(200):
        h0, h1 = (int(x) for x in rng.integers(0, 1 << 64, 2, dtype=np.uint64))
        for i in range(0, 9):
            exp = h0 if i == 0 else h1 if i == 1 else ((h0 + i * h1) & ((1 << 64) - 1)) % P
            assert shim.shim_bloom_hash_i(h0, h1, i) == exp
    out = (C.c_uint64 * 16)()
    for _ in range(300):  # the incremental form used by the kernel == bloomfilter's (h0 + i*h1) % prime % bits
        h0, h1 = (int(x) for x in rng.integers(0, 1 << 64, 2, dtype=np.uint64))
        if _ % 3 == 0:
            h1 = (1 << 64) - int(rng.integers(1, 1000))  # wrap-around heavy
        for k, bits in [(7, 76_680_472), (1, 8), (2, 24), (13, (1 << 40) + 8)]:
            n = shim.shim_bloom_probe_all(h0, h1, k, bits, out)
            exp = [(h0 if i == 0 else h1 if i == 1 else ((h0 + i * h1) & ((1 << 64) - 1)) % P) % bits for i in range(k)]
            assert n == k and [out[i] for i in range(k)] == exp
    assert shim.shim_bloom_hash_i(P, 0, 2) == 0 and shim.shim_bloom_hash_i((1 << 64) - 1, 0, 5) == ((1 << 64) - 1) % P


def test_realign16(shim):
    src = bytes(range(100, 132))
    out = C.create_string_buffer(16)
    for sh in range(16):
        shim.shim_realign16(src, sh, out)
        assert out.raw == src[sh:sh + 16]


def test_i128_order(shim):
    vals = [0, 1, -1, 1 << 64, -(1 << 64), (1 << 127) - 1, -(1 << 127), 1_700_000_000_000_000_000, (1 << 64) - 1, 1 << 63]
    for a in vals:
        for b in vals:
            got = shim.shim_ts_greater(a.to_bytes(16, "little", signed=True), b.to_bytes(16, "little", signed=True))
            assert got == int(a > b)


def test_timestamp_range_check_matches_oracle(shim):
    """ts_decodes (WAL replay kernel) == the oracle's restatement of time's from_unix_timestamp_nanos, including the
    wrapping i64 cast of the seconds."""
    rng = np.random.default_rng(9)
    lo, hi = -377705116800 * 10**9, 253402300799 * 10**9 + 999_999_999
    vals = [0, 1, -1, lo, lo - 1, lo + 1, hi, hi + 1, hi - 1, 1 << 100, -(1 << 100), (1 << 127) - 1, -(1 << 127),
            10**9 << 64, -(10**9 << 64), (10**9 << 64) + 5, 999_999_999, -999_999_999, 10**9, -(10**9)]
    vals += [int(rng.integers(-2**62, 2**62)) * int(rng.integers(1, 2**40)) for _ in range(3000)]
    vals += [int(rng.integers(-2**63, 2**63)) << int(rng.integers(0, 64)) for _ in range(3000)]
    for v in vals:
        v = max(-(1 << 127), min((1 << 127) - 1, v))
        assert bool(shim.shim_ts_decodes(v.to_bytes(16, "little", signed=True))) == oracle.timestamp_decodes(v), v


def test_murmur3_32_and_ring_owner(shim):
    """The routing kernel's arithmetic (device_fns.cuh) against the sklearn-generated goldens and the oracle's ring walk."""
    import json
    g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "murmur3_32.json")))
    for v in g["vectors"]:
        b = bytes.fromhex(v["hex"])
        assert shim.shim_murmur3_32(b + b"\x77" * 9, len(b), v["seed"]) == v["hash"], v  # garbage past the end is ignored
    ring, _ = oracle.shard_ring(8)
    rng = np.random.default_rng(4)
    probes = [0, 1, 2**32 - 1] + [int(h) + d for h in ring for d in (-1, 0, 1)] + [int(x) for x in rng.integers(0, 2**32, 3000)]
    for n in (1, 2, 3, 8):
        sub = np.ascontiguousarray(ring[:n])
        for h in probes:
            h %= 2**32
            assert shim.shim_ring_owner(sub.ctypes.data, n, h) == oracle.ring_owner(sub, h), (n, h)


def test_window32_and_blend32(shim):
    """The 32-byte realignment of the gather's entry-boundary blocks: any 32 bytes out of a 64-byte window, and the blend of the
    tail of one entry with the head of the next at any byte."""
    rng = np.random.default_rng(17)
    for _ in range(200):
        src = bytes(rng.integers(0, 256, 64, dtype=np.uint8))
        for s0 in range(32):
            out = C.create_string_buffer(32)
            shim.shim_window32(src, s0, out)
            assert out.raw == src[s0:s0 + 32], s0
    for _ in range(50):
        t32, h32 = bytes(rng.integers(0, 256, 32, dtype=np.uint8)), bytes(rng.integers(0, 256, 32, dtype=np.uint8))
        for t in range(33):
            out = C.create_string_buffer(32)
            shim.shim_blend32(t32, h32, t, out)
            assert out.raw == t32[:t] + h32[t:], t
