"""ctypes front-end of the CPU oracle (oracle/dbeel_oracle.c).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package (dbeel_b200/) never imports this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

BLOOM_FP = 0.01  # lsm_tree.rs:48 BLOOM_MAX_ALLOWED_ERROR
DEFAULT_BLOOM_MIN_SIZE = 1_048_576  # mod.rs:19
DEFAULT_TREE_CAPACITY = 8192  # mod.rs:18


class _Run(C.Structure):
    _fields_ = [("data", C.c_void_p), ("data_len", C.c_uint64),
                ("index", C.c_void_p), ("index_len", C.c_uint64)]


class _Out(C.Structure):
    _fields_ = [("data", C.c_void_p), ("data_cap", C.c_uint64), ("data_len", C.c_uint64),
                ("index", C.c_void_p), ("index_cap", C.c_uint64), ("index_len", C.c_uint64),
                ("bloom", C.c_void_p), ("bloom_cap", C.c_uint64), ("bloom_len", C.c_uint64),
                ("items_written", C.c_uint64)]


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "dbeel_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.orc_compact.restype = C.c_int
        L.orc_compact.argtypes = [C.POINTER(_Run), C.c_uint32, C.c_int, C.c_uint64, C.c_double,
                                  C.c_char_p, C.c_int, C.POINTER(_Out)]
        L.orc_siphash13.restype = C.c_uint64
        L.orc_siphash13.argtypes = [C.c_uint64, C.c_uint64, C.c_char_p, C.c_uint64]
        L.orc_bloom_bitmap_bytes.restype = C.c_uint64
        L.orc_bloom_bitmap_bytes.argtypes = [C.c_uint64, C.c_double]
        L.orc_bloom_k_num.restype = C.c_uint32
        L.orc_bloom_k_num.argtypes = [C.c_uint64, C.c_uint64]
        L.orc_bloom_file_size.restype = C.c_uint64
        L.orc_bloom_file_size.argtypes = [C.c_uint64, C.c_double]
        L.orc_bloom_check.restype = C.c_int
        L.orc_bloom_check.argtypes = [C.c_void_p, C.c_uint64, C.c_char_p, C.c_uint64]
        L.orc_rb_new.restype = C.c_void_p
        L.orc_rb_new.argtypes = [C.c_uint32]
        L.orc_rb_free.argtypes = [C.c_void_p]
        L.orc_rb_len.restype = C.c_uint32
        L.orc_rb_len.argtypes = [C.c_void_p]
        L.orc_rb_clear.argtypes = [C.c_void_p]
        L.orc_rb_set.restype = C.c_int
        L.orc_rb_set.argtypes = [C.c_void_p, C.c_char_p, C.c_uint64, C.c_char_p, C.c_uint64, C.c_char_p]
        L.orc_rb_shape.restype = C.c_uint32
        L.orc_rb_shape.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
        L.orc_rb_flush.restype = C.c_int
        L.orc_rb_flush.argtypes = [C.c_void_p, C.c_int, C.POINTER(_Out)]
        L.orc_memtable_flush.restype = C.c_int
        L.orc_memtable_flush.argtypes = [C.POINTER(_Run), C.c_uint64, C.c_uint32, C.c_int,
                                         C.POINTER(_Out), C.POINTER(C.c_uint64)]
        L.orc_sstable_lookup.restype = C.c_int
        L.orc_sstable_lookup.argtypes = [C.POINTER(_Run), C.c_void_p, C.c_uint64, C.c_char_p, C.c_uint64,
                                         C.POINTER(C.c_uint64), C.POINTER(C.c_int)]
        L.orc_timestamp_decodes.restype = C.c_int
        L.orc_timestamp_decodes.argtypes = [C.c_char_p]
        L.orc_wal_flush.restype = C.c_int
        L.orc_wal_flush.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_int, C.POINTER(_Out), C.POINTER(C.c_uint64)]
        L.orc_get_many.restype = None
        L.orc_get_many.argtypes = [C.POINTER(_Run), C.POINTER(C.c_void_p), C.POINTER(C.c_uint64), C.c_uint32, C.c_void_p,
                                   C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_densify.restype = C.c_uint64
        L.orc_densify.argtypes = [C.POINTER(_Run), C.c_void_p, C.c_uint64, C.c_void_p]
        L.orc_murmur3_32.restype = C.c_uint32
        L.orc_murmur3_32.argtypes = [C.c_char_p, C.c_uint64, C.c_uint32]
        L.orc_ring_owner.restype = C.c_uint32
        L.orc_ring_owner.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32]
        L.orc_route.restype = C.c_uint64
        L.orc_route.argtypes = [C.POINTER(_Run), C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def _u8(a) -> np.ndarray:
    if isinstance(a, np.ndarray):
        assert a.dtype == np.uint8 and a.flags.c_contiguous
        return a
    return np.frombuffer(bytes(a), dtype=np.uint8)


def _mk_runs(runs: Sequence[Tuple[object, object]]):
    keep = [(_u8(d), _u8(i)) for d, i in runs]
    arr = (_Run * max(1, len(keep)))()
    for j, (d, i) in enumerate(keep):
        arr[j] = _Run(d.ctypes.data, d.size, i.ctypes.data, i.size)
    return arr, keep


def _mk_out(data_cap: int, index_cap: int, bloom_cap: int):
    d = np.empty(max(1, data_cap), dtype=np.uint8)
    i = np.empty(max(1, index_cap), dtype=np.uint8)
    b = np.empty(max(1, bloom_cap), dtype=np.uint8)
    o = _Out(d.ctypes.data, data_cap, 0, i.ctypes.data, index_cap, 0, b.ctypes.data, bloom_cap, 0, 0)
    return o, (d, i, b)


class OracleError(RuntimeError):
    pass


def compact(runs: Sequence[Tuple[object, object]], keep_tombstones: bool,
            bloom_min_size: int = DEFAULT_BLOOM_MIN_SIZE, seed: bytes = bytes(32),
            emulate_page_cache: bool = False, fp: float = BLOOM_FP):
    """LSMTree::compact merge core.  Returns (data, index, bloom|None, items_written) as
    numpy uint8 arrays (views trimmed to the written length)."""
    assert len(seed) == 32
    arr, keep = _mk_runs(runs)
    data_cap = sum(d.size for d, _ in keep)
    index_cap = sum(i.size for _, i in keep)
    items = sum(i.size // 16 for _, i in keep)
    bloom_cap = lib().orc_bloom_file_size(items, fp) if items and data_cap > bloom_min_size else 0
    out, (d, i, b) = _mk_out(data_cap, index_cap, bloom_cap)
    rc = lib().orc_compact(arr, len(keep), int(keep_tombstones), bloom_min_size, fp, seed,
                           int(emulate_page_cache), C.byref(out))
    if rc:
        raise OracleError(f"orc_compact rc={rc}")
    bloom = b[:out.bloom_len] if out.bloom_len else None
    return d[:out.data_len], i[:out.index_len], bloom, int(out.items_written)


def memtable_flushes(batch: Tuple[object, object], capacity: int = DEFAULT_TREE_CAPACITY,
                     emulate_page_cache: bool = False):
    """Replay an arrival-ordered batch through the red-black-tree memtable, flushing every
    time it fills (and once more for the remainder).  Returns a list of (data, index, n)."""
    arr, keep = _mk_runs([batch])
    d0, i0 = keep[0]
    n = i0.size // 16
    pos = 0
    outs = []
    while pos < n:
        out, (d, i, _) = _mk_out(d0.size, i0.size, 0)
        consumed = C.c_uint64(0)
        rc = lib().orc_memtable_flush(arr, pos, capacity, int(emulate_page_cache), C.byref(out),
                                      C.byref(consumed))
        if rc:
            raise OracleError(f"orc_memtable_flush rc={rc}")
        outs.append((d[:out.data_len].copy(), i[:out.index_len].copy(), int(out.items_written)))
        pos += consumed.value
    return outs


def siphash13(k0: int, k1: int, msg: bytes) -> int:
    return lib().orc_siphash13(k0, k1, msg, len(msg))


def bloom_bitmap_bytes(items: int, fp: float = BLOOM_FP) -> int:
    return lib().orc_bloom_bitmap_bytes(items, fp)


def bloom_k_num(bits: int, items: int) -> int:
    return lib().orc_bloom_k_num(bits, items)


def bloom_file_size(items: int, fp: float = BLOOM_FP) -> int:
    return lib().orc_bloom_file_size(items, fp)


def bloom_check(bloom_file, key: bytes) -> bool:
    b = _u8(bloom_file)
    r = lib().orc_bloom_check(b.ctypes.data, b.size, key, len(key))
    if r < 0:
        raise OracleError("malformed bloom file")
    return bool(r)


class RbTree:
    """rbtree_arena::RedBlackTree<Vec<u8>, EntryValue> restatement (shape-observable)."""

    def __init__(self, capacity: int):
        self._t = lib().orc_rb_new(capacity)
        self._keep = []

    def __del__(self):
        if getattr(self, "_t", None):
            lib().orc_rb_free(self._t)
            self._t = None

    def __len__(self):
        return lib().orc_rb_len(self._t)

    def clear(self):
        lib().orc_rb_clear(self._t)
        self._keep.clear()

    def set(self, key: bytes, value: bytes, ts: int = 0) -> Optional[bool]:
        """True if an existing key was replaced, False if inserted; raises when full."""
        k = C.create_string_buffer(key, len(key))
        v = C.create_string_buffer(value, len(value))
        self._keep += [k, v]
        r = lib().orc_rb_set(self._t, k, len(key), v, len(value), int(ts).to_bytes(16, "little", signed=True))
        if r < 0:
            raise OracleError("ReachedCapacity")
        return bool(r)

    def shape(self):
        buf = np.zeros(4096, dtype=np.uint8)
        n = lib().orc_rb_shape(self._t, buf.ctypes.data, buf.size)
        return [(int(buf[j]), int(buf[j + 1])) for j in range(0, n, 2)]

    def flush(self, cap_bytes: int = 1 << 20):
        out, (d, i, _) = _mk_out(cap_bytes, cap_bytes, 0)
        rc = lib().orc_rb_flush(self._t, 0, C.byref(out))
        if rc:
            raise OracleError(f"orc_rb_flush rc={rc}")
        return d[:out.data_len].copy(), i[:out.index_len].copy(), int(out.items_written)


def sstable_lookup(run, bloom, key: bytes):
    """get_entry's per-SSTable step: (found, index record number | None, bloom_said_no)."""
    arr, keep = _mk_runs([run])
    b = _u8(bloom) if bloom is not None else None
    rec, no = C.c_uint64(0), C.c_int(0)
    f = lib().orc_sstable_lookup(arr, b.ctypes.data if b is not None else None, b.size if b is not None else 0,
                                 key, len(key), C.byref(rec), C.byref(no))
    return bool(f), (int(rec.value) if f else None), bool(no.value)


def get_many(tables, keys_blob: np.ndarray, key_off: np.ndarray):
    """get_entry's SSTable loop for a batch: tables = [(data, index, bloom | None)] oldest first, keys packed as
    (bytes back to back, n + 1 uint64 offsets).  Returns (table int32[n], record uint64[n], bloom_rejects uint32[n])."""
    arr, keep = _mk_runs([(d, i) for d, i, _ in tables])
    nt = len(tables)
    blooms = [(_u8(b) if b is not None and len(b) else None) for _, _, b in tables]
    bp = (C.c_void_p * max(1, nt))(*[(b.ctypes.data if b is not None else None) for b in blooms])
    bl = (C.c_uint64 * max(1, nt))(*[(b.size if b is not None else 0) for b in blooms])
    n = key_off.size - 1
    blob = np.ascontiguousarray(keys_blob, dtype=np.uint8)
    off = np.ascontiguousarray(key_off, dtype=np.uint64)
    t, r, j = np.empty(n, np.int32), np.empty(n, np.uint64), np.empty(n, np.uint32)
    lib().orc_get_many(arr, bp, bl, nt, blob.ctypes.data if blob.size else None, off.ctypes.data, n, t.ctypes.data,
                       r.ctypes.data, j.ctypes.data)
    return t, r, j


def timestamp_decodes(ts: int) -> bool:
    """utils/timestamp_nanos.rs:15-24 through time 0.3: does this i128 nanosecond count deserialize?"""
    return bool(lib().orc_timestamp_decodes(int(ts).to_bytes(16, "little", signed=True)))


def wal_flush(wal, capacity: int = DEFAULT_TREE_CAPACITY, emulate_page_cache: bool = False):
    """read_memtable_from_wal_file (lsm_tree.rs:552-574) + the recovery flush (:478-513).
    Returns (data, index, items_written, entries_replayed); raises OracleError("ReachedCapacity") like `set(..)?`."""
    w = _u8(wal)
    out, (d, i, _) = _mk_out(w.size, w.size // 16 + 16, 0)
    seen = C.c_uint64(0)
    rc = lib().orc_wal_flush(w.ctypes.data if w.size else None, w.size, capacity, int(emulate_page_cache), C.byref(out),
                             C.byref(seen))
    if rc == 4:
        raise OracleError("ReachedCapacity")
    if rc:
        raise OracleError(f"orc_wal_flush rc={rc}")
    return d[:out.data_len].copy(), i[:out.index_len].copy(), int(out.items_written), int(seen.value)


def murmur3_32(data: bytes, seed: int = 0) -> int:
    """murmur3 0.5.2's murmur3_32 = MurmurHash3_x86_32 (shards.rs:95-101 call it with seed 0)."""
    return int(lib().orc_murmur3_32(bytes(data), len(data), seed))


def shard_ring(n_shards: int, node: str = "dbeel") -> np.ndarray:
    """Ascending hashes of the shard names "<node>-<cpu id>" (shards.rs:213-214, args.rs default name) and the cpu id at
    each ring position: returns (hashes[u32], ids[u32])."""
    pairs = sorted((murmur3_32(f"{node}-{i}".encode()), i) for i in range(n_shards))
    return np.array([h for h, _ in pairs], np.uint32), np.array([i for _, i in pairs], np.uint32)


def ring_owner(ring: np.ndarray, key_hash: int) -> int:
    ring = np.ascontiguousarray(ring, np.uint32)
    return int(lib().orc_ring_owner(ring.ctypes.data, ring.size, key_hash))


def route(batch: Tuple[object, object], ring: np.ndarray):
    """owns_key (shards.rs:586-598) for every arrival of a batch: (ring position per arrival, key hash per arrival)."""
    arr, keep = _mk_runs([batch])
    ring = np.ascontiguousarray(ring, np.uint32)
    n = keep[0][1].size // 16
    shard = np.empty(n, np.uint32)
    hashes = np.empty(n, np.uint32)
    done = lib().orc_route(arr, ring.ctypes.data, ring.size, shard.ctypes.data, hashes.ctypes.data)
    return shard[:done], hashes[:done]


def densify(data, sparse_index):
    """Dense arrival batch (data, index) of the entries a sparse index names inside `data` (a routed stream's slice)."""
    arr, keep = _mk_runs([(data, sparse_index)])
    idx = keep[0][1]
    fs = idx.reshape(-1, 16)[:, 12:16].copy().view("<u4").ravel()
    total = int(fs.astype(np.uint64).sum())
    od = np.empty(max(1, total), np.uint8)
    oi = np.empty(max(1, idx.size), np.uint8)
    done = lib().orc_densify(arr, od.ctypes.data, total, oi.ctypes.data)
    if done != idx.size // 16:
        raise OracleError(f"orc_densify stopped at record {done}")
    return od[:total], oi[:idx.size]
