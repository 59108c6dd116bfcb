"""On-disk record formats of a dbeel SSTable, host side (numpy / bytes).

Mirrors the reference's types (paths under /root/reference):

* ``Entry`` / ``EntryValue``  -- src/storage_engine/mod.rs:52-73, serialized with bincode 1.3.3
  ``DefaultOptions + reject_trailing_bytes + fixint`` (src/utils/bincode.rs:8-16) and the
  timestamp as i128 unix nanos (src/utils/timestamp_nanos.rs:6-12):

      .data record  =  klen:u64le | key | dlen:u64le | data | ts:i128le      (32+klen+dlen bytes)

* ``EntryOffset``             -- src/storage_engine/mod.rs:45-50, exactly 16 bytes
  (lsm_tree.rs:408-413), written by EntryWriter::write (entry_writer.rs:76-86):

      .index record =  offset:u64le | key_size:u32le (=8+klen) | full_size:u32le

* ``TOMBSTONE``               -- mod.rs:14, an entry whose data is empty.

A *run* here is the pair ``(data, index)`` of uint8 arrays holding one SSTable's two files.
An *arrival batch* (input of a memtable flush) uses the same two-array layout, except that
records are in arrival order and keys may repeat.
"""
from __future__ import annotations

import os
from typing import Iterable, List, Sequence, Tuple

import numpy as np

INDEX_ENTRY_SIZE = 16  # mod.rs:33
ENTRY_OVERHEAD = 32  # two u64 length prefixes + i128 timestamp
INDEX_PADDING = 20  # mod.rs:21 -- file names are "{index:020}.{ext}"
DATA_FILE_EXT = "data"
INDEX_FILE_EXT = "index"
BLOOM_FILE_EXT = "bloom"
COMPACT_DATA_FILE_EXT = "compact_data"
COMPACT_INDEX_FILE_EXT = "compact_index"
COMPACT_BLOOM_FILE_EXT = "compact_bloom"
COMPACT_ACTION_FILE_EXT = "compact_action"

EntryTuple = Tuple[bytes, bytes, int]  # (key, data, timestamp nanos)
Run = Tuple[np.ndarray, np.ndarray]


def encode_entry(key: bytes, data: bytes, ts: int) -> bytes:
    return (len(key).to_bytes(8, "little") + key + len(data).to_bytes(8, "little") + data
            + int(ts).to_bytes(16, "little", signed=True))


def encode_index_record(offset: int, klen: int, full_size: int) -> bytes:
    return offset.to_bytes(8, "little") + (8 + klen).to_bytes(4, "little") + full_size.to_bytes(4, "little")


def build_run(entries: Iterable[EntryTuple]) -> Run:
    """Serialize entries in the given order, exactly as EntryWriter::write would
    (entry_writer.rs:71-98).  Sorting / uniqueness are the caller's business."""
    d = bytearray()
    i = bytearray()
    for key, data, ts in entries:
        rec = encode_entry(key, data, ts)
        i += encode_index_record(len(d), len(key), len(rec))
        d += rec
    return (np.frombuffer(bytes(d), dtype=np.uint8).copy(), np.frombuffer(bytes(i), dtype=np.uint8).copy())


def parse_run(data, index) -> List[EntryTuple]:
    """Decode a run through its .index file (the binary_search access path, lsm_tree.rs:605-670)."""
    d = bytes(np.asarray(data, dtype=np.uint8))
    ix = bytes(np.asarray(index, dtype=np.uint8))
    out = []
    for r in range(len(ix) // INDEX_ENTRY_SIZE):
        off = int.from_bytes(ix[16 * r:16 * r + 8], "little")
        ks = int.from_bytes(ix[16 * r + 8:16 * r + 12], "little")
        fs = int.from_bytes(ix[16 * r + 12:16 * r + 16], "little")
        rec = d[off:off + fs]
        klen = int.from_bytes(rec[:8], "little")
        assert ks == 8 + klen, (r, ks, klen)
        dlen = int.from_bytes(rec[8 + klen:16 + klen], "little")
        assert fs == ENTRY_OVERHEAD + klen + dlen, (r, fs, klen, dlen)
        out.append((rec[8:8 + klen], rec[16 + klen:16 + klen + dlen],
                    int.from_bytes(rec[16 + klen + dlen:], "little", signed=True)))
    return out


def build_run_fixed(keys: np.ndarray, docs: np.ndarray, ts: np.ndarray, tombstone: np.ndarray) -> Run:
    """Vectorised run builder for equal-length keys and docs.

    keys [n, klen] uint8, docs [n, dlen] uint8, ts [n] int64 (nanos, fits the low 8 bytes;
    the high 8 bytes are the sign extension), tombstone [n] bool (those rows get dlen = 0).
    """
    n, klen = keys.shape
    dlen = docs.shape[1]
    dense = np.empty((n, ENTRY_OVERHEAD + klen + dlen), dtype=np.uint8)
    dense[:, 16 + klen:16 + klen + dlen] = docs
    return build_run_dense(dense, keys, ts, tombstone)


def build_run_dense(dense: np.ndarray, keys: np.ndarray, ts: np.ndarray, tombstone: np.ndarray) -> Run:
    """Like build_run_fixed, but the caller hands in the [n, 32+klen+dlen] row buffer with
    the document bytes already in place (columns 16+klen .. 16+klen+dlen); the fixed
    fields are filled in here.  Avoids one pass over the payload."""
    n, klen = keys.shape
    full = dense.shape[1]
    dlen = full - ENTRY_OVERHEAD - klen
    dense[:, 0:8] = np.frombuffer(np.uint64(klen).tobytes(), dtype=np.uint8)
    dense[:, 8:8 + klen] = keys
    dense[:, 8 + klen:16 + klen] = np.frombuffer(np.uint64(dlen).tobytes(), dtype=np.uint8)
    ts64 = np.ascontiguousarray(ts, dtype="<i8")
    dense[:, full - 16:full - 8] = ts64.view(np.uint8).reshape(n, 8)
    dense[:, full - 8:] = np.where(ts64 < 0, 0xFF, 0).astype(np.uint8)[:, None]

    sizes = np.where(tombstone, full - dlen, full).astype(np.uint64)
    offsets = np.zeros(n, dtype=np.uint64)
    np.cumsum(sizes[:-1], out=offsets[1:])
    index = np.empty((n, 4), dtype="<u4")
    index[:, 0:2] = offsets.astype("<u8").view("<u4").reshape(n, 2)
    index[:, 2] = 8 + klen
    index[:, 3] = sizes.astype("<u4")

    tomb_rows = np.flatnonzero(tombstone)
    flat = dense.reshape(-1)
    if tomb_rows.size == 0:
        return flat, index.view(np.uint8).reshape(-1)
    # compact in place, front to back: every row only ever moves towards lower addresses
    head = 16 + klen
    prev = 0  # first row not yet placed
    for r in tomb_rows.tolist():
        dst = int(offsets[prev])
        if dst != prev * full:
            flat[dst:dst + (r - prev) * full] = flat[prev * full:r * full].copy()
        o = int(offsets[r])
        row = dense[r].copy()
        flat[o:o + head] = row[:head]
        flat[o + 8 + klen:o + head] = 0  # dlen = 0
        flat[o + head:o + head + 16] = row[full - 16:]
        prev = r + 1
    if prev < n:
        dst = int(offsets[prev])
        flat[dst:dst + (n - prev) * full] = flat[prev * full:n * full].copy()
    total = int(sizes.sum())
    return flat[:total], index.view(np.uint8).reshape(-1)


PAGE_SIZE = 4096  # src/storage_engine/mod.rs (DMA block): every WAL write starts on a page boundary


def build_wal(entries: Iterable[EntryTuple], pad_byte: int = 0) -> np.ndarray:
    """The write-ahead log the reference leaves behind for these writes (lsm_tree.rs:740-744,805-811): each
    bincode Entry at a 4096-aligned offset, padded to ``size + 4096 - size % 4096`` bytes (a whole extra page when
    the size is a page multiple).  ``pad_byte`` fills the padding (the reference's DMA buffers are not cleared)."""
    parts = []
    for k, v, ts in entries:
        e = encode_entry(k, v, ts)
        padded = len(e) + PAGE_SIZE - len(e) % PAGE_SIZE
        parts.append(e + bytes([pad_byte]) * (padded - len(e)))
    return np.frombuffer(b"".join(parts), dtype=np.uint8).copy() if parts else np.zeros(0, np.uint8)


def run_entry_count(index) -> int:
    return len(index) // INDEX_ENTRY_SIZE  # lsm_tree.rs:452-453, 978-979


def file_name(index: int, ext: str) -> str:
    """lsm_tree.rs:284-288 get_file_path: zero-padded to 20 digits."""
    return f"{index:0{INDEX_PADDING}d}.{ext}"


def read_run_files(directory: str, index: int) -> Run:
    d = np.fromfile(os.path.join(directory, file_name(index, DATA_FILE_EXT)), dtype=np.uint8)
    i = np.fromfile(os.path.join(directory, file_name(index, INDEX_FILE_EXT)), dtype=np.uint8)
    return d, i


def write_run_files(directory: str, index: int, run: Run, bloom=None) -> None:
    np.asarray(run[0], dtype=np.uint8).tofile(os.path.join(directory, file_name(index, DATA_FILE_EXT)))
    np.asarray(run[1], dtype=np.uint8).tofile(os.path.join(directory, file_name(index, INDEX_FILE_EXT)))
    if bloom is not None:
        np.asarray(bloom, dtype=np.uint8).tofile(os.path.join(directory, file_name(index, BLOOM_FILE_EXT)))


def input_bytes(runs: Sequence[Run]) -> int:
    """BASELINE.json's unit: MB/s of input bytes = sum over runs of len(.data)+len(.index)."""
    return int(sum(len(d) + len(i) for d, i in runs))
