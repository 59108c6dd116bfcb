#!/usr/bin/env python
"""Golden vectors for murmur3_32 (the crate `murmur3 0.5.2` behind src/shards.rs:95-101 is not vendored).
Source of truth here: scikit-learn's murmurhash3_32 (its own C++ MurmurHash3_x86_32, independent of oracle/ and of the CUDA
code), cross-checked against the published SMHasher verification values listed below.
    python tests/golden/gen_murmur3_sklearn.py > tests/golden/murmur3_32.json"""
import json

import numpy as np
from sklearn.utils import murmurhash3_32

PUBLISHED = [  # (bytes hex, seed, expected) -- SMHasher / reference implementation known answers
    ("", 0, 0x00000000), ("", 1, 0x514E28B7), ("", 0xFFFFFFFF, 0x81F16F39), ("ffffffff", 0, 0x76293B50),
    ("21436587", 0, 0xF55B516B), ("21436587", 0x5082EDEE, 0x2362F9DE), ("214365", 0, 0x7E4A8634), ("2143", 0, 0xA0F7B07A),
    ("21", 0, 0x72661CF4), ("00000000", 0, 0x2362F9DE), ("000000", 0, 0x85F0B427), ("0000", 0, 0x30F4C306), ("00", 0, 0x514E28B7),
    (b"Hello, world!".hex(), 1234, 0xFAF6CDB3), (b"Hello, world!".hex(), 0x9747B28C, 0x24884CBA),
    (b"The quick brown fox jumps over the lazy dog".hex(), 0x9747B28C, 0x2FA826CD),
]


def main():
    vec = []
    for hx, seed, exp in PUBLISHED:
        got = murmurhash3_32(bytes.fromhex(hx), seed=seed, positive=True)
        assert got == exp, (hx, seed, hex(got), hex(exp))
        vec.append({"hex": hx, "seed": seed, "hash": got, "published": True})
    rng = np.random.default_rng(32)
    for n in list(range(0, 40)) + [63, 64, 65, 127, 255, 1000]:
        b = bytes(rng.integers(0, 256, n, dtype=np.uint8))
        for seed in (0, 0x9747B28C):
            vec.append({"hex": b.hex(), "seed": seed, "hash": murmurhash3_32(b, seed=seed, positive=True)})
    for i in (0, 1, 7, 12345, 15999999):  # the benchmark's key shape: msgpack fixstr "k%015d"
        b = b"\xb0" + (b"k%015d" % i)
        vec.append({"hex": b.hex(), "seed": 0, "hash": murmurhash3_32(b, seed=0, positive=True)})
    ring = [{"name": f"dbeel-{i}", "hash": murmurhash3_32(f"dbeel-{i}".encode(), seed=0, positive=True)} for i in range(8)]
    print(json.dumps({"source": "sklearn.utils.murmurhash3_32 (MurmurHash3_x86_32)", "vectors": vec, "ring_dbeel_8": ring}, indent=0))


if __name__ == "__main__":
    main()
