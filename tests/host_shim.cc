// Host build of dbeel_b200/csrc/device_fns.cuh (the same text nvcc compiles as device code),
// exported with a C ABI so tests/test_device_fns_host.py can drive it without a GPU.
#include <stdint.h>
#include <string.h>

#include "../dbeel_b200/csrc/device_fns.cuh"

using namespace dbeel;

static uint64_t ld_le(const uint8_t *p, uint64_t avail) { // garbage-tolerant loader: pads with 0xAA
    uint8_t b[8];
    for (int i = 0; i < 8; i++) b[i] = (uint64_t)i < avail ? p[i] : 0xAA;
    uint64_t v;
    memcpy(&v, b, 8);
    return v;
}

extern "C" {

void shim_make_rec(const uint8_t *key, uint64_t klen, uint32_t L, uint32_t gid, uint32_t out[4]) {
    uint64_t rem = klen - L;
    uint64_t w0 = ld_le(key + L, rem), w1 = ld_le(key + L + 8, rem > 8 ? rem - 8 : 0);
    Rec r = make_rec(w0, w1, rem, gid);
    out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
}

int shim_rec_cmp(const uint32_t a[4], const uint32_t b[4], int *undecided) {
    Rec ra = {a[0], a[1], a[2], a[3]}, rb = {b[0], b[1], b[2], b[3]};
    return rec_cmp_window(ra, rb, undecided);
}

void shim_sip_pair(const uint64_t k[4], const uint8_t *key, uint64_t klen, uint64_t out[2]) {
    sip13_pair_vec_u8(k, klen, [&](uint64_t j) { return ld_le(key + 8 * j, klen - 8 * j); }, &out[0], &out[1]);
}

uint64_t shim_fastmod(uint64_t h, uint64_t d) {
    uint64_t magic = (uint64_t)((((unsigned __int128)1) << 64) / d);
    return fastmod(h, d, magic);
}

uint64_t shim_bloom_hash_i(uint64_t h0, uint64_t h1, uint32_t i) { return bloom_hash_i(h0, h1, i); }

uint32_t shim_bloom_probe_all(uint64_t h0, uint64_t h1, uint32_t k_num, uint64_t bits, uint64_t *out) {
    uint32_t n = 0;
    const uint64_t magic = (uint64_t)((((unsigned __int128)1) << 64) / bits);
    bloom_probe_all(h0, h1, k_num, bits, magic, [&](uint64_t bit) { out[n++] = bit; });
    return n;
}

void shim_realign16(const uint8_t src32[32], uint32_t sh, uint8_t out16[16]) {
    uint32_t A[4], B[4], O[4];
    memcpy(A, src32, 16);
    memcpy(B, src32 + 16, 16);
    realign16(A, B, sh, O);
    memcpy(out16, O, 16);
}

void shim_window32(const uint8_t src64[64], uint32_t s0, uint8_t out32[32]) {
    uint32_t w[16], o[8];
    memcpy(w, src64, 64);
    window32(w, s0, o);
    memcpy(out32, o, 32);
}

void shim_blend32(const uint8_t t32[32], const uint8_t h32[32], uint32_t t, uint8_t out32[32]) {
    uint32_t T[8], H[8], O[8];
    memcpy(T, t32, 32);
    memcpy(H, h32, 32);
    blend32(T, H, t, O);
    memcpy(out32, O, 32);
}

int shim_ts_decodes(const uint8_t ts[16]) {
    uint64_t lo, hi;
    memcpy(&lo, ts, 8); memcpy(&hi, ts + 8, 8);
    return ts_decodes(lo, hi) ? 1 : 0;
}

int shim_ts_greater(const uint8_t a[16], const uint8_t b[16]) {
    uint64_t al, ah, bl, bh;
    memcpy(&al, a, 8); memcpy(&ah, a + 8, 8); memcpy(&bl, b, 8); memcpy(&bh, b + 8, 8);
    return ts_greater(al, ah, bl, bh) ? 1 : 0;
}

uint32_t shim_murmur3_32(const uint8_t *key, uint64_t len, uint32_t seed) {
    return murmur3_32(len, seed, [&](uint64_t q) { return ld_le(key + 8 * q, len - 8 * q); });
}

uint32_t shim_ring_owner(const uint32_t *ring, uint32_t n, uint32_t h) {
    return ring_owner(n, h, [&](uint32_t s) { return ring[s]; });
}
}
