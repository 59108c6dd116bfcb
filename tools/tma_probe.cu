// tma_probe.cu -- microbenchmark behind the payload-gather design.
//
// Question: how fast can B200 move VARIABLE-LENGTH, BYTE-ALIGNED records (SSTable entries: 32 B .. KBs, arbitrary source
// and destination byte offsets) with the TMA doing the byte realignment?  A `cp.async.bulk.tensor` over a u8 tensor takes
// its coordinates in ELEMENTS (= bytes), so a box may start at any byte; the tensor here is the flat address range seen as
// rows of 256 bytes that overlap (row stride 256, row width 511), so the box {256, k} at (c0, r) is the contiguous byte
// range [256 r + c0, 256 (r + k) + c0).  Load: global (any byte offset) -> dense smem slot; store: smem slot -> global
// (any byte offset).  No per-byte instructions at all.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/bin/tma_probe tools/tma_probe.cu
//   tools/bin/tma_probe [entry_len=305] [n_entries=4000000] [warps=4] [stages=4] [ctas_per_sm=1]
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#define CK(x)                                                                                   \
    do {                                                                                        \
        cudaError_t e_ = (x);                                                                   \
        if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } \
    } while (0)

// descriptor table: [0..15] box {16 (m+1), 1}; [16 + j] box {256, 2^j} for j = 0..5 (256 B .. 8 KB)
constexpr int kSmall = 16, kBig = 6, kMaps = kSmall + kBig;
struct Maps {
    CUtensorMap m[kMaps];
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void tma_load_2d(uint32_t smem_dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(smem_dst),
                 "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap *map, int c0, int c1, uint32_t smem_src) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];" ::"l"(map), "r"(c0), "r"(c1), "r"(smem_src) : "memory");
}

// Pieces of an entry of n bytes (n >= 32): power-of-two multiples of 256 from the front, then the remainder (< 256) as ONE box
// of ceil16(rem) bytes that ENDS at the entry's last byte (it overlaps up to 15 bytes of what precedes it: same bytes twice).
// fn(map index, byte offset inside the entry, smem offset inside the slot, bytes)
template <class F>
__device__ __forceinline__ void for_pieces(uint32_t n, F fn) {
    uint32_t done = 0, big = n >> 8;
    while (big) {
        uint32_t j = 31 - __clz(big);
        if (j > kBig - 1) j = kBig - 1;
        fn(kSmall + j, done, done, 256u << j);
        done += 256u << j;
        big -= 1u << j;
    }
    const uint32_t rem = n - done;
    if (rem) {
        const uint32_t sz = (rem + 15) & ~15u;
        // done == 0: the entry is shorter than 256: front box of floor16 bytes + a 16-byte tail box if needed
        if (done == 0) {
            const uint32_t fl = n & ~15u;
            fn(fl / 16 - 1, 0, 0, fl);
            if (n & 15) fn(0, n - 16, 256, 16);
        } else {
            fn(sz / 16 - 1, n - sz, done, sz);
        }
    }
}

__global__ void __launch_bounds__(256) k_tma_copy(const __grid_constant__ Maps maps, uint64_t base, const uint64_t *src_off,
                                                 const uint64_t *dst_off, const uint32_t *len, uint32_t n, uint32_t slot, int stages) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bars[8 * 8];
    __shared__ uint64_t s_dst[8][8][32];
    __shared__ uint32_t s_len[8][8][32];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    uint64_t *bar = bars + warp * 8;
    if (lane == 0)
        for (int s = 0; s < stages; s++) mbar_init(&bar[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    const uint32_t ring = smem_u32(smem) + warp * (uint32_t)stages * 32u * slot;
    const uint32_t n_batches = (n + 31) / 32;
    const uint32_t gw = blockIdx.x * nw + warp, tw = gridDim.x * nw;
    const uint32_t mine = gw < n_batches ? (n_batches - gw + tw - 1) / tw : 0;
    const uint32_t lag = (uint32_t)stages - 2; // the batch completed at iteration it is batch it - lag
    for (uint32_t it = 0; it < mine + lag; it++) {
        if (it < mine) {
            const uint32_t st = it % (uint32_t)stages;
            // stage st was last used by the batch of iteration it - stages; its stores were committed at iteration
            // it - stages + lag = it - 2: only the newest group (it - 1) may still be reading shared memory
            asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            __syncwarp();
            const uint32_t e = (gw + it * tw) * 32 + lane;
            uint64_t so = 0;
            uint32_t ln = 0;
            if (e < n) { so = src_off[e]; ln = len[e]; s_dst[warp][st][lane] = dst_off[e]; }
            s_len[warp][st][lane] = ln;
            uint32_t bytes = 0;
            if (ln) for_pieces(ln, [&](uint32_t, uint32_t, uint32_t, uint32_t sz) { bytes += sz; });
            for (int o = 16; o; o >>= 1) bytes += __shfl_xor_sync(0xFFFFFFFFu, bytes, o);
            if (lane == 0) mbar_expect_tx(&bar[st], bytes);
            __syncwarp();
            if (ln) {
                const uint32_t sl = ring + (st * 32u + lane) * slot;
                for_pieces(ln, [&](uint32_t mi, uint32_t eo, uint32_t so_, uint32_t) {
                    const uint64_t a = so + eo - base;
                    tma_load_2d(sl + so_, &maps.m[mi], (int)(a & 255), (int)(a >> 8), &bar[st]);
                });
            }
        }
        if (it >= lag) { // complete batch it - lag
            const uint32_t ci = it - lag;
            const uint32_t cs = ci % (uint32_t)stages;
            while (!mbar_try_wait(&bar[cs], (ci / (uint32_t)stages) & 1)) {}
            const uint32_t ln = s_len[warp][cs][lane];
            if (ln) {
                const uint32_t sl = ring + (cs * 32u + lane) * slot;
                const uint64_t d = s_dst[warp][cs][lane];
                for_pieces(ln, [&](uint32_t mi, uint32_t eo, uint32_t so_, uint32_t) {
                    const uint64_t a = d + eo - base;
                    tma_store_2d(&maps.m[mi], (int)(a & 255), (int)(a >> 8), sl + so_);
                });
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// reference: one warp per entry, byte copy
__global__ void k_ref_copy(const uint8_t *src, uint8_t *dst, const uint64_t *src_off, const uint64_t *dst_off, const uint32_t *len, uint32_t n) {
    const uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= n) return;
    const uint8_t *s = src + src_off[w];
    uint8_t *d = dst + dst_off[w];
    for (uint32_t i = lane; i < len[w]; i += 32) d[i] = s[i];
}

__global__ void k_fill(uint32_t *p, uint64_t nwords) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < nwords; i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t x = i * 0x9E3779B97F4A7C15ull;
        x ^= x >> 29;
        p[i] = (uint32_t)(x * 0xBF58476D1CE4E5B9ull >> 32);
    }
}

__global__ void k_diff(const uint4 *a, const uint4 *b, uint64_t nvec, unsigned long long *bad) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < nvec; i += (uint64_t)gridDim.x * blockDim.x) {
        uint4 x = a[i], y = b[i];
        if (x.x != y.x || x.y != y.y || x.z != y.z || x.w != y.w) atomicAdd(bad, 1ull);
    }
}


typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                             const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// ------------------------------------------------------------------------------------ self tests (one process each)
__device__ __forceinline__ void tma_load_1d_t(uint32_t smem_dst, const CUtensorMap *map, int c0, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.1d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2}], [%3];" ::"r"(smem_dst),
                 "l"(map), "r"(c0), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_store_1d_t(const CUtensorMap *map, int c0, uint32_t smem_src) {
    asm volatile("cp.async.bulk.tensor.1d.global.shared::cta.tile.bulk_group [%0, {%1}], [%2];" ::"l"(map), "r"(c0), "r"(smem_src) : "memory");
}

// which: 0 = the single __grid_constant__ descriptor, 1 = arr.m[idx] (param array, dynamic index), 2 = descriptor in global memory
// op: 0 = load box -> smem -> plain stores to out; 1 = plain loads from `in` -> smem -> TMA store of the box
__global__ void k_selftest(const __grid_constant__ CUtensorMap one, const __grid_constant__ Maps arr, const CUtensorMap *gmap, int which, int idx,
                           int rank, int c0, int c1, uint32_t bytes, int op, const uint8_t *in, uint8_t *out) {
    __shared__ __align__(1024) uint8_t buf[8192];
    __shared__ __align__(8) uint64_t bar;
    const CUtensorMap *m = which == 0 ? &one : (which == 1 ? &arr.m[idx] : gmap);
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (op == 0) {
        if (threadIdx.x == 0) {
            mbar_expect_tx(&bar, bytes);
            if (rank == 1) tma_load_1d_t(smem_u32(buf), m, c0, &bar);
            else tma_load_2d(smem_u32(buf), m, c0, c1, &bar);
        }
        while (!mbar_try_wait(&bar, 0)) {}
        for (uint32_t i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = buf[i];
    } else {
        for (uint32_t i = threadIdx.x; i < bytes; i += blockDim.x) buf[i] = in[i];
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (threadIdx.x == 0) {
            if (rank == 1) tma_store_1d_t(m, c0, smem_u32(buf));
            else tma_store_2d(m, c0, c1, smem_u32(buf));
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        }
    }
}

static EncodeFn get_encode() {
    EncodeFn encode = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", reinterpret_cast<void **>(&encode), cudaEnableDefault, &qres));
    return encode;
}

// tools/bin/tma_probe selftest <layout 0 flat1d | 1 rows256 | 2 overlapped rows> <box0> <box1> <byte offset> <which> <op>
static int selftest(int argc, char **argv) {
    const int layout = atoi(argv[2]), box0 = atoi(argv[3]), box1 = atoi(argv[4]);
    const uint64_t byte_off = strtoull(argv[5], nullptr, 10);
    const int which = atoi(argv[6]), op = atoi(argv[7]);
    CK(cudaSetDevice(0));
    const uint64_t W = 1ull << 24;
    uint8_t *src, *out;
    CK(cudaMalloc(&src, W + 8192));
    CK(cudaMalloc(&out, W + 8192));
    std::vector<uint8_t> h(W);
    for (uint64_t i = 0; i < W; i++) h[i] = (uint8_t)((i * 2654435761u) >> 13);
    CK(cudaMemcpy(src, h.data(), W, cudaMemcpyHostToDevice));
    CK(cudaMemset(out, 0xEE, W));
    EncodeFn encode = get_encode();
    Maps arr;
    memset(&arr, 0, sizeof arr);
    CUtensorMap one;
    const uint32_t rank = layout == 0 ? 1 : 2;
    cuuint64_t gdim[2] = {layout == 0 ? W : (layout == 1 ? 256ull : 511ull), W / 256 - 2};
    cuuint64_t gstr[1] = {256};
    cuuint32_t box[2] = {(cuuint32_t)box0, (cuuint32_t)box1};
    cuuint32_t estr[2] = {1, 1};
    void *base = op == 0 ? src : out;
    CUresult r = encode(&one, CU_TENSOR_MAP_DATA_TYPE_UINT8, rank, base, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("selftest: encode failed (%d)\n", (int)r); return 3; }
    for (int i = 0; i < kMaps; i++) arr.m[i] = one;
    CUtensorMap *gmap;
    CK(cudaMalloc(&gmap, sizeof(CUtensorMap)));
    CK(cudaMemcpy(gmap, &one, sizeof one, cudaMemcpyHostToDevice));
    const uint32_t bytes = (uint32_t)box0 * (uint32_t)box1;
    const int c0 = layout == 0 ? (int)byte_off : (int)(byte_off & 255), c1 = (int)(byte_off >> 8);
    k_selftest<<<1, 128>>>(one, arr, gmap, which, 7, (int)rank, c0, c1, bytes, op, src + 4096, out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("selftest: kernel failed: %s\n", cudaGetErrorString(e)); return 4; }
    std::vector<uint8_t> got(W);
    CK(cudaMemcpy(got.data(), out, W, cudaMemcpyDeviceToHost));
    uint64_t bad = 0;
    if (op == 0) {
        for (uint32_t i = 0; i < bytes; i++) bad += got[i] != h[byte_off + i];
    } else {
        for (uint64_t i = 0; i < W; i++) {
            const uint8_t exp = (i >= byte_off && i < byte_off + bytes) ? h[4096 + (i - byte_off)] : 0xEE;
            bad += got[i] != exp;
        }
    }
    printf("selftest layout %d box {%d,%d} byte offset %llu which %d op %s: %s (%llu bad bytes)\n", layout, box0, box1,
           (unsigned long long)byte_off, which, op ? "store" : "load", bad ? "FAIL" : "PASS", (unsigned long long)bad);
    return bad ? 2 : 0;
}

int main(int argc, char **argv) {
    if (argc > 7 && !strcmp(argv[1], "selftest")) return selftest(argc, argv);
    const uint32_t L = argc > 1 ? atoi(argv[1]) : 305;
    const uint32_t n = argc > 2 ? atoi(argv[2]) : 4000000;
    const int warps = argc > 3 ? atoi(argv[3]) : 4;
    int stages = argc > 4 ? atoi(argv[4]) : 4;
    const int cps = argc > 5 ? atoi(argv[5]) : 1;
    const int jitter = argc > 6 ? atoi(argv[6]) : 1; // entry lengths L +- jitter*(0..15)
    if (stages < 3) stages = 3;
    if (stages > 8) stages = 8;
    CK(cudaSetDevice(0));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));

    // entries: source = 8 interleaved "runs" with 20 % holes (dropped entries), destination dense
    std::vector<uint64_t> so(n), dof(n);
    std::vector<uint32_t> ln(n);
    uint64_t s = 3, d = 5; // deliberately odd starts
    uint64_t rng = 88172645463325252ull;
    for (uint32_t i = 0; i < n; i++) {
        rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17;
        uint32_t l = L + (jitter ? (uint32_t)(rng & 15) * (uint32_t)jitter : 0);
        if (l < 32) l = 32;
        ln[i] = l;
        so[i] = s;
        dof[i] = d;
        s += l;
        if ((rng >> 20) % 5 == 0) s += L; // a dropped entry in between
        d += l;
    }
    const uint64_t src_bytes = (s + 4096 + 15) & ~15ull, dst_bytes = (d + 4096 + 15) & ~15ull;
    uint8_t *buf;
    // one allocation: [src | dst_tma | dst_ref] so a single descriptor window covers everything
    CK(cudaMalloc(&buf, src_bytes + 2 * dst_bytes));
    uint8_t *src = buf, *dst = buf + src_bytes, *ref = dst + dst_bytes;
    k_fill<<<1024, 256>>>(reinterpret_cast<uint32_t *>(src), src_bytes / 4);
    CK(cudaMemset(dst, 0, 2 * dst_bytes));
    uint64_t *d_so, *d_do, *d_do_ref;
    uint32_t *d_ln;
    CK(cudaMalloc(&d_so, 8ull * n)); CK(cudaMalloc(&d_do, 8ull * n)); CK(cudaMalloc(&d_do_ref, 8ull * n)); CK(cudaMalloc(&d_ln, 4ull * n));
    const uint64_t base = (uint64_t)(uintptr_t)buf & ~255ull;
    std::vector<uint64_t> a(n);
    for (uint32_t i = 0; i < n; i++) a[i] = (uint64_t)(uintptr_t)src + so[i];
    CK(cudaMemcpy(d_so, a.data(), 8ull * n, cudaMemcpyHostToDevice));
    for (uint32_t i = 0; i < n; i++) a[i] = (uint64_t)(uintptr_t)dst + dof[i];
    CK(cudaMemcpy(d_do, a.data(), 8ull * n, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_do_ref, dof.data(), 8ull * n, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_ln, ln.data(), 4ull * n, cudaMemcpyHostToDevice));
    uint64_t *d_so_rel;
    CK(cudaMalloc(&d_so_rel, 8ull * n));
    CK(cudaMemcpy(d_so_rel, so.data(), 8ull * n, cudaMemcpyHostToDevice));

    // descriptors
    EncodeFn encode = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", reinterpret_cast<void **>(&encode), cudaEnableDefault, &qres));
    if (!encode) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
    Maps maps;
    const uint64_t window = src_bytes + 2 * dst_bytes + 512;
    cuuint64_t gdim[2] = {511, (window + 255) / 256};
    cuuint64_t gstr[1] = {256};
    cuuint32_t estr[2] = {1, 1};
    for (int i = 0; i < kMaps; i++) {
        cuuint32_t box[2] = {i < kSmall ? 16u * (i + 1) : 256u, i < kSmall ? 1u : 1u << (i - kSmall)};
        CUresult r = encode(&maps.m[i], CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, reinterpret_cast<void *>(base), gdim, gstr, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode %d failed: %d\n", i, (int)r); return 1; }
    }

    // slot per lane: big pieces at their entry offsets, remainder piece behind them, all 128-byte aligned
    uint32_t maxl = L + 15 * jitter;
    uint32_t slot = maxl < 256 ? 384 : (((maxl >> 8) << 8) + 256);
    slot = (slot + 127) & ~127u;
    while ((size_t)warps * stages * 32 * slot > 200 * 1024 / (size_t)cps && stages > 3) stages--;
    const size_t smem = (size_t)warps * stages * 32 * slot;
    if (smem > 227 * 1024) { printf("ring of %zu B does not fit\n", smem); return 1; }
    CK(cudaFuncSetAttribute(k_tma_copy, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = prop.multiProcessorCount * cps;
    printf("entries %u x ~%u B (%.1f MB out), warps/CTA %d, stages %d, CTAs/SM %d, slot %u B, smem/CTA %zu B\n", n, L, d / 1e6, warps, stages,
           cps, slot, smem);

    k_ref_copy<<<(n + 7) / 8, 256>>>(src, ref, d_so_rel, d_do_ref, d_ln, n);
    CK(cudaDeviceSynchronize());

    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e9f;
    for (int rep = 0; rep < 6; rep++) {
        CK(cudaEventRecord(e0));
        k_tma_copy<<<grid, warps * 32, smem>>>(maps, base, d_so, d_do, d_ln, n, slot, stages);
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep && ms < best) best = ms;
    }
    unsigned long long *d_bad, h_bad = 0;
    CK(cudaMalloc(&d_bad, 8));
    CK(cudaMemset(d_bad, 0, 8));
    k_diff<<<1024, 256>>>(reinterpret_cast<const uint4 *>(dst), reinterpret_cast<const uint4 *>(ref), dst_bytes / 16, d_bad);
    CK(cudaMemcpy(&h_bad, d_bad, 8, cudaMemcpyDeviceToHost));
    // pieces per entry
    double pieces = 0;
    for (uint32_t i = 0; i < n; i++) {
        uint32_t l = ln[i], big = l >> 8, c = 0;
        while (big) { uint32_t j = 31 - __builtin_clz(big); if (j > kBig - 1) j = kBig - 1; big -= 1u << j; c++; }
        if (l & 255) c += (l < 256 && (l & 15)) ? 2 : 1;
        pieces += c;
    }
    printf("tma copy: %.3f ms  -> %.1f GB/s (read+write)  %.1f M entries/s  %.2f pieces/entry  %.2f G TMA ops/s (%.1f cyc/op/SM @1.9GHz)  mismatching vectors: %llu\n",
           best, 2.0 * d / best / 1e6, n / best / 1e3, pieces / n, 2 * pieces / best / 1e6, 1.9e6 * best * prop.multiProcessorCount / (2 * pieces), h_bad);

    // memcpy reference
    CK(cudaEventRecord(e0));
    for (int i = 0; i < 5; i++) CK(cudaMemcpyAsync(ref, dst, d, cudaMemcpyDeviceToDevice));
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("cudaMemcpy D2D of the same bytes: %.3f ms -> %.1f GB/s (read+write)\n", ms / 5, 2.0 * d / (ms / 5) / 1e6);
    return h_bad ? 2 : 0;
}
