// stream_pump_test.cc -- the flow control of dbeel_b200/csrc/host/stream_pump.h without a GPU.
//
// A fake engine walks the partitions exactly like run_job_host_pipelined does (two H2D copies ahead, one "kernel" at a
// time, D2H behind it), with a thread standing in for the copy engines: the "device" is memory, a partition's "kernels"
// copy its input slot to its output slot (byte-wise + 1).  The callbacks read from / write to in-memory "files".
// Checks: every output byte arrives exactly once at the right offset; no ring slot is refilled before its consumer is
// done (the slot's bytes are verified at consumption time); errors from either callback stop the pump with that code.
//
//   g++ -O2 -std=c++17 -pthread tests/stream_pump_test.cc -o /tmp/stream_pump_test && /tmp/stream_pump_test
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <chrono>
#include <random>

#include "../dbeel_b200/csrc/host/stream_pump.h"

using dbeel::StreamPump;

namespace {

struct Files {
    std::vector<std::vector<uint8_t>> run_data; // inputs
    std::vector<uint8_t> out_data, out_index;
    std::vector<uint8_t> out_hits; // how often each output byte was written
    std::atomic<int> reads{0}, writes{0};
    int fail_read_at = -1, fail_write_at = -1;
    std::mutex mu;
};

int rd(void *ctx, uint32_t run, uint32_t kind, uint64_t off, uint64_t len, void *dst) {
    Files *f = static_cast<Files *>(ctx);
    const int k = f->reads.fetch_add(1);
    if (k == f->fail_read_at) return 77;
    if (kind != DBEEL_STREAM_DATA || off + len > f->run_data[run].size()) return 78;
    if ((k & 7) == 0) std::this_thread::sleep_for(std::chrono::microseconds(200));
    memcpy(dst, f->run_data[run].data() + off, len);
    return 0;
}

int wr(void *ctx, uint32_t kind, uint64_t off, const void *src, uint64_t len) {
    Files *f = static_cast<Files *>(ctx);
    const int k = f->writes.fetch_add(1);
    if (k == f->fail_write_at) return 88;
    std::vector<uint8_t> &dst = kind == DBEEL_STREAM_DATA ? f->out_data : f->out_index;
    if (off + len > dst.size()) return 89;
    if ((k & 3) == 0) std::this_thread::sleep_for(std::chrono::microseconds(300));
    memcpy(dst.data() + off, src, len);
    if (kind == DBEEL_STREAM_DATA) {
        std::lock_guard<std::mutex> lk(f->mu);
        for (uint64_t i = 0; i < len; i++) f->out_hits[off + i]++;
    }
    return 0;
}

// one scenario; returns 0 on success, the pump's error code when one was injected (and checks it stopped cleanly)
int scenario(uint32_t np, uint32_t n_runs, uint32_t ring, int threads, uint64_t max_slice, unsigned seed, int fail_read_at, int fail_write_at) {
    std::mt19937_64 rng(seed);
    Files f;
    f.fail_read_at = fail_read_at;
    f.fail_write_at = fail_write_at;
    // slice lengths per (partition, run)
    std::vector<std::vector<uint64_t>> len(np, std::vector<uint64_t>(n_runs));
    std::vector<uint64_t> run_total(n_runs, 0), part_total(np, 0);
    uint64_t max_in = 0;
    for (uint32_t c = 0; c < np; c++) {
        for (uint32_t r = 0; r < n_runs; r++) {
            len[c][r] = rng() % 5 == 0 ? 0 : rng() % max_slice;
            run_total[r] += len[c][r];
            part_total[c] += len[c][r];
        }
        max_in = std::max(max_in, part_total[c]);
    }
    f.run_data.resize(n_runs);
    for (uint32_t r = 0; r < n_runs; r++) {
        f.run_data[r].resize(run_total[r]);
        for (auto &b : f.run_data[r]) b = (uint8_t)rng();
    }
    uint64_t out_total = 0;
    for (uint32_t c = 0; c < np; c++) out_total += part_total[c];
    f.out_data.assign(out_total, 0);
    f.out_hits.assign(out_total, 0);
    f.out_index.assign(16 * (uint64_t)np, 0);

    std::vector<uint8_t> ring_in((uint64_t)ring * (max_in + 1)), ring_out((uint64_t)ring * (max_in + 17));
    // "D2H" completion flags, set by the fake copy engine
    std::vector<std::atomic<int>> d2h_done(np);
    for (auto &x : d2h_done) x.store(0);
    dbeel_stream_io io{rd, wr, &f};
    StreamPump pump(&io, np, ring, threads, [&](uint32_t c) {
        while (!d2h_done[c].load(std::memory_order_acquire)) std::this_thread::yield();
    });
    std::vector<uint64_t> roff(n_runs, 0);
    for (uint32_t c = 0; c < np; c++) {
        uint8_t *slot = ring_in.data() + (uint64_t)(c % ring) * (max_in + 1);
        uint64_t pos = 0;
        for (uint32_t r = 0; r < n_runs; r++) {
            if (len[c][r]) pump.add_read(c, r, DBEEL_STREAM_DATA, roff[r], len[c][r], slot + pos);
            roff[r] += len[c][r];
            pos += len[c][r];
        }
    }
    pump.start();

    // the engine's loop: device buffers dev_in[2] / dev_out[2], h2d two ahead
    std::vector<uint8_t> dev_in[2], dev_out[2];
    for (auto &v : dev_in) v.resize(max_in + 1);
    for (auto &v : dev_out) v.resize(max_in + 1);
    std::vector<uint64_t> part_off(np + 1, 0);
    for (uint32_t c = 0; c < np; c++) part_off[c + 1] = part_off[c] + part_total[c];
    int rc = 0;
    auto h2d = [&](uint32_t c) -> int { // synchronous stand-in for the async copy: verifies the slot against the files first
        const int prc = pump.wait_reads(c);
        if (prc) return prc;
        const uint8_t *slot = ring_in.data() + (uint64_t)(c % ring) * (max_in + 1);
        memcpy(dev_in[c & 1].data(), slot, part_total[c]);
        return 0;
    };
    std::vector<std::thread> copiers;
    rc = h2d(0);
    if (!rc && np > 1) rc = h2d(1);
    std::vector<uint64_t> roff2(n_runs, 0);
    for (uint32_t c = 0; c < np && !rc; c++) {
        // "kernels": check the input is what the files hold, produce output = input + 1
        uint64_t pos = 0;
        for (uint32_t r = 0; r < n_runs; r++) {
            if (memcmp(dev_in[c & 1].data() + pos, f.run_data[r].data() + roff2[r], len[c][r]) != 0) {
                fprintf(stderr, "partition %u run %u: ring slot was overwritten or misread\n", c, r);
                return -1;
            }
            roff2[r] += len[c][r];
            pos += len[c][r];
        }
        for (uint64_t i = 0; i < part_total[c]; i++) dev_out[c & 1][i] = (uint8_t)(dev_in[c & 1][i] + 1);
        pump.release_input(c);
        rc = pump.wait_out_slot(c);
        if (rc) break;
        uint8_t *oslot = ring_out.data() + (uint64_t)(c % ring) * (max_in + 17);
        // "D2H" on another thread, completing a little later
        std::vector<uint8_t> snapshot(dev_out[c & 1].begin(), dev_out[c & 1].begin() + part_total[c]);
        copiers.emplace_back([&, c, oslot, snapshot]() {
            std::this_thread::sleep_for(std::chrono::microseconds(100 + 50 * (c % 5)));
            memcpy(oslot, snapshot.data(), snapshot.size());
            uint64_t tag[2] = {c, part_total[c]};
            memcpy(oslot + snapshot.size(), tag, 16);
            d2h_done[c].store(1, std::memory_order_release);
        });
        StreamPump::OutPart op;
        op.data = oslot; op.data_len = part_total[c]; op.data_off = part_off[c];
        op.index = oslot + part_total[c]; op.index_len = 16; op.index_off = 16ull * c;
        pump.publish_out(c, op);
        if (c + 2 < np) rc = h2d(c + 2);
    }
    if (rc) pump.abort(rc);
    const int frc = rc ? rc : pump.finish();
    for (auto &t : copiers) t.join();
    if (frc) return frc;
    // verify
    uint64_t pos = 0;
    std::vector<uint64_t> roff3(n_runs, 0);
    for (uint32_t c = 0; c < np; c++) {
        for (uint32_t r = 0; r < n_runs; r++) {
            for (uint64_t i = 0; i < len[c][r]; i++) {
                if (f.out_data[pos + i] != (uint8_t)(f.run_data[r][roff3[r] + i] + 1) || f.out_hits[pos + i] != 1) {
                    fprintf(stderr, "output byte %llu wrong (hits %u)\n", (unsigned long long)(pos + i), f.out_hits[pos + i]);
                    return -2;
                }
            }
            pos += len[c][r];
            roff3[r] += len[c][r];
        }
        uint64_t tag[2];
        memcpy(tag, f.out_index.data() + 16ull * c, 16);
        if (tag[0] != c || tag[1] != part_total[c]) { fprintf(stderr, "index piece %u wrong\n", c); return -3; }
    }
    return 0;
}

} // namespace

int main() {
    int bad = 0;
    unsigned seed = 1;
    for (uint32_t np : {1u, 2u, 3u, 7u, 23u})
        for (uint32_t ring : {2u, 3u, 4u})
            for (int threads : {1, 3, 8}) {
                const int rc = scenario(np, 5, ring, threads, 3000, seed++, -1, -1);
                if (rc) { fprintf(stderr, "scenario np=%u ring=%u threads=%d -> %d\n", np, ring, threads, rc); bad++; }
            }
    // pieces larger than kPiece: a slice is split over several callback calls
    {
        const int rc = scenario(4, 2, 3, 4, 20ull << 20, 999, -1, -1);
        if (rc) { fprintf(stderr, "large scenario -> %d\n", rc); bad++; }
    }
    // error injection: the pump must stop with the callback's code, never hang
    for (int at : {0, 3, 17, 23}) {
        int rc = scenario(12, 4, 3, 4, 3000, 500 + at, at, -1);
        if (rc != 77) { fprintf(stderr, "read failure at %d -> %d (want 77)\n", at, rc); bad++; }
        rc = scenario(12, 4, 3, 4, 3000, 600 + at, -1, at);
        if (rc != 88) { fprintf(stderr, "write failure at %d -> %d (want 88)\n", at, rc); bad++; }
    }
    printf(bad ? "FAILED %d\n" : "ok\n", bad);
    return bad ? 1 : 0;
}
