// stream_pump.h -- the threads between a file (the caller's read / write callbacks) and the pinned rings of the streaming
// host path (dbeel_compact_stream, row N3: the storage edge).  Plain C++: no CUDA in here, so the flow control is
// exercised on a box without a GPU (tests/stream_pump_test.cc); the engine supplies "wait until partition c's D2H has
// landed" as a callable.
//
// The pipeline of dbeel_compact.cu walks the key-range partitions of a compaction in order.  With files on both sides:
//
//   reader threads   pull partition c's slices of every run into ring slot c mod R of the pinned INPUT ring -- as soon
//                    as the partition that used the slot before (c - R) has been consumed (its kernels are done)
//   engine thread    waits for partition c's reads, enqueues its H2D / kernels / D2H (into slot c mod R of the pinned
//                    OUTPUT ring, once partition c - R has left it), publishes what the D2H will deliver
//   writer threads   walk the partitions in order: wait for the D2H, push the slot's bytes through the write callback
//                    in pieces, hand the slot back
//
// so file reads, both PCIe directions and file writes all overlap, and the pinned memory is R slots however large the
// SSTables are.  Any callback error aborts the pump; the first error code is what every wait returns from then on.
#pragma once
#include <stdint.h>

#include <atomic>
#include <condition_variable>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#include "../../../include/dbeel_compact.h"

namespace dbeel {

class StreamPump {
  public:
    struct ReadTask {
        uint32_t part, run, kind;
        uint64_t off, len;
        uint8_t *dst;
    };
    struct OutPart { // where partition c's output sits in the pinned ring and where it goes in the files
        const uint8_t *data = nullptr;
        uint64_t data_len = 0, data_off = 0;
        const uint8_t *index = nullptr;
        uint64_t index_len = 0, index_off = 0;
    };
    static constexpr uint64_t kPiece = 8ull << 20; // bytes per callback call

    // wait_out(c): blocks until partition c's output has arrived in host memory (the engine: cudaEventSynchronize).
    // thread_init(): run once on every pump thread (the engine: cudaSetDevice).
    StreamPump(const dbeel_stream_io *io, uint32_t n_parts, uint32_t ring, int n_threads, std::function<void(uint32_t)> wait_out,
               std::function<void()> thread_init = nullptr)
        : io_(io), np_(n_parts), ring_(ring ? ring : 1), nt_(n_threads > 0 ? n_threads : 1), wait_out_(std::move(wait_out)),
          thread_init_(std::move(thread_init)), r_left_(n_parts, 0), w_left_(n_parts, 0), w_next_(new std::atomic<uint64_t>[n_parts ? n_parts : 1]),
          outs_(n_parts) {
        for (uint32_t c = 0; c < n_parts; c++) w_next_[c].store(0);
    }
    StreamPump(const StreamPump &) = delete;
    StreamPump &operator=(const StreamPump &) = delete;
    ~StreamPump() {
        abort(DBEEL_ERR_INVALID_ARG); // no-op for the error code when the pump finished cleanly
        join();
    }

    // Before start(), in partition order.  A slice longer than kPiece becomes several tasks.
    void add_read(uint32_t part, uint32_t run, uint32_t kind, uint64_t off, uint64_t len, uint8_t *dst) {
        for (uint64_t done = 0; done < len; done += kPiece) {
            const uint64_t n = len - done < kPiece ? len - done : kPiece;
            tasks_.push_back(ReadTask{part, run, kind, off + done, n, dst + done});
            r_left_[part]++;
        }
    }

    void start() {
        started_ = true;
        for (int i = 0; i < nt_; i++) threads_.emplace_back([this] { reader(); });
        for (int i = 0; i < nt_; i++) threads_.emplace_back([this] { writer(); });
    }

    int wait_reads(uint32_t part) {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return failed_ || r_left_[part] == 0; });
        return failed_ ? err_ : 0;
    }

    void release_input(uint32_t part) {
        std::lock_guard<std::mutex> lk(mu_);
        if (part + 1 > consumed_) consumed_ = part + 1;
        cv_.notify_all();
    }

    // The output ring slot of `part` is free once partition part - ring has been written out.
    int wait_out_slot(uint32_t part) {
        if (part < ring_) return 0;
        const uint32_t prev = part - ring_;
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return failed_ || (published_ > prev && w_left_[prev] == 0); });
        return failed_ ? err_ : 0;
    }

    void publish_out(uint32_t part, const OutPart &o) {
        std::lock_guard<std::mutex> lk(mu_);
        outs_[part] = o;
        w_left_[part] = pieces(o.data_len) + pieces(o.index_len);
        published_ = part + 1;
        cv_.notify_all();
    }

    // Everything published has been written (or the pump failed).  Joins the threads.
    int finish() {
        {
            std::unique_lock<std::mutex> lk(mu_);
            cv_.wait(lk, [&] {
                if (failed_) return true;
                if (published_ < np_) return false;
                for (uint32_t c = 0; c < np_; c++)
                    if (w_left_[c]) return false;
                return true;
            });
            done_ = true;
            cv_.notify_all();
        }
        join();
        return failed_ ? err_ : 0;
    }

    void abort(int code) {
        std::lock_guard<std::mutex> lk(mu_);
        if (!failed_ && !done_) {
            failed_ = true;
            err_ = code;
        }
        cv_.notify_all();
    }

  private:
    static uint64_t pieces(uint64_t len) { return (len + kPiece - 1) / kPiece; }

    void join() {
        for (auto &t : threads_)
            if (t.joinable()) t.join();
        threads_.clear();
    }

    void fail_locked(int code) {
        if (!failed_) {
            failed_ = true;
            err_ = code ? code : DBEEL_ERR_INVALID_ARG;
        }
    }

    void reader() {
        if (thread_init_) thread_init_();
        while (true) {
            const size_t k = r_next_.fetch_add(1);
            if (k >= tasks_.size()) return;
            const ReadTask &t = tasks_[k];
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return failed_ || t.part < consumed_ + ring_; });
                if (failed_) return;
            }
            const int rc = io_->read(io_->ctx, t.run, t.kind, t.off, t.len, t.dst);
            std::lock_guard<std::mutex> lk(mu_);
            if (rc) fail_locked(rc);
            r_left_[t.part]--;
            cv_.notify_all();
            if (failed_) return;
        }
    }

    void writer() {
        if (thread_init_) thread_init_();
        for (uint32_t c = 0; c < np_; c++) {
            OutPart o;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return failed_ || published_ > c; });
                if (failed_) return;
                o = outs_[c];
            }
            const uint64_t nd = pieces(o.data_len), total = nd + pieces(o.index_len);
            if (total == 0) continue;
            if (wait_out_) wait_out_(c);
            while (true) {
                const uint64_t k = w_next_[c].fetch_add(1);
                if (k >= total) break;
                const bool is_data = k < nd;
                const uint64_t q = is_data ? k : k - nd, len_all = is_data ? o.data_len : o.index_len;
                const uint64_t off = q * kPiece, n = len_all - off < kPiece ? len_all - off : kPiece;
                const int rc = io_->write(io_->ctx, is_data ? DBEEL_STREAM_DATA : DBEEL_STREAM_INDEX, (is_data ? o.data_off : o.index_off) + off,
                                          (is_data ? o.data : o.index) + off, n);
                std::lock_guard<std::mutex> lk(mu_);
                if (rc) fail_locked(rc);
                w_left_[c]--;
                cv_.notify_all();
                if (failed_) return;
            }
        }
    }

    const dbeel_stream_io *io_;
    const uint32_t np_, ring_;
    const int nt_;
    std::function<void(uint32_t)> wait_out_;
    std::function<void()> thread_init_;
    std::vector<ReadTask> tasks_;
    std::atomic<size_t> r_next_{0};
    std::mutex mu_;
    std::condition_variable cv_;
    // all below under mu_
    std::vector<uint32_t> r_left_;
    std::vector<uint64_t> w_left_;
    std::unique_ptr<std::atomic<uint64_t>[]> w_next_;
    std::vector<OutPart> outs_;
    uint32_t consumed_ = 0;  // partitions [0, consumed_) have released their input slot
    uint32_t published_ = 0; // partitions [0, published_) have their output described
    bool failed_ = false, done_ = false, started_ = false;
    int err_ = 0;
    std::vector<std::thread> threads_;
};

} // namespace dbeel
