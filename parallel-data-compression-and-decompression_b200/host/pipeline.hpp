// pipeline.hpp — the worker pool both host pipelines share (internal to the host library).
//
// The reference overlaps nothing on the compress side but producer/consumer hand-off (compression.cpp:160-194) and runs one
// thread per archive on the decompress side (decompression.cpp:174). Here a rank keeps W workers (ZWZ_WORKERS, default 3),
// each with its own zwz_ctx — its own CUDA stream, device arenas and page-locked staging — and every worker takes whole
// batches through read -> GPU -> write on its own. Kernels of different workers overlap on the device (the MD5 of a few long
// files is a serial chain that occupies a handful of SMs for ~100 ms; the other workers' deflate/inflate kernels fill the
// rest of the GPU meanwhile), copies overlap kernels, and file I/O overlaps both. Results are committed in batch order, so
// the archive bytes and the console output do not depend on W or on timing.
#pragma once
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdio>
#include <unistd.h>
#include <exception>
#include <functional>
#include <mutex>
#include <stdexcept>
#include <thread>
#include <vector>

#include "zwz_host.hpp"

namespace zwzhost {

zwz_ctx *ctx_for(int device);
zwz_ctx *worker_ctx(int device, int worker); // worker 0 shares ctx_for(device); the others get their own, created on first use
int worker_count();                          // ZWZ_WORKERS (default: host cores / ranks on this box / 2), clamped to [1, 12]
int io_threads();                            // ZWZ_IO_THREADS: reader/writer threads a worker uses inside one batch (default 4)
std::string md5_of_file_ctx(zwz_ctx *ctx, const std::string &file_path);

// batch b may commit only after batches 0..b-1 have
class OrderedCommit {
  public:
    void wait_turn(size_t index) {
        std::unique_lock<std::mutex> lock(mu_);
        cv_.wait(lock, [&] { return next_ == index || aborted_; });
        if (aborted_) throw std::runtime_error("zwz: pipeline aborted");
    }
    void done(size_t index) {
        {
            std::lock_guard<std::mutex> lock(mu_);
            if (next_ == index) next_ = index + 1;
        }
        cv_.notify_all();
    }
    void abort() {
        {
            std::lock_guard<std::mutex> lock(mu_);
            aborted_ = true;
        }
        cv_.notify_all();
    }

  private:
    std::mutex mu_;
    std::condition_variable cv_;
    size_t next_ = 0;
    bool aborted_ = false;
};

// runs body(worker_id) on n threads (n == 1: on the caller's thread); the first exception is rethrown on the caller
inline void run_workers(int n, OrderedCommit &order, const std::function<void(int)> &body) {
    if (n <= 1) {
        body(0);
        return;
    }
    std::exception_ptr first;
    std::mutex mu;
    std::vector<std::thread> threads;
    for (int w = 0; w < n; ++w)
        threads.emplace_back([&, w] {
            try {
                body(w);
            } catch (...) {
                {
                    std::lock_guard<std::mutex> lock(mu);
                    if (!first) first = std::current_exception();
                }
                order.abort();
            }
        });
    for (auto &t : threads) t.join();
    if (first) std::rethrow_exception(first);
}

// fn(i) for i in [0, n) on up to `threads` threads (the caller's included); fn must not throw. Used for the file I/O inside
// one batch: 370 000 small files are bound by open/read/close (resp. open/write/close) latency, not by bandwidth.
template <class F> inline void parallel_for(size_t n, int threads, F &&fn) {
    if (n == 0) return;
    const size_t nt = std::min<size_t>((size_t) std::max(1, threads), n);
    if (nt <= 1) {
        for (size_t i = 0; i < n; ++i) fn(i);
        return;
    }
    std::atomic<size_t> next{0};
    const size_t grain = std::max<size_t>(1, n / (nt * 8));
    auto body = [&] {
        for (;;) {
            size_t a = next.fetch_add(grain);
            if (a >= n) return;
            size_t b = std::min(n, a + grain);
            for (size_t i = a; i < b; ++i) fn(i);
        }
    };
    std::vector<std::thread> ts;
    for (size_t t = 1; t < nt; ++t) ts.emplace_back(body);
    body();
    for (auto &t : ts) t.join();
}

// ---- the segment ledger: how the ranks of one run tell each other a few numbers about the segments of a file that is cut
// over several GPUs (sizes, offsets). Tiny files under <output dir>/.zwz_segments/, published with write + rename (a reader
// sees all of an entry or none) and polled by the ranks that need them — the same shared-directory channel the ranks use for
// the file record (main.cpp); the reference's ranks exchange nothing but that path either (main.cpp:27-35). The `.zwz`
// format itself has no index to exchange: what the ranks need from each other is where a segment's records go.
inline std::string ledger_dir(const std::string &output_dir) { return output_dir + "/.zwz_segments"; }
inline void ledger_publish(const std::string &output_dir, const std::string &key, int seg, const char *what, uint64_t a, uint64_t b = 0) {
    const std::string final_name = ledger_dir(output_dir) + "/" + key + "." + std::to_string(seg) + what;
    const std::string tmp = final_name + ".tmp" + std::to_string((long) getpid());
    std::FILE *f = std::fopen(tmp.c_str(), "wb");
    if (!f) throw std::runtime_error("zwz: cannot write the segment ledger");
    uint64_t v[2] = {a, b};
    std::fwrite(v, 8, 2, f);
    std::fclose(f);
    if (std::rename(tmp.c_str(), final_name.c_str()) != 0) throw std::runtime_error("zwz: cannot publish to the segment ledger");
}
inline uint64_t ledger_wait(const std::string &output_dir, const std::string &key, int seg, const char *what, uint64_t *b = nullptr) {
    const std::string name = ledger_dir(output_dir) + "/" + key + "." + std::to_string(seg) + what;
    const double t0 = now_seconds();
    for (;;) {
        std::FILE *f = std::fopen(name.c_str(), "rb");
        if (f) {
            uint64_t v[2] = {0, 0};
            size_t k = std::fread(v, 8, 2, f);
            std::fclose(f);
            if (k == 2) {
                if (b) *b = v[1];
                return v[0];
            }
        }
        if (now_seconds() - t0 > 900.0)
            throw std::runtime_error("zwz: timed out waiting for segment " + std::to_string(seg) + " of " + key + " from another rank");
        usleep(500);
    }
}

// grow-only page-locked host buffer (copies from/to it run at full PCIe speed and asynchronously)
struct PinnedBuf {
    zwz_ctx *ctx = nullptr;
    uint8_t *p = nullptr;
    size_t cap = 0;
    explicit PinnedBuf(zwz_ctx *c) : ctx(c) {}
    ~PinnedBuf() {
        if (p) zwz_free_pinned(ctx, p);
    }
    PinnedBuf(const PinnedBuf &) = delete;
    PinnedBuf &operator=(const PinnedBuf &) = delete;
    void reserve(size_t n) {
        if (n <= cap) return;
        if (p) zwz_free_pinned(ctx, p);
        p = nullptr;
        cap = n + n / 8 + 4096;
        if (zwz_malloc_pinned(ctx, cap, (void **) &p) != ZWZ_OK) throw std::runtime_error("zwz: pinned allocation failed");
    }
    uint8_t *data() { return p; }
};

} // namespace zwzhost
