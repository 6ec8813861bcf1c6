// verify.cpp — md5_of_file (verification.cpp:6-30) on the GPU, and the process-wide zwz_ctx.
#include "pipeline.hpp"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <map>
#include <mutex>
#include <stdexcept>
#include <thread>

namespace zwzhost {

// one context per device per process, created on first use. No GPU => hard error: this build has no CPU path.
zwz_ctx *ctx_for(int device) {
    static std::mutex mu;
    static std::map<int, zwz_ctx *> all;
    std::lock_guard<std::mutex> lock(mu);
    auto it = all.find(device);
    if (it != all.end()) return it->second;
    zwz_ctx *c = nullptr;
    double t0 = now_seconds();
    // the rank's device index was derived without asking CUDA (distribute.cpp); fold it into what the runtime really offers
    const int nd = zwz_device_count();
    int rc = zwz_init(nd > 0 ? device % nd : device, &c);
    stats().t_init += now_seconds() - t0;
    if (rc != ZWZ_OK || !c) {
        std::cerr << "zwz: cannot initialise CUDA device " << device << " (error " << rc << "); there is no CPU fallback" << std::endl;
        throw std::runtime_error("zwz_init failed");
    }
    if (timing_level() >= 2) zwz_profile_enable(c, 1);
    zwz_ctx_tune(c, ZWZ_TUNE_DEFLATE_SUBBATCH_BYTES, (uint64_t) 32 << 20); // batches of <= 128 MB: small scratch, quick start-up
    all[device] = c;
    return c;
}

void warm_device() {
    try {
        ctx_for(config().device);
    } catch (...) {
    }
}

// the extra contexts of the worker pool (pipeline.hpp): cheap once the device's primary context exists
zwz_ctx *worker_ctx(int device, int worker) {
    if (worker <= 0) return ctx_for(device);
    ctx_for(device); // the first initialisation of the device is the slow one: do it once, outside the lock below
    static std::mutex mu;
    static std::map<std::pair<int, int>, zwz_ctx *> all;
    std::lock_guard<std::mutex> lock(mu);
    auto key = std::make_pair(device, worker);
    auto it = all.find(key);
    if (it != all.end()) return it->second;
    zwz_ctx *c = nullptr;
    const int nd = zwz_device_count();
    if (zwz_init(nd > 0 ? device % nd : device, &c) != ZWZ_OK || !c) throw std::runtime_error("zwz_init failed for a worker context");
    if (timing_level() >= 2) zwz_profile_enable(c, 1);
    zwz_ctx_tune(c, ZWZ_TUNE_DEFLATE_SUBBATCH_BYTES, (uint64_t) 32 << 20);
    all[key] = c;
    return c;
}

int worker_count() {
    const char *e = std::getenv("ZWZ_WORKERS");
    int w;
    if (e && *e) {
        w = std::atoi(e);
    } else {
        // every worker is a host thread that reads, serialises and writes besides driving its CUDA stream: half the cores this
        // rank may use (the ranks of one box share them), at least 3
        int cores = (int) std::thread::hardware_concurrency();
        int local_ranks = std::max(1, std::min(config().world_size, std::max(1, visible_gpu_count())));
        // measured on a 16-core box (profiles/round2/cli): 3 and 4 workers tie, 8 lose a second or more — every worker brings
        // its own page-locked buffers and device arenas, and allocating those is the larger part of a 2 GB job
        w = std::max(2, std::min(4, cores / local_ranks / 4));
    }
    return w < 1 ? 1 : (w > 12 ? 12 : w);
}

int io_threads() {
    const char *e = std::getenv("ZWZ_IO_THREADS");
    int t = (e && *e) ? std::atoi(e) : 4;
    return t < 1 ? 1 : (t > 32 ? 32 : t);
}

// Streams the file through the device in pieces (the reference streams 1 024-byte reads through MD5_Update,
// verification.cpp:15-19; the update granularity does not change the digest). Returns "" if the file cannot be opened,
// like the reference (verification.cpp:8-11).
std::string md5_of_file_on(int device, const std::string &file_path) { return md5_of_file_ctx(ctx_for(device), file_path); }

std::string md5_of_file_ctx(zwz_ctx *ctx, const std::string &file_path) {
    FILE *f = std::fopen(file_path.c_str(), "rb");
    if (!f) {
        std::cerr << "Cannot open file: " << file_path << std::endl;
        return "";
    }
    const size_t piece = (size_t) 64 << 20; // multiple of 64
    struct Guard { // every exit — the throws below included — closes the file and returns both staging buffers
        zwz_ctx *ctx;
        FILE *f;
        void *pin = nullptr, *dev = nullptr;
        ~Guard() {
            if (f) std::fclose(f);
            if (pin) zwz_free_pinned(ctx, pin);
            if (dev) zwz_free_device(ctx, dev);
        }
    } g{ctx, f};
    if (zwz_malloc_pinned(ctx, piece, &g.pin) != ZWZ_OK || zwz_malloc_device(ctx, piece + 64, &g.dev) != ZWZ_OK)
        throw std::runtime_error(std::string("zwz: staging allocation failed: ") + zwz_last_error(ctx));
    void *const pin = g.pin, *const dev = g.dev;
    uint32_t state[4];
    zwz_md5_state_init(state, 1);
    uint64_t total = 0;
    uint8_t digest[16];
    bool done = false;
    while (!done) {
        size_t got = std::fread(pin, 1, piece, f);
        total += got;
        uint64_t off = 0;
        if (got == piece) {
            uint64_t len = got;
            if (zwz_memcpy_h2d(ctx, dev, pin, got) != ZWZ_OK || zwz_md5_update_device(ctx, state, (const uint8_t *) dev, &off, &len, 1, nullptr) != ZWZ_OK)
                throw std::runtime_error(std::string("zwz: md5 update failed: ") + zwz_last_error(ctx));
        } else {
            uint64_t full = got & ~(uint64_t) 63, tail = got - full;
            if (got && zwz_memcpy_h2d(ctx, dev, pin, got) != ZWZ_OK) throw std::runtime_error("zwz: h2d failed");
            if (full && zwz_md5_update_device(ctx, state, (const uint8_t *) dev, &off, &full, 1, nullptr) != ZWZ_OK)
                throw std::runtime_error(std::string("zwz: md5 update failed: ") + zwz_last_error(ctx));
            if (zwz_md5_final_device(ctx, state, (const uint8_t *) dev, &full, &tail, &total, 1, digest, nullptr) != ZWZ_OK)
                throw std::runtime_error(std::string("zwz: md5 final failed: ") + zwz_last_error(ctx));
            done = true;
        }
    }
    char hex[33];
    zwz_md5_hex(digest, hex);
    hex[32] = 0;
    return std::string(hex, 32);
}

std::string md5_of_file(const std::string &file_path) { return md5_of_file_on(config().device, file_path); }

bool is_md5_match(const std::string &file_path, const std::string &expected_md5) { return md5_of_file(file_path) == expected_md5; }

} // namespace zwzhost
