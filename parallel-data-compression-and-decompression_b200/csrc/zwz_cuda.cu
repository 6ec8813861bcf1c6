// zwz_cuda.cu — implementation of the C ABI in include/zwz_cuda.h over the sm_100a kernels in this directory.
//
// One zwz_ctx per GPU: a non-blocking stream, grow-only device arenas (chunk descriptors, match/token scratch, bulk
// staging for the host-buffer entry points) and a pinned arena for descriptors/results. No CPU fallback: every entry
// point either runs the CUDA kernels or returns an error.
#include "zwz_cuda.h"
#include "zwz_rt.h"

#include "zwz_common.cuh"
#include "md5.cuh"
#include "inflate.cuh"
#include "inflate_lanes.cuh"
#include "deflate_match.cuh"
#include "deflate_encode.cuh"

#include <algorithm>
#include <new>
#include <vector>

namespace zwz {

// warp-per-chunk gather of the variable-length streams out of their slots into one packed buffer
ZWZ_KERNEL pack_streams_kernel(const uint8_t *__restrict__ slots, const uint64_t *__restrict__ slot_off, const uint32_t *__restrict__ res,
                               uint8_t *packed, const uint64_t *__restrict__ packed_off, uint32_t n) {
    uint32_t c = blockIdx.x * (blockDim.x >> 5) + warp_id();
    if (c >= n) return;
    const uint8_t *s = slots + slot_off[c];
    uint8_t *d = packed + packed_off[c];
    uint32_t bytes = res[4u * c] + res[4u * c + 1u];
    unsigned lane = lane_id();
    // slots are 4-byte aligned; use word copies when the destination happens to be too
    if ((((uintptr_t) d) & 3u) == 0u) {
        uint32_t nw = bytes >> 2;
        for (uint32_t i = lane; i < nw; i += 32u) ((uint32_t *) d)[i] = ((const uint32_t *) s)[i];
        for (uint32_t i = (nw << 2) + lane; i < bytes; i += 32u) d[i] = s[i];
    } else {
        for (uint32_t i = lane; i < bytes; i += 32u) d[i] = s[i];
    }
}

// Adler-32 of arbitrary ranges, one warp each (exported for tests; the deflate path fuses it into lz_match_kernel)
ZWZ_KERNEL adler32_kernel(const uint8_t *__restrict__ data, const uint64_t *__restrict__ off, const uint32_t *__restrict__ len,
                          uint32_t *adler, uint32_t n) {
    uint32_t c = blockIdx.x * (blockDim.x >> 5) + warp_id();
    if (c >= n) return;
    uint32_t a = enc_adler_global(data + off[c], len[c]);
    if (lane_id() == 0) adler[c] = a;
}

// concatenation of inflated records into contiguous files: one warp per record
ZWZ_KERNEL gather_records_kernel(const uint8_t *__restrict__ slots, const uint64_t *__restrict__ src_off, const uint64_t *__restrict__ dst_off,
                                 const uint32_t *__restrict__ len, uint8_t *dst, uint32_t n) {
    uint32_t c = blockIdx.x * (blockDim.x >> 5) + warp_id();
    if (c >= n) return;
    const uint8_t *s = slots + src_off[c];
    uint8_t *d = dst + dst_off[c];
    uint32_t bytes = len[c];
    unsigned lane = lane_id();
    if (((((uintptr_t) d) | ((uintptr_t) s)) & 15u) == 0u) {
        uint32_t nv = bytes >> 4;
        for (uint32_t i = lane; i < nv; i += 32u) ((uint4 *) d)[i] = ((const uint4 *) s)[i];
        for (uint32_t i = (nv << 4) + lane; i < bytes; i += 32u) d[i] = s[i];
    } else {
        for (uint32_t i = lane; i < bytes; i += 32u) d[i] = s[i];
    }
}

struct Arena {
    void *p = nullptr;
    size_t cap = 0;
    bool pinned = false;
};

// Deferred MD5 of one host-buffer batch (zwz_compress_files_async / zwz_decompress_records_async): the bytes to hash stay
// resident in `data` while the digests are computed on the context's second stream; the caller collects them with zwz_wait.
// Two slots per context, used alternately: the MD5 of batch k runs while batch k+1 goes through the main stream.
struct DigestSlot {
    Arena data;   // device: compress side = the raw files of the batch; decompress side = the concatenated output files
    Arena meta;   // device: MD5 descriptors + digests
    Arena pin;    // page-locked: descriptors (upload source) + digests (download target)
    zwz_rt::zwz_sync_event_t ev_in{}, ev_done{};
    bool pending = false;
    uint32_t n = 0;
    size_t digest_off = 0;     // where the digests land inside `pin`
    uint8_t *digest_dst = nullptr;
    uint64_t ticket = 0;
};

} // namespace zwz

struct zwz_ctx {
    int device = 0;
    int sm_count = 0, cc_major = 0, cc_minor = 0;
    size_t total_mem = 0, smem_optin = 0;
    zwz_stream_t stream = nullptr;
    zwz_stream_t md5_stream = nullptr; // deferred digests (DigestSlot)
    zwz::DigestSlot slot[2];
    int next_slot = 0;
    uint64_t ticket_seq = 0;
    std::string err;
    uint64_t launches = 0;
    zwz::Arena meta, scratch, bulk_in, bulk_out, packed, pin_meta, pin_aux, counter;
    int inflate_mode = 0; // 0/1 = one warp per stream, 2 = one lane per stream, 3 = one warp per stream without the lane-parallel block decoder (ZWZ_INFLATE_MODE=warp|lanes|careful)
    bool arena_async = true; // ZWZ_ARENA_SYNC=1: plain cudaMalloc/cudaFree arenas (for A/B timing)
    bool trace = false;   // ZWZ_TRACE=1: wall-clock phases of the host-buffer calls on stderr (syncs the stream at every mark)
    size_t batch_raw_bytes = (size_t) 4 << 30; // raw bytes per internal deflate sub-batch (scratch = 6x that; ZWZ_BATCH_RAW_MB overrides)
    size_t last_res_off = 0, last_slot_off = 0; // where the last deflate call left results / slot offsets inside `meta`
    // optional per-kernel timing
    bool profiling = false;
    struct Span {
        int kind;
        zwz_rt::zwz_event_t a, b;
    };
    std::vector<Span> spans;
    double prof_ms[ZWZ_PROF_N] = {0};
    uint64_t prof_n[ZWZ_PROF_N] = {0};
};

namespace {

using zwz::Arena;

int fail(zwz_ctx *ctx, int code, const char *what) {
    if (ctx) {
        std::string cuda;
        zwz_rt::last_error(cuda);
        ctx->err = what;
        if (!cuda.empty()) ctx->err += std::string(": ") + cuda;
    }
    return code;
}

int reserve(zwz_ctx *ctx, Arena &a, size_t bytes, bool pinned) {
    if (a.cap >= bytes && a.p) return ZWZ_OK;
    const size_t old_cap = a.p ? a.cap : 0;
    if (a.p) {
        zwz_rt::stream_sync(ctx->stream);
        if (a.pinned) zwz_rt::free_pinned(a.p); else if (ctx->arena_async) zwz_rt::free_arena(a.p, ctx->stream); else zwz_rt::free_device(a.p);
        a.p = nullptr;
        a.cap = 0;
    }
    // grow-only, with headroom (batches of one job differ by a few percent) and geometric from the second time on: a job whose
    // batches hold more and more, smaller and smaller files (the size-sorted deal) would otherwise re-allocate its descriptor
    // arenas batch after batch — and a re-allocation waits for the device, i.e. for every other worker's kernels
    // (profiles/round2_notes.md: 50-500 ms `reserve` phases in the end-to-end trace)
    size_t want = bytes + bytes / 4 + 4096;
    if (old_cap) want = std::max(want, std::min(old_cap * 2, bytes + ((size_t) 1 << 30)));
    int rc = pinned ? zwz_rt::malloc_pinned(&a.p, want) : (ctx->arena_async ? zwz_rt::malloc_arena(&a.p, want, ctx->stream) : zwz_rt::malloc_device(&a.p, want));
    if (rc) {
        a.p = nullptr;
        return fail(ctx, ZWZ_E_NOMEM, pinned ? "pinned allocation failed" : "device allocation failed");
    }
    a.cap = want;
    a.pinned = pinned;
    return ZWZ_OK;
}

void release(zwz_ctx *ctx, Arena &a) {
    if (!a.p) return;
    if (a.pinned) zwz_rt::free_pinned(a.p); else if (ctx->arena_async) zwz_rt::free_arena(a.p, ctx->stream); else zwz_rt::free_device(a.p);
    a.p = nullptr;
    a.cap = 0;
}

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// RAII bracket: records an event pair around one kernel launch when profiling is on
struct ProfSpan {
    zwz_ctx *ctx;
    zwz_stream_t st;
    zwz_ctx::Span sp;
    bool on;
    ProfSpan(zwz_ctx *c, int kind, zwz_stream_t s) : ctx(c), st(s), on(c->profiling) {
        sp.kind = kind;
        if (on) zwz_rt::event_record(&sp.a, st);
    }
    ~ProfSpan() {
        if (on) {
            zwz_rt::event_record(&sp.b, st);
            ctx->spans.push_back(sp);
        }
    }
};

int check_launch(zwz_ctx *ctx, const char *what) {
    ctx->launches++;
    std::string msg;
    if (zwz_rt::last_error(msg)) {
        ctx->err = std::string(what) + ": " + msg;
        return ZWZ_E_CUDA;
    }
    return ZWZ_OK;
}

// ZWZ_TRACE=1: phase times of a host-buffer call (debugging aid: every mark synchronises the stream)
struct Trace {
    zwz_ctx *ctx;
    const char *call;
    double t0;
    std::string line;
    static double now() {
        struct timespec ts;
        clock_gettime(CLOCK_MONOTONIC, &ts);
        return (double) ts.tv_sec + 1e-9 * (double) ts.tv_nsec;
    }
    Trace(zwz_ctx *c, const char *name) : ctx(c), call(name), t0(c->trace ? now() : 0.0) {}
    void mark(const char *what) {
        if (!ctx->trace) return;
        zwz_rt::stream_sync(ctx->stream);
        double t = now();
        char buf[64];
        snprintf(buf, sizeof buf, " %s %.2f ms,", what, (t - t0) * 1e3);
        line += buf;
        t0 = t;
    }
    ~Trace() {
        if (ctx->trace && !line.empty()) fprintf(stderr, "[zwz trace %p] %s:%s\n", (void *) ctx, call, line.c_str());
    }
};

#ifndef ZWZ_L0_DEPTH
#define ZWZ_L0_DEPTH 12
#define ZWZ_L0_NICE 48
#endif
struct LevelParams {
    uint32_t depth, nice;
};
LevelParams level_params(int level) {
    // search effort per level (hash-chain candidates per position, stop length). 0 = default: the cheapest setting that
    // stays inside the 3 % size tolerance on every corpus class (tests/test_kernels.py::test_deflate_ratio...)
    static const LevelParams t[10] = {{ZWZ_L0_DEPTH, ZWZ_L0_NICE}, {4, 16}, {6, 24}, {8, 32}, {16, 64}, {24, 96}, {32, 128}, {64, 160}, {128, 258}, {512, 258}};
    if (level < 0 || level > 9) level = 0;
    return t[level];
}

} // namespace

extern "C" {

int zwz_abi_version(void) { return ZWZ_ABI_VERSION; }
int zwz_device_count(void) { return zwz_rt::device_count(); }

int zwz_init(int device, zwz_ctx **out) {
    if (!out) return ZWZ_E_ARG;
    *out = nullptr;
    if (device < 0 || device >= zwz_rt::device_count()) return ZWZ_E_NODEVICE;
    if (zwz_rt::set_device(device)) return ZWZ_E_NODEVICE;
    zwz_ctx *ctx = new (std::nothrow) zwz_ctx();
    if (!ctx) return ZWZ_E_NOMEM;
    ctx->device = device;
    if (zwz_rt::device_props(device, &ctx->sm_count, &ctx->cc_major, &ctx->cc_minor, &ctx->total_mem, &ctx->smem_optin)) {
        delete ctx;
        return ZWZ_E_NODEVICE;
    }
#ifndef ZWZ_EMU
    if (ctx->cc_major != 10) { // the fatbin holds sm_100a SASS only
        delete ctx;
        return ZWZ_E_NODEVICE;
    }
#endif
    if (ctx->smem_optin < zwz::MatchClass<3>::kSmem || zwz_rt::stream_create(&ctx->stream) ||
        zwz_rt::set_max_dyn_smem((const void *) zwz::lz_match_kernel<0>, zwz::MatchClass<0>::kSmem) ||
        zwz_rt::set_max_dyn_smem((const void *) zwz::lz_match_kernel<1>, zwz::MatchClass<1>::kSmem) ||
        zwz_rt::set_max_dyn_smem((const void *) zwz::lz_match_kernel<2>, zwz::MatchClass<2>::kSmem) ||
        zwz_rt::set_max_dyn_smem((const void *) zwz::lz_match_kernel<3>, zwz::MatchClass<3>::kSmem) ||
        zwz_rt::set_max_dyn_smem((const void *) zwz::md5_files_staged_kernel, ZWZ_MD5S_SMEM) ||
        zwz_rt::set_max_dyn_smem((const void *) zwz::inflate_kernel, ZWZ_INF_SMEM) ||
        zwz_rt::set_max_dyn_smem((const void *) zwz::inflate_lanes_kernel, ZWZ_IL_SMEM)) {
        delete ctx;
        return ZWZ_E_NODEVICE;
    }
    if (zwz_rt::stream_create(&ctx->md5_stream)) {
        delete ctx;
        return ZWZ_E_NODEVICE;
    }
    for (auto &sl : ctx->slot)
        if (zwz_rt::sync_event_create(&sl.ev_in) || zwz_rt::sync_event_create(&sl.ev_done)) {
            delete ctx;
            return ZWZ_E_NODEVICE;
        }
    zwz_rt::keep_pool_memory(device);
    zwz_rt::preload_kernel((const void *) zwz::deflate_encode_kernel);
    zwz_rt::preload_kernel((const void *) zwz::inflate_kernel);
    zwz_rt::preload_kernel((const void *) zwz::md5_files_kernel);
    zwz_rt::preload_kernel((const void *) zwz::pack_streams_kernel);
    zwz_rt::preload_kernel((const void *) zwz::gather_records_kernel);
    zwz_rt::preload_kernel((const void *) zwz::adler32_kernel);
    if (const char *e = getenv("ZWZ_ARENA_SYNC")) ctx->arena_async = !(*e && *e != '0');
    if (const char *e = getenv("ZWZ_TRACE")) ctx->trace = *e && *e != '0';
    if (const char *e = getenv("ZWZ_INFLATE_MODE")) ctx->inflate_mode = !strcmp(e, "warp") ? 1 : (!strcmp(e, "lanes") ? 2 : (!strcmp(e, "careful") ? 3 : 0));
    if (const char *e = getenv("ZWZ_BATCH_RAW_MB")) {
        long v = atol(e);
        if (v >= 1) ctx->batch_raw_bytes = (size_t) v << 20;
    }
    *out = ctx;
    return ZWZ_OK;
}

void zwz_destroy(zwz_ctx *ctx) {
    if (!ctx) return;
    zwz_rt::set_device(ctx->device);
    zwz_rt::stream_sync(ctx->stream);
    zwz_rt::stream_sync(ctx->md5_stream);
    for (auto &sl : ctx->slot) {
        release(ctx, sl.data);
        release(ctx, sl.meta);
        release(ctx, sl.pin);
        zwz_rt::sync_event_destroy(sl.ev_in);
        zwz_rt::sync_event_destroy(sl.ev_done);
    }
    release(ctx, ctx->meta);
    release(ctx, ctx->scratch);
    release(ctx, ctx->bulk_in);
    release(ctx, ctx->bulk_out);
    release(ctx, ctx->packed);
    release(ctx, ctx->pin_meta);
    release(ctx, ctx->pin_aux);
    release(ctx, ctx->counter);
    zwz_rt::stream_sync(ctx->stream);
    zwz_rt::stream_destroy(ctx->stream);
    zwz_rt::stream_destroy(ctx->md5_stream);
    delete ctx;
}

const char *zwz_last_error(const zwz_ctx *ctx) { return ctx ? ctx->err.c_str() : "null ctx"; }

int zwz_ctx_tune(zwz_ctx *ctx, int what, uint64_t value) {
    if (!ctx) return ZWZ_E_ARG;
    if (what == ZWZ_TUNE_DEFLATE_SUBBATCH_BYTES) {
        if (value < ((uint64_t) 1 << 20)) return fail(ctx, ZWZ_E_ARG, "sub-batch below 1 MiB");
        ctx->batch_raw_bytes = (size_t) value;
        return ZWZ_OK;
    }
    return fail(ctx, ZWZ_E_ARG, "unknown tuning knob");
}

int zwz_device_props(const zwz_ctx *ctx, int *sm_count, int *cc_major, int *cc_minor, size_t *total_mem) {
    if (!ctx) return ZWZ_E_ARG;
    if (sm_count) *sm_count = ctx->sm_count;
    if (cc_major) *cc_major = ctx->cc_major;
    if (cc_minor) *cc_minor = ctx->cc_minor;
    if (total_mem) *total_mem = ctx->total_mem;
    return ZWZ_OK;
}

uint64_t zwz_launch_count(const zwz_ctx *ctx) { return ctx ? ctx->launches : 0; }

int zwz_profile_enable(zwz_ctx *ctx, int on) {
    if (!ctx) return ZWZ_E_ARG;
    ctx->profiling = on != 0;
    return ZWZ_OK;
}
int zwz_profile_read(zwz_ctx *ctx, double *ms, uint64_t *launches, int reset) {
    if (!ctx) return ZWZ_E_ARG;
#ifndef ZWZ_EMU
    if (cudaDeviceSynchronize() != cudaSuccess) return fail(ctx, ZWZ_E_CUDA, "device sync failed");
#endif
    for (auto &sp : ctx->spans) {
        ctx->prof_ms[sp.kind] += zwz_rt::event_elapsed_and_free(sp.a, sp.b);
        ctx->prof_n[sp.kind] += 1;
    }
    ctx->spans.clear();
    for (int k = 0; k < ZWZ_PROF_N; ++k) {
        if (ms) ms[k] = ctx->prof_ms[k];
        if (launches) launches[k] = ctx->prof_n[k];
        if (reset) {
            ctx->prof_ms[k] = 0;
            ctx->prof_n[k] = 0;
        }
    }
    return ZWZ_OK;
}

int zwz_malloc_device(zwz_ctx *ctx, size_t bytes, void **ptr) {
    if (!ctx || !ptr) return ZWZ_E_ARG;
    zwz_rt::set_device(ctx->device);
    return zwz_rt::malloc_device(ptr, bytes) ? fail(ctx, ZWZ_E_NOMEM, "device allocation failed") : ZWZ_OK;
}
int zwz_free_device(zwz_ctx *ctx, void *ptr) {
    if (!ctx) return ZWZ_E_ARG;
    zwz_rt::set_device(ctx->device);
    zwz_rt::stream_sync(ctx->stream);
    return zwz_rt::free_device(ptr) ? fail(ctx, ZWZ_E_CUDA, "cudaFree failed") : ZWZ_OK;
}
int zwz_malloc_pinned(zwz_ctx *ctx, size_t bytes, void **ptr) {
    if (!ctx || !ptr) return ZWZ_E_ARG;
    zwz_rt::set_device(ctx->device);
    return zwz_rt::malloc_pinned(ptr, bytes) ? fail(ctx, ZWZ_E_NOMEM, "pinned allocation failed") : ZWZ_OK;
}
int zwz_free_pinned(zwz_ctx *ctx, void *ptr) {
    if (!ctx) return ZWZ_E_ARG;
    return zwz_rt::free_pinned(ptr) ? fail(ctx, ZWZ_E_CUDA, "cudaFreeHost failed") : ZWZ_OK;
}
int zwz_memcpy_h2d(zwz_ctx *ctx, void *dst, const void *src, size_t bytes) {
    if (!ctx) return ZWZ_E_ARG;
    zwz_rt::set_device(ctx->device);
    if (zwz_rt::memcpy_h2d(dst, src, bytes, ctx->stream) || zwz_rt::stream_sync(ctx->stream)) return fail(ctx, ZWZ_E_CUDA, "h2d copy failed");
    return ZWZ_OK;
}
int zwz_memcpy_d2h(zwz_ctx *ctx, void *dst, const void *src, size_t bytes) {
    if (!ctx) return ZWZ_E_ARG;
    zwz_rt::set_device(ctx->device);
    if (zwz_rt::memcpy_d2h(dst, src, bytes, ctx->stream) || zwz_rt::stream_sync(ctx->stream)) return fail(ctx, ZWZ_E_CUDA, "d2h copy failed");
    return ZWZ_OK;
}
int zwz_sync(zwz_ctx *ctx) {
    if (!ctx) return ZWZ_E_ARG;
    return zwz_rt::stream_sync(ctx->stream) ? fail(ctx, ZWZ_E_CUDA, "stream sync failed") : ZWZ_OK;
}

// ======================================================================================================================
// deflate
// ======================================================================================================================
int zwz_deflate_batch_device(zwz_ctx *ctx, const uint8_t *d_raw, const uint64_t *off, const uint32_t *len, uint32_t n, uint8_t *d_out,
                             const uint64_t *out_off, zwz_deflate_result *res, int level, void *stream_v) {
    if (!ctx) return ZWZ_E_ARG;
    if (n == 0) return ZWZ_OK;
    if (!off || !len || !out_off || !res || !d_out) return fail(ctx, ZWZ_E_ARG, "null argument");
    zwz_rt::set_device(ctx->device);
    zwz_stream_t st = stream_v ? (zwz_stream_t) stream_v : ctx->stream;
    uint64_t total_raw = 0;
    for (uint32_t i = 0; i < n; ++i) {
        if (len[i] > ZWZ_CHUNK_SIZE) return fail(ctx, ZWZ_E_ARG, "chunk longer than 65535 bytes");
        if (out_off[i] & 3u) return fail(ctx, ZWZ_E_ARG, "out_off must be a multiple of 4");
        total_raw += len[i];
    }
    if (total_raw && !d_raw) return fail(ctx, ZWZ_E_ARG, "null raw buffer");

    // descriptors: [raw_off u64][scr_off u64][out_off u64][raw_len u32][order u32] per chunk, then results + adler on the device
    const size_t meta_bytes = (size_t) n * (8 + 8 + 8 + 4 + 4);
    const size_t dev_meta_bytes = align_up(meta_bytes, 256) + (size_t) n * 16 + (size_t) n * 8 + 256;
    int rc;
    if ((rc = reserve(ctx, ctx->pin_meta, std::max(meta_bytes, (size_t) n * 16), true))) return rc;
    if ((rc = reserve(ctx, ctx->meta, dev_meta_bytes, false))) return rc;

    uint64_t *h_raw_off = (uint64_t *) ctx->pin_meta.p;
    uint64_t *h_scr_off = h_raw_off + n;
    uint64_t *h_out_off = h_scr_off + n;
    uint32_t *h_len = (uint32_t *) (h_out_off + n);
    uint32_t *h_order = h_len + n;

    // sub-batches bounded by scratch: scratch offsets restart at 0 in every sub-batch
    std::vector<uint32_t> sub_begin;
    size_t max_scr = 0;
    {
        size_t scr = 0, raw = 0;
        sub_begin.push_back(0);
        for (uint32_t i = 0; i < n; ++i) {
            if (raw + len[i] > ctx->batch_raw_bytes && i > sub_begin.back()) {
                max_scr = std::max(max_scr, scr);
                sub_begin.push_back(i);
                scr = 0;
                raw = 0;
            }
            h_raw_off[i] = off[i];
            h_scr_off[i] = scr;
            h_out_off[i] = out_off[i];
            h_len[i] = len[i];
            // uint32 entries: best match per position (+ pad so the parser may peek one past the end), then the u16 position
            // lists of lz_match_kernel (deflate_match.cuh: dm_scratch_match_words)
            scr += align_up((size_t) len[i] + 2, 32) + align_up(((size_t) len[i] + 1) / 2 + 32, 32) + ZWZ_SCR_FLAG_WORDS; // zwz_common.cuh: scratch layout
            raw += len[i];
        }
        max_scr = std::max(max_scr, scr);
        sub_begin.push_back(n);
    }
    if ((rc = reserve(ctx, ctx->scratch, max_scr * 4 + 256, false))) return rc;
    // size classes (deflate_match.cuh): per sub-batch, chunk indices grouped by class, longest first inside a class so the
    // persistent CTAs end together
    const size_t nsub = sub_begin.size() - 1;
    std::vector<uint32_t> cls_begin(nsub * 5);
    for (size_t s = 0; s < nsub; ++s) {
        uint32_t b = sub_begin[s], e = sub_begin[s + 1];
        uint32_t *o = h_order + b;
        uint32_t k = 0;
        static const uint32_t kClassHi[4] = {zwz::MatchClass<0>::kCap, zwz::MatchClass<1>::kCap, zwz::MatchClass<2>::kCap, zwz::MatchClass<3>::kCap};
        for (int cls = 0; cls < 4; ++cls) {
            cls_begin[s * 5 + cls] = k;
            uint32_t lo_len = cls == 0 ? 0u : kClassHi[cls - 1] + 1u, hi_len = kClassHi[cls];
            uint32_t k0 = k;
            for (uint32_t i = b; i < e; ++i)
                if (len[i] >= lo_len && len[i] <= hi_len) o[k++] = i - b;
            auto longer = [&](uint32_t x, uint32_t y) { return len[b + x] > len[b + y]; };
            if (!std::is_sorted(o + k0, o + k, longer)) std::stable_sort(o + k0, o + k, longer); // size-sorted shards (the reference's deal) skip this
        }
        cls_begin[s * 5 + 4] = k;
    }
    if ((rc = reserve(ctx, ctx->counter, nsub * 32 + 256, false))) return rc;
    if (zwz_rt::memset_device(ctx->counter.p, 0, nsub * 32, st)) return fail(ctx, ZWZ_E_CUDA, "memset failed");

    uint8_t *dm = (uint8_t *) ctx->meta.p;
    if (zwz_rt::memcpy_h2d(dm, ctx->pin_meta.p, meta_bytes, st)) return fail(ctx, ZWZ_E_CUDA, "descriptor upload failed");
    uint32_t *d_res = (uint32_t *) (dm + align_up(meta_bytes, 256));
    uint32_t *d_adler = d_res + (size_t) n * 4;
    uint32_t *d_nmatch = d_adler + n;
    ctx->last_res_off = align_up(meta_bytes, 256);
    ctx->last_slot_off = 2 * (size_t) n * 8;

    const LevelParams lp = level_params(level);
    for (size_t s = 0; s + 1 < sub_begin.size(); ++s) {
        uint32_t b = sub_begin[s], e = sub_begin[s + 1];
        zwz::DeflateJob job;
        job.raw = d_raw;
        job.raw_off = (const uint64_t *) dm + b;
        job.scr_off = (const uint64_t *) dm + n + b;
        job.out_off = (const uint64_t *) dm + 2 * (size_t) n + b;
        job.raw_len = (const uint32_t *) ((const uint64_t *) dm + 3 * (size_t) n) + b;
        job.scratch = (uint32_t *) ctx->scratch.p;
        job.adler = d_adler + b;
        job.nmatch = d_nmatch + b;
        job.out = d_out;
        job.res = d_res + (size_t) b * 4;
        job.n = e - b;
        job.depth = lp.depth;
        job.nice = lp.nice;
        job.work_counter = nullptr;
        const uint32_t *d_order = (const uint32_t *) ((const uint64_t *) dm + 3 * (size_t) n) + n + b;
        uint32_t *d_counters = (uint32_t *) ctx->counter.p + s * 8; // [0..3] match classes, [4] encoder
        for (int cls = 0; cls < 4; ++cls) {
            uint32_t w0 = cls_begin[s * 5 + cls], w1 = cls_begin[s * 5 + cls + 1];
            if (w1 == w0) continue;
            ProfSpan ps(ctx, ZWZ_PROF_MATCH, st);
            const uint32_t nwork = w1 - w0, sms = (uint32_t) ctx->sm_count;
            if (cls == 0) {
                ZWZ_LAUNCH(zwz::lz_match_kernel<0>, std::min(nwork, sms * 6u), zwz::MatchClass<0>::kThreads, zwz::MatchClass<0>::kSmem, st, job,
                           d_order + w0, nwork, d_counters + cls);
            } else if (cls == 1) {
                ZWZ_LAUNCH(zwz::lz_match_kernel<1>, std::min(nwork, sms * 3u), zwz::MatchClass<1>::kThreads, zwz::MatchClass<1>::kSmem, st, job,
                           d_order + w0, nwork, d_counters + cls);
            } else if (cls == 2) {
                ZWZ_LAUNCH(zwz::lz_match_kernel<2>, std::min(nwork, sms * 2u), zwz::MatchClass<2>::kThreads, zwz::MatchClass<2>::kSmem, st, job,
                           d_order + w0, nwork, d_counters + cls);
            } else {
                ZWZ_LAUNCH(zwz::lz_match_kernel<3>, std::min(nwork, sms), zwz::MatchClass<3>::kThreads, zwz::MatchClass<3>::kSmem, st, job,
                           d_order + w0, nwork, d_counters + cls);
            }
            ctx->launches++;
        }
        if ((rc = check_launch(ctx, "lz_match_kernel"))) return rc;
        ctx->launches--; // check_launch counted one more
        uint32_t grid2 = std::min<uint32_t>((job.n + ZWZ_DE_WARPS - 1) / ZWZ_DE_WARPS, (uint32_t) ctx->sm_count * 8u);
        {
            ProfSpan ps(ctx, ZWZ_PROF_ENCODE, st);
            ZWZ_LAUNCH(zwz::deflate_encode_kernel, grid2, ZWZ_DE_WARPS * 32, ZWZ_DE_SMEM, st, job, d_counters + 4);
        }
        if ((rc = check_launch(ctx, "deflate_encode_kernel"))) return rc;
    }
    // results come back through the pinned arena (descriptors are no longer needed once the kernels are queued ... but
    // the H2D above must have completed before we overwrite it: same stream => ordered)
    if (zwz_rt::memcpy_d2h(ctx->pin_meta.p, d_res, (size_t) n * 16, st) || zwz_rt::stream_sync(st))
        return fail(ctx, ZWZ_E_CUDA, "deflate kernels failed");
    memcpy(res, ctx->pin_meta.p, (size_t) n * 16);
    return ZWZ_OK;
}

static int md5_launch(zwz_ctx *ctx, uint32_t *state, const uint8_t *d_data, const uint64_t *off, const uint64_t *len, const uint64_t *total_len,
                      uint32_t n, uint8_t *digest, int finalize, void *stream_v);
static int slot_finish(zwz_ctx *ctx, zwz::DigestSlot &sl);
static int slot_acquire(zwz_ctx *ctx, zwz::DigestSlot **out);
static int slot_queue_md5(zwz_ctx *ctx, zwz::DigestSlot &sl, const uint64_t *off, const uint64_t *len, uint32_t n, uint8_t *digest, uint64_t *ticket);

static int deflate_host_impl(zwz_ctx *ctx, const uint8_t *raw, const uint64_t *off, const uint32_t *len, uint32_t n, uint8_t *out, uint64_t out_cap,
                             uint64_t *packed_off, zwz_deflate_result *res, int level, const uint64_t *md5_off, const uint64_t *md5_len,
                             uint32_t md5_n, uint8_t *digest, uint64_t *ticket) {
    if (!ctx) return ZWZ_E_ARG;
    if (ticket) *ticket = 0;
    if (!packed_off) return fail(ctx, ZWZ_E_ARG, "null argument");
    packed_off[0] = 0;
    if (n == 0) return ZWZ_OK;
    if (!off || !len || !res || !out) return fail(ctx, ZWZ_E_ARG, "null argument");
    zwz_rt::set_device(ctx->device);
    // span of the host buffer the chunks touch
    uint64_t lo = ~0ull, hi = 0;
    for (uint32_t i = 0; i < n; ++i) {
        if (len[i] > ZWZ_CHUNK_SIZE) return fail(ctx, ZWZ_E_ARG, "chunk longer than 65535 bytes");
        lo = std::min(lo, off[i]);
        hi = std::max(hi, off[i] + len[i]);
    }
    if (hi < lo) lo = hi = 0;
    std::vector<uint64_t> roff(n), slot(n + 1);
    slot[0] = 0;
    for (uint32_t i = 0; i < n; ++i) {
        roff[i] = off[i] - lo;
        slot[i + 1] = slot[i] + zwz_deflate_bound(len[i]);
    }
    int rc;
    Trace tr(ctx, "compress");
    // the raw bytes live in a digest slot: their MD5 runs on the second stream and may outlive this call (ticket != NULL)
    zwz::DigestSlot *sl = nullptr;
    if ((rc = slot_acquire(ctx, &sl))) return rc;
    if ((rc = reserve(ctx, sl->data, (size_t) (hi - lo) + 64, false))) return rc;
    if ((rc = reserve(ctx, ctx->bulk_out, (size_t) slot[n] + 64, false))) return rc;
    tr.mark("reserve");
    if (zwz_rt::memcpy_h2d(sl->data.p, raw + lo, (size_t) (hi - lo), ctx->stream)) return fail(ctx, ZWZ_E_CUDA, "raw upload failed");
    tr.mark("h2d");
    if (digest && md5_n) { // MD5 of the source files from the same resident copy (no second upload), beside the deflate kernels
        std::vector<uint64_t> mo(md5_n);
        for (uint32_t i = 0; i < md5_n; ++i) mo[i] = md5_off[i] - lo;
        if ((rc = slot_queue_md5(ctx, *sl, mo.data(), md5_len, md5_n, digest, ticket))) return rc;
    }
    struct Disarm { // an error below must not leave a pointer into the caller's memory behind in the slot
        zwz::DigestSlot *s;
        bool ok = false;
        ~Disarm() {
            if (!ok) s->digest_dst = nullptr;
        }
    } disarm{sl};
    if ((rc = zwz_deflate_batch_device(ctx, (const uint8_t *) sl->data.p, roff.data(), len, n, (uint8_t *) ctx->bulk_out.p, slot.data(), res,
                                       level, nullptr)))
        return rc;
    tr.mark("md5+deflate");
    for (uint32_t i = 0; i < n; ++i) packed_off[i + 1] = packed_off[i] + res[i].len0 + res[i].len1;
    if (packed_off[n] > out_cap) return fail(ctx, ZWZ_E_CAPACITY, "output buffer too small");
    // gather on the device so only the compressed bytes cross the bus
    const size_t pm = (size_t) (n + 1) * 8;
    if ((rc = reserve(ctx, ctx->packed, (size_t) packed_off[n] + align_up(pm, 256) + 256, false))) return rc;
    if ((rc = reserve(ctx, ctx->pin_meta, pm, true))) return rc;
    tr.mark("reserve2");
    memcpy(ctx->pin_meta.p, packed_off, pm);
    uint8_t *d_poff = (uint8_t *) ctx->packed.p;
    uint8_t *d_packed = d_poff + align_up(pm, 256);
    if (zwz_rt::memcpy_h2d(d_poff, ctx->pin_meta.p, pm, ctx->stream)) return fail(ctx, ZWZ_E_CUDA, "offset upload failed");
    // slot offsets and results are still in ctx->meta from the call above
    const uint8_t *dm = (const uint8_t *) ctx->meta.p;
    const uint64_t *d_slot_off = (const uint64_t *) (dm + ctx->last_slot_off);
    const uint32_t *d_res = (const uint32_t *) (dm + ctx->last_res_off);
    {
        ProfSpan ps(ctx, ZWZ_PROF_PACK, ctx->stream);
        ZWZ_LAUNCH(zwz::pack_streams_kernel, (n + 7) / 8, 256, 0, ctx->stream, (const uint8_t *) ctx->bulk_out.p, d_slot_off, d_res, d_packed,
                   (const uint64_t *) d_poff, n);
    }
    if ((rc = check_launch(ctx, "pack_streams_kernel"))) return rc;
    tr.mark("pack");
    if (zwz_rt::memcpy_d2h(out, d_packed, (size_t) packed_off[n], ctx->stream) || zwz_rt::stream_sync(ctx->stream))
        return fail(ctx, ZWZ_E_CUDA, "compressed download failed");
    tr.mark("d2h");
    disarm.ok = true;
    if (!ticket) return slot_finish(ctx, *sl); // synchronous form: the digests are part of the result
    return ZWZ_OK;
}

int zwz_pack_streams_device(zwz_ctx *ctx, const uint8_t *d_slots, const uint64_t *slot_off, const zwz_deflate_result *res, uint32_t n,
                            uint8_t *d_packed, uint64_t *packed_off, void *stream_v) {
    if (!ctx) return ZWZ_E_ARG;
    if (!packed_off) return fail(ctx, ZWZ_E_ARG, "null argument");
    packed_off[0] = 0;
    if (n == 0) return ZWZ_OK;
    if (!slot_off || !res || !d_slots || !d_packed) return fail(ctx, ZWZ_E_ARG, "null argument");
    zwz_rt::set_device(ctx->device);
    zwz_stream_t st = stream_v ? (zwz_stream_t) stream_v : ctx->stream;
    for (uint32_t i = 0; i < n; ++i) packed_off[i + 1] = packed_off[i] + res[i].len0 + res[i].len1;
    const size_t m_poff = (size_t) n * 8, m_res = m_poff + (size_t) (n + 1) * 8, meta_bytes = m_res + (size_t) n * 16;
    int rc;
    if ((rc = reserve(ctx, ctx->pin_meta, meta_bytes, true))) return rc;
    if ((rc = reserve(ctx, ctx->meta, meta_bytes + 256, false))) return rc;
    uint8_t *hp = (uint8_t *) ctx->pin_meta.p;
    memcpy(hp, slot_off, (size_t) n * 8);
    memcpy(hp + m_poff, packed_off, (size_t) (n + 1) * 8);
    memcpy(hp + m_res, res, (size_t) n * 16);
    uint8_t *dm = (uint8_t *) ctx->meta.p;
    if (zwz_rt::memcpy_h2d(dm, hp, meta_bytes, st)) return fail(ctx, ZWZ_E_CUDA, "descriptor upload failed");
    {
        ProfSpan ps(ctx, ZWZ_PROF_PACK, st);
        ZWZ_LAUNCH(zwz::pack_streams_kernel, (n + 7) / 8, 256, 0, st, d_slots, (const uint64_t *) dm, (const uint32_t *) (dm + m_res), d_packed,
                   (const uint64_t *) (dm + m_poff), n);
    }
    if ((rc = check_launch(ctx, "pack_streams_kernel"))) return rc;
    if (zwz_rt::stream_sync(st)) return fail(ctx, ZWZ_E_CUDA, "pack kernel failed");
    return ZWZ_OK;
}

int zwz_deflate_batch(zwz_ctx *ctx, const uint8_t *raw, const uint64_t *off, const uint32_t *len, uint32_t n, uint8_t *out, uint64_t out_cap,
                      uint64_t *packed_off, zwz_deflate_result *res, int level) {
    return deflate_host_impl(ctx, raw, off, len, n, out, out_cap, packed_off, res, level, nullptr, nullptr, 0, nullptr, nullptr);
}

uint64_t zwz_count_chunks(const uint64_t *file_off, uint32_t nf) {
    uint64_t c = 0;
    for (uint32_t i = 0; i < nf; ++i) c += (file_off[i + 1] - file_off[i]) / ZWZ_CHUNK_SIZE + 1;
    return c;
}

static int compress_files_impl(zwz_ctx *ctx, const uint8_t *data, const uint64_t *file_off, uint32_t nf, int level, uint8_t *out, uint64_t out_cap,
                               uint64_t *packed_off, zwz_deflate_result *res, uint8_t *digest, uint64_t *ticket) {
    if (!ctx) return ZWZ_E_ARG;
    if (!file_off || !packed_off) return fail(ctx, ZWZ_E_ARG, "null argument");
    uint64_t nc = zwz_count_chunks(file_off, nf);
    if (nc > 0xffffffffull) return fail(ctx, ZWZ_E_ARG, "too many chunks in one batch");
    std::vector<uint64_t> coff((size_t) nc), flen(nf);
    std::vector<uint32_t> clen((size_t) nc);
    size_t k = 0;
    for (uint32_t i = 0; i < nf; ++i) { // compression.cpp:52-64
        uint64_t S = file_off[i + 1] - file_off[i];
        flen[i] = S;
        for (uint64_t o = 0;; o += ZWZ_CHUNK_SIZE) {
            uint64_t left = S - o;
            coff[k] = file_off[i] + o;
            clen[k] = (uint32_t) (left < ZWZ_CHUNK_SIZE ? left : ZWZ_CHUNK_SIZE);
            ++k;
            if (left < ZWZ_CHUNK_SIZE) break;
        }
    }
    return deflate_host_impl(ctx, data, coff.data(), clen.data(), (uint32_t) nc, out, out_cap, packed_off, res, level, file_off, flen.data(),
                             digest ? nf : 0, digest, ticket);
}

int zwz_compress_files(zwz_ctx *ctx, const uint8_t *data, const uint64_t *file_off, uint32_t nf, int level, uint8_t *out, uint64_t out_cap,
                       uint64_t *packed_off, zwz_deflate_result *res, uint8_t *digest) {
    return compress_files_impl(ctx, data, file_off, nf, level, out, out_cap, packed_off, res, digest, nullptr);
}

int zwz_compress_files_async(zwz_ctx *ctx, const uint8_t *data, const uint64_t *file_off, uint32_t nf, int level, uint8_t *out, uint64_t out_cap,
                             uint64_t *packed_off, zwz_deflate_result *res, uint8_t *digest, uint64_t *ticket) {
    if (!ticket) return ctx ? fail(ctx, ZWZ_E_ARG, "null ticket") : ZWZ_E_ARG;
    return compress_files_impl(ctx, data, file_off, nf, level, out, out_cap, packed_off, res, digest, ticket);
}

int zwz_wait(zwz_ctx *ctx, uint64_t ticket) {
    if (!ctx) return ZWZ_E_ARG;
    if (ticket == 0) return ZWZ_OK;
    zwz_rt::set_device(ctx->device);
    for (auto &sl : ctx->slot)
        if (sl.pending && sl.ticket == ticket) return slot_finish(ctx, sl);
    return ZWZ_OK; // already delivered (a later batch needed the slot)
}

// ======================================================================================================================
// inflate
// ======================================================================================================================
int zwz_inflate_batch_device(zwz_ctx *ctx, const uint8_t *d_comp, const uint64_t *off, const uint32_t *len, uint32_t n, uint8_t *d_raw_out,
                             const uint64_t *raw_off, uint32_t *raw_len, uint32_t *status, uint32_t flags, void *stream_v) {
    if (!ctx) return ZWZ_E_ARG;
    if (n == 0) return ZWZ_OK;
    if (!off || !len || !raw_off || !raw_len || !status) return fail(ctx, ZWZ_E_ARG, "null argument");
    zwz_rt::set_device(ctx->device);
    zwz_stream_t st = stream_v ? (zwz_stream_t) stream_v : ctx->stream;
    // Mapping: one warp per stream (inflate.cuh) is the product path. One lane per stream (inflate_lanes.cuh,
    // ZWZ_INFLATE_MODE=lanes) needs ~40x fewer warp-instructions per byte but only three warps fit an SM next to its
    // per-lane tables, and at that occupancy it runs latency-bound: 140 ms against 68 ms per C2 step on B200. It stays as a
    // tested alternative mapping. It keeps its Adler-32 sums unreduced and its bit counts in 32 bits, hence the size limits.
    bool lanes = ctx->inflate_mode == 2; // opt-in only: measured slower than the warp mapping on B200 (inflate)
    for (uint32_t i = 0; lanes && i < n; ++i)
        if (len[i] >= (1u << 28) || raw_off[i + 1] - raw_off[i] > (1ull << 24)) lanes = false;
    const size_t m_off = 0, m_roff = (size_t) n * 8, m_len = m_roff + (size_t) (n + 1) * 8, m_order = m_len + (size_t) n * 4,
                 meta_bytes = m_order + (lanes ? (size_t) n * 4 : 0);
    const size_t r_base = align_up(meta_bytes, 256);
    int rc;
    if ((rc = reserve(ctx, ctx->pin_meta, std::max(meta_bytes, (size_t) n * 8), true))) return rc;
    if ((rc = reserve(ctx, ctx->meta, r_base + (size_t) n * 8 + 512, false))) return rc;
    uint8_t *hp = (uint8_t *) ctx->pin_meta.p;
    memcpy(hp + m_off, off, (size_t) n * 8);
    memcpy(hp + m_roff, raw_off, (size_t) (n + 1) * 8);
    memcpy(hp + m_len, len, (size_t) n * 4);
    if (lanes) {
        // streams by decreasing compressed size (counting sort on len / 32), so that the 32 lanes of a warp finish together
        // and the longest groups start first
        constexpr uint32_t NB = 4096;
        std::vector<uint32_t> head(NB + 1, 0u);
        auto bucket = [](uint32_t l) { return NB - 1u - std::min<uint32_t>(l >> 5, NB - 1u); };
        for (uint32_t i = 0; i < n; ++i) head[bucket(len[i]) + 1]++;
        for (uint32_t b = 0; b < NB; ++b) head[b + 1] += head[b];
        uint32_t *order = (uint32_t *) (hp + m_order);
        for (uint32_t i = 0; i < n; ++i) order[head[bucket(len[i])]++] = i;
    }
    uint8_t *dm = (uint8_t *) ctx->meta.p;
    if (zwz_rt::memcpy_h2d(dm, hp, meta_bytes, st)) return fail(ctx, ZWZ_E_CUDA, "descriptor upload failed");
    uint32_t *d_rlen = (uint32_t *) (dm + r_base);
    uint32_t *d_stat = d_rlen + n;
    uint32_t *d_counter = d_stat + n; // work queue head of the persistent warps
    if (zwz_rt::memset_device(d_counter, 0, 4, st)) return fail(ctx, ZWZ_E_CUDA, "memset failed");
    {
        ProfSpan ps(ctx, ZWZ_PROF_INFLATE, st);
        if (lanes) {
            uint32_t grid = std::min<uint32_t>((n + 31u) / 32u, (uint32_t) ctx->sm_count * 3u);
            ZWZ_LAUNCH(zwz::inflate_lanes_kernel, grid, 32, ZWZ_IL_SMEM, st, d_comp, (const uint64_t *) (dm + m_off),
                       (const uint32_t *) (dm + m_len), d_raw_out, (const uint64_t *) (dm + m_roff), d_rlen, d_stat,
                       (const uint32_t *) (dm + m_order), n, flags, d_counter);
        } else {
            const uint32_t per_sm = std::min<uint32_t>(32u, std::max<uint32_t>(1u, (uint32_t) ((228u * 1024u) / (ZWZ_INF_SMEM + 1024u))));
            uint32_t grid = std::min<uint32_t>((n + ZWZ_INF_WARPS - 1) / ZWZ_INF_WARPS, (uint32_t) ctx->sm_count * per_sm);
            if (ctx->inflate_mode == 3) flags |= ZWZ_INFLATE_CAREFUL;
            ZWZ_LAUNCH(zwz::inflate_kernel, grid, ZWZ_INF_WARPS * 32, ZWZ_INF_SMEM, st, d_comp, (const uint64_t *) (dm + m_off),
                       (const uint32_t *) (dm + m_len), d_raw_out, (const uint64_t *) (dm + m_roff), d_rlen, d_stat, n, flags, d_counter);
        }
    }
    if ((rc = check_launch(ctx, "inflate_kernel"))) return rc;
    if (zwz_rt::memcpy_d2h(hp, d_rlen, (size_t) n * 8, st) || zwz_rt::stream_sync(st)) return fail(ctx, ZWZ_E_CUDA, "inflate kernel failed");
    memcpy(raw_len, hp, (size_t) n * 4);
    memcpy(status, hp + (size_t) n * 4, (size_t) n * 4);
    return ZWZ_OK;
}

int zwz_inflate_batch(zwz_ctx *ctx, const uint8_t *comp, const uint64_t *off, const uint32_t *len, uint32_t n, uint8_t *raw_out,
                      const uint64_t *raw_off, uint32_t *raw_len, uint32_t *status, uint32_t flags) {
    if (!ctx) return ZWZ_E_ARG;
    if (n == 0) return ZWZ_OK;
    if (!off || !len || !raw_off || !raw_len || !status) return fail(ctx, ZWZ_E_ARG, "null argument");
    zwz_rt::set_device(ctx->device);
    uint64_t lo = ~0ull, hi = 0;
    for (uint32_t i = 0; i < n; ++i) {
        lo = std::min(lo, off[i]);
        hi = std::max(hi, off[i] + len[i]);
    }
    if (hi < lo) lo = hi = 0;
    std::vector<uint64_t> coff(n), roff(n + 1);
    for (uint32_t i = 0; i < n; ++i) coff[i] = off[i] - lo;
    for (uint32_t i = 0; i <= n; ++i) roff[i] = raw_off[i] - raw_off[0];
    int rc;
    if ((rc = reserve(ctx, ctx->bulk_in, (size_t) (hi - lo) + 64, false))) return rc;
    if ((rc = reserve(ctx, ctx->bulk_out, (size_t) roff[n] + 64, false))) return rc;
    if (zwz_rt::memcpy_h2d(ctx->bulk_in.p, comp + lo, (size_t) (hi - lo), ctx->stream)) return fail(ctx, ZWZ_E_CUDA, "compressed upload failed");
    if ((rc = zwz_inflate_batch_device(ctx, (const uint8_t *) ctx->bulk_in.p, coff.data(), len, n, (uint8_t *) ctx->bulk_out.p, roff.data(), raw_len,
                                       status, flags, nullptr)))
        return rc;
    if (zwz_rt::memcpy_d2h(raw_out + raw_off[0], ctx->bulk_out.p, (size_t) roff[n], ctx->stream) || zwz_rt::stream_sync(ctx->stream))
        return fail(ctx, ZWZ_E_CUDA, "raw download failed");
    return ZWZ_OK;
}

static int decompress_records_impl(zwz_ctx *ctx, const uint8_t *comp, const uint64_t *off, const uint32_t *len, const uint32_t *rec_cap,
                                   const uint32_t *rec_file, uint32_t n, uint32_t nf, uint8_t *files_out, uint64_t out_cap, uint64_t *file_off_out,
                                   uint32_t *raw_len, uint32_t *status, uint8_t *digest, uint32_t flags, uint64_t *ticket) {
    if (!ctx) return ZWZ_E_ARG;
    if (ticket) *ticket = 0;
    if (!file_off_out) return fail(ctx, ZWZ_E_ARG, "null argument");
    for (uint32_t f = 0; f <= nf; ++f) file_off_out[f] = 0;
    if (n == 0) {
        if (digest && nf) { // every file is empty
            std::vector<uint64_t> z(nf, 0);
            int rc0 = reserve(ctx, ctx->bulk_in, 64, false);
            if (rc0) return rc0;
            return md5_launch(ctx, nullptr, (const uint8_t *) ctx->bulk_in.p, z.data(), z.data(), nullptr, nf, digest, 1, nullptr);
        }
        return ZWZ_OK;
    }
    if (!off || !len || !rec_cap || !rec_file || !raw_len || !status) return fail(ctx, ZWZ_E_ARG, "null argument");
    zwz_rt::set_device(ctx->device);
    uint64_t lo = ~0ull, hi = 0;
    for (uint32_t i = 0; i < n; ++i) {
        lo = std::min(lo, off[i]);
        hi = std::max(hi, off[i] + len[i]);
        if (rec_file[i] >= nf || (i && rec_file[i] < rec_file[i - 1])) return fail(ctx, ZWZ_E_ARG, "rec_file must be non-decreasing and < nf");
    }
    std::vector<uint64_t> coff(n), slot(n + 1);
    slot[0] = 0;
    for (uint32_t i = 0; i < n; ++i) {
        coff[i] = off[i] - lo;
        slot[i + 1] = slot[i] + (((uint64_t) rec_cap[i] + 15u) & ~15ull);
    }
    int rc;
    Trace tr(ctx, "decompress_records");
    if ((rc = reserve(ctx, ctx->bulk_in, (size_t) (hi - lo) + 64, false))) return rc;
    if ((rc = reserve(ctx, ctx->bulk_out, (size_t) slot[n] + 64, false))) return rc;
    tr.mark("reserve");
    if (zwz_rt::memcpy_h2d(ctx->bulk_in.p, comp + lo, (size_t) (hi - lo), ctx->stream)) return fail(ctx, ZWZ_E_CUDA, "compressed upload failed");
    tr.mark("h2d");
    // per-record capacity is rec_cap, but slots are 16-byte aligned so the gather can use 128-bit copies
    std::vector<uint64_t> cap_off(n + 1);
    for (uint32_t i = 0; i <= n; ++i) cap_off[i] = slot[i];
    // inflate_kernel takes capacity = raw_off[i+1]-raw_off[i]; the up-to-15 padding bytes per slot are harmless extra room,
    // but OUTPUT_FULL must be judged against rec_cap — checked below.
    if ((rc = zwz_inflate_batch_device(ctx, (const uint8_t *) ctx->bulk_in.p, coff.data(), len, n, (uint8_t *) ctx->bulk_out.p, cap_off.data(),
                                       raw_len, status, flags, nullptr)))
        return rc;
    tr.mark("inflate");
    bool retry = false;
    for (uint32_t i = 0; i < n; ++i) {
        if (status[i] == ZWZ_STREAM_OUTPUT_FULL || raw_len[i] > rec_cap[i]) {
            if (raw_len[i] > rec_cap[i]) {
                status[i] = ZWZ_STREAM_OUTPUT_FULL;
                retry = true;
            }
        }
    }
    if (retry) return ZWZ_OK; // caller inspects status/raw_len and calls again with larger capacities
    // layout of the concatenated files
    std::vector<uint64_t> dst(n);
    uint64_t acc = 0;
    uint32_t f = 0;
    for (uint32_t i = 0; i < n; ++i) {
        while (f < rec_file[i]) file_off_out[++f] = acc;
        dst[i] = acc;
        acc += raw_len[i];
    }
    while (f < nf) file_off_out[++f] = acc;
    if (acc > out_cap) return fail(ctx, ZWZ_E_CAPACITY, "output buffer too small");
    const size_t m_dst = (size_t) n * 8, m_len = 2 * (size_t) n * 8, meta_bytes = m_len + (size_t) n * 4;
    // the concatenated files live in a digest slot: their MD5 runs on the second stream and may outlive this call
    zwz::DigestSlot *sl = nullptr;
    if ((rc = slot_acquire(ctx, &sl))) return rc;
    if ((rc = reserve(ctx, sl->data, (size_t) acc + 256, false))) return rc;
    if ((rc = reserve(ctx, ctx->packed, align_up(meta_bytes, 256) + 256, false))) return rc;
    // NOT pin_meta: other calls refill pin_meta while this upload may still be in flight (page-locked memory is read by
    // the DMA engine after cudaMemcpyAsync returns) — a reuse race that corrupted gather descriptors once per ~10^5 records.
    if ((rc = reserve(ctx, ctx->pin_aux, meta_bytes, true))) return rc;
    tr.mark("reserve2");
    uint8_t *hp = (uint8_t *) ctx->pin_aux.p;
    memcpy(hp, slot.data(), (size_t) n * 8);
    memcpy(hp + m_dst, dst.data(), (size_t) n * 8);
    memcpy(hp + m_len, raw_len, (size_t) n * 4);
    uint8_t *dmeta = (uint8_t *) ctx->packed.p;
    uint8_t *d_files = (uint8_t *) sl->data.p;
    if (zwz_rt::memcpy_h2d(dmeta, hp, meta_bytes, ctx->stream)) return fail(ctx, ZWZ_E_CUDA, "descriptor upload failed");
    {
        ProfSpan ps(ctx, ZWZ_PROF_PACK, ctx->stream);
        ZWZ_LAUNCH(zwz::gather_records_kernel, (n + 7) / 8, 256, 0, ctx->stream, (const uint8_t *) ctx->bulk_out.p, (const uint64_t *) dmeta,
                   (const uint64_t *) (dmeta + m_dst), (const uint32_t *) (dmeta + m_len), d_files, n);
    }
    if ((rc = check_launch(ctx, "gather_records_kernel"))) return rc;
    tr.mark("gather");
    if (digest && nf) { // beside the download of the files
        std::vector<uint64_t> fo(nf), fl(nf);
        for (uint32_t i = 0; i < nf; ++i) {
            fo[i] = file_off_out[i];
            fl[i] = file_off_out[i + 1] - file_off_out[i];
        }
        if ((rc = slot_queue_md5(ctx, *sl, fo.data(), fl.data(), nf, digest, ticket))) return rc;
    }
    tr.mark("md5 queued");
    if (zwz_rt::memcpy_d2h(files_out, d_files, (size_t) acc, ctx->stream) || zwz_rt::stream_sync(ctx->stream)) {
        sl->digest_dst = nullptr;
        return fail(ctx, ZWZ_E_CUDA, "raw download failed");
    }
    tr.mark("d2h");
    if (!ticket) return slot_finish(ctx, *sl);
    return ZWZ_OK;
}

int zwz_decompress_records(zwz_ctx *ctx, const uint8_t *comp, const uint64_t *off, const uint32_t *len, const uint32_t *rec_cap,
                           const uint32_t *rec_file, uint32_t n, uint32_t nf, uint8_t *files_out, uint64_t out_cap, uint64_t *file_off_out,
                           uint32_t *raw_len, uint32_t *status, uint8_t *digest, uint32_t flags) {
    return decompress_records_impl(ctx, comp, off, len, rec_cap, rec_file, n, nf, files_out, out_cap, file_off_out, raw_len, status, digest, flags,
                                   nullptr);
}

int zwz_decompress_records_async(zwz_ctx *ctx, const uint8_t *comp, const uint64_t *off, const uint32_t *len, const uint32_t *rec_cap,
                                 const uint32_t *rec_file, uint32_t n, uint32_t nf, uint8_t *files_out, uint64_t out_cap, uint64_t *file_off_out,
                                 uint32_t *raw_len, uint32_t *status, uint8_t *digest, uint32_t flags, uint64_t *ticket) {
    if (!ticket) return ctx ? fail(ctx, ZWZ_E_ARG, "null ticket") : ZWZ_E_ARG;
    return decompress_records_impl(ctx, comp, off, len, rec_cap, rec_file, n, nf, files_out, out_cap, file_off_out, raw_len, status, digest, flags,
                                   ticket);
}

// ======================================================================================================================
// MD5 / Adler-32
// ======================================================================================================================
// Queues the MD5 kernel over n files on `st`: descriptors go up from `pin`, digests (and chained states) come back into `pin`
// at *digest_off (resp. m_state). Nothing is waited for.
static int md5_enqueue(zwz_ctx *ctx, zwz_stream_t st, Arena &dev_meta, Arena &pin, uint32_t *state, const uint8_t *d_data, const uint64_t *off,
                       const uint64_t *len, const uint64_t *total_len, uint32_t n, int finalize, size_t *digest_off) {
    const size_t m_len = (size_t) n * 8, m_tot = 2 * (size_t) n * 8, m_state = 3 * (size_t) n * 8, meta_bytes = m_state + (size_t) n * 16;
    const size_t r_base = align_up(meta_bytes, 256);
    int rc;
    if ((rc = reserve(ctx, pin, r_base + (size_t) n * 16, true))) return rc;
    if ((rc = reserve(ctx, dev_meta, r_base + (size_t) n * 16 + 256, false))) return rc;
    uint8_t *hp = (uint8_t *) pin.p;
    memcpy(hp, off, (size_t) n * 8);
    memcpy(hp + m_len, len, (size_t) n * 8);
    if (total_len) memcpy(hp + m_tot, total_len, (size_t) n * 8);
    if (state) memcpy(hp + m_state, state, (size_t) n * 16);
    uint8_t *dm = (uint8_t *) dev_meta.p;
    if (zwz_rt::memcpy_h2d(dm, hp, meta_bytes, st)) return fail(ctx, ZWZ_E_CUDA, "descriptor upload failed");
    uint8_t *d_digest = dm + r_base;
    {
        // few long files: the staged kernel (coalesced cp.async into shared memory); many short ones: one lane streams its file
        uint64_t total = 0;
        for (uint32_t i = 0; i < n; ++i) total += len[i];
        const bool staged = total / n >= 65536u;
        ProfSpan ps(ctx, ZWZ_PROF_MD5, st);
        if (staged) {
            ZWZ_LAUNCH(zwz::md5_files_staged_kernel, (n + 31u) / 32u, ZWZ_MD5S_THREADS, ZWZ_MD5S_SMEM, st, d_data, (const uint64_t *) dm,
                       (const uint64_t *) (dm + m_len), total_len ? (const uint64_t *) (dm + m_tot) : (const uint64_t *) nullptr,
                       state ? (uint32_t *) (dm + m_state) : (uint32_t *) nullptr, d_digest, n, finalize);
        } else {
            ZWZ_LAUNCH(zwz::md5_files_kernel, (n + 127) / 128, 128, 0, st, d_data, (const uint64_t *) dm, (const uint64_t *) (dm + m_len),
                       total_len ? (const uint64_t *) (dm + m_tot) : (const uint64_t *) nullptr,
                       state ? (uint32_t *) (dm + m_state) : (uint32_t *) nullptr, d_digest, n, finalize);
        }
    }
    if ((rc = check_launch(ctx, "md5_files_kernel"))) return rc;
    if (finalize && zwz_rt::memcpy_d2h(hp + r_base, d_digest, (size_t) n * 16, st)) return fail(ctx, ZWZ_E_CUDA, "digest download failed");
    if (state && zwz_rt::memcpy_d2h(hp + m_state, dm + m_state, (size_t) n * 16, st)) return fail(ctx, ZWZ_E_CUDA, "state download failed");
    *digest_off = r_base;
    return ZWZ_OK;
}

static int md5_launch(zwz_ctx *ctx, uint32_t *state, const uint8_t *d_data, const uint64_t *off, const uint64_t *len, const uint64_t *total_len,
                      uint32_t n, uint8_t *digest, int finalize, void *stream_v) {
    if (!ctx) return ZWZ_E_ARG;
    if (n == 0) return ZWZ_OK;
    if (!off || !len || (finalize && !digest)) return fail(ctx, ZWZ_E_ARG, "null argument");
    zwz_rt::set_device(ctx->device);
    zwz_stream_t st = stream_v ? (zwz_stream_t) stream_v : ctx->stream;
    size_t doff = 0;
    int rc = md5_enqueue(ctx, st, ctx->meta, ctx->pin_meta, state, d_data, off, len, total_len, n, finalize, &doff);
    if (rc) return rc;
    if (zwz_rt::stream_sync(st)) return fail(ctx, ZWZ_E_CUDA, "md5 kernel failed");
    const uint8_t *hp = (const uint8_t *) ctx->pin_meta.p;
    if (finalize) memcpy(digest, hp + doff, (size_t) n * 16);
    if (state) memcpy(state, hp + 3 * (size_t) n * 8, (size_t) n * 16);
    return ZWZ_OK;
}

// ---- deferred digests (DigestSlot) ---------------------------------------------------------------------------------
// the slot a new batch may use: its previous batch's digests are delivered first if nobody has waited for them yet
static int slot_finish(zwz_ctx *ctx, zwz::DigestSlot &sl) {
    if (!sl.pending) return ZWZ_OK;
    sl.pending = false;
    if (zwz_rt::sync_event_wait(sl.ev_done)) return fail(ctx, ZWZ_E_CUDA, "md5 kernel failed");
    if (sl.digest_dst && sl.n) memcpy(sl.digest_dst, (const uint8_t *) sl.pin.p + sl.digest_off, (size_t) sl.n * 16);
    return ZWZ_OK;
}
static int slot_acquire(zwz_ctx *ctx, zwz::DigestSlot **out) {
    zwz::DigestSlot &sl = ctx->slot[ctx->next_slot];
    ctx->next_slot ^= 1;
    int rc = slot_finish(ctx, sl);
    *out = &sl;
    return rc;
}
// MD5 of n files inside sl.data, queued on the second stream behind everything the main stream has queued so far
static int slot_queue_md5(zwz_ctx *ctx, zwz::DigestSlot &sl, const uint64_t *off, const uint64_t *len, uint32_t n, uint8_t *digest, uint64_t *ticket) {
    if (zwz_rt::sync_event_record(sl.ev_in, ctx->stream) || zwz_rt::stream_wait_event(ctx->md5_stream, sl.ev_in))
        return fail(ctx, ZWZ_E_CUDA, "stream dependency failed");
    int rc = md5_enqueue(ctx, ctx->md5_stream, sl.meta, sl.pin, nullptr, (const uint8_t *) sl.data.p, off, len, nullptr, n, 1, &sl.digest_off);
    if (rc) return rc;
    if (zwz_rt::sync_event_record(sl.ev_done, ctx->md5_stream)) return fail(ctx, ZWZ_E_CUDA, "event record failed");
    sl.pending = true;
    sl.n = n;
    sl.digest_dst = digest;
    sl.ticket = ++ctx->ticket_seq;
    if (ticket) *ticket = sl.ticket;
    return ZWZ_OK;
}

int zwz_md5_batch_device(zwz_ctx *ctx, const uint8_t *d_data, const uint64_t *off, const uint64_t *len, uint32_t n, uint8_t *digest, void *stream) {
    return md5_launch(ctx, nullptr, d_data, off, len, nullptr, n, digest, 1, stream);
}

int zwz_md5_batch(zwz_ctx *ctx, const uint8_t *data, const uint64_t *off, const uint64_t *len, uint32_t n, uint8_t *digest) {
    if (!ctx) return ZWZ_E_ARG;
    if (n == 0) return ZWZ_OK;
    if (!off || !len || !digest) return fail(ctx, ZWZ_E_ARG, "null argument");
    zwz_rt::set_device(ctx->device);
    uint64_t lo = ~0ull, hi = 0;
    for (uint32_t i = 0; i < n; ++i) {
        lo = std::min(lo, off[i]);
        hi = std::max(hi, off[i] + len[i]);
    }
    if (hi < lo) lo = hi = 0;
    std::vector<uint64_t> o(n);
    for (uint32_t i = 0; i < n; ++i) o[i] = off[i] - lo;
    int rc;
    if ((rc = reserve(ctx, ctx->bulk_in, (size_t) (hi - lo) + 64, false))) return rc;
    if (zwz_rt::memcpy_h2d(ctx->bulk_in.p, data + lo, (size_t) (hi - lo), ctx->stream)) return fail(ctx, ZWZ_E_CUDA, "data upload failed");
    return md5_launch(ctx, nullptr, (const uint8_t *) ctx->bulk_in.p, o.data(), len, nullptr, n, digest, 1, nullptr);
}

void zwz_md5_state_init(uint32_t *state, uint32_t n) {
    for (uint32_t i = 0; i < n; ++i) {
        state[4 * i + 0] = 0x67452301u;
        state[4 * i + 1] = 0xefcdab89u;
        state[4 * i + 2] = 0x98badcfeu;
        state[4 * i + 3] = 0x10325476u;
    }
}

int zwz_md5_update_device(zwz_ctx *ctx, uint32_t *state, const uint8_t *d_data, const uint64_t *off, const uint64_t *len, uint32_t n, void *stream) {
    if (!ctx) return ZWZ_E_ARG;
    if (!state) return fail(ctx, ZWZ_E_ARG, "null state");
    for (uint32_t i = 0; i < n; ++i)
        if (len[i] & 63u) return fail(ctx, ZWZ_E_ARG, "md5 update length must be a multiple of 64");
    return md5_launch(ctx, state, d_data, off, len, nullptr, n, nullptr, 0, stream);
}

int zwz_md5_final_device(zwz_ctx *ctx, uint32_t *state, const uint8_t *d_tail, const uint64_t *off, const uint64_t *len, const uint64_t *total_len,
                         uint32_t n, uint8_t *digest, void *stream) {
    if (!ctx) return ZWZ_E_ARG;
    if (!state || !total_len) return fail(ctx, ZWZ_E_ARG, "null argument");
    return md5_launch(ctx, state, d_tail, off, len, total_len, n, digest, 1, stream);
}

void zwz_md5_hex(const uint8_t digest[16], char hex[32]) {
    static const char d[] = "0123456789abcdef";
    for (int i = 0; i < 16; ++i) {
        hex[2 * i] = d[digest[i] >> 4];
        hex[2 * i + 1] = d[digest[i] & 15];
    }
}

int zwz_adler32_batch_device(zwz_ctx *ctx, const uint8_t *d_data, const uint64_t *off, const uint32_t *len, uint32_t n, uint32_t *adler, void *stream_v) {
    if (!ctx) return ZWZ_E_ARG;
    if (n == 0) return ZWZ_OK;
    if (!off || !len || !adler) return fail(ctx, ZWZ_E_ARG, "null argument");
    zwz_rt::set_device(ctx->device);
    zwz_stream_t st = stream_v ? (zwz_stream_t) stream_v : ctx->stream;
    const size_t m_len = (size_t) n * 8, meta_bytes = m_len + (size_t) n * 4, r_base = align_up(meta_bytes, 256);
    int rc;
    if ((rc = reserve(ctx, ctx->pin_meta, meta_bytes, true))) return rc;
    if ((rc = reserve(ctx, ctx->meta, r_base + (size_t) n * 4 + 256, false))) return rc;
    uint8_t *hp = (uint8_t *) ctx->pin_meta.p;
    memcpy(hp, off, (size_t) n * 8);
    memcpy(hp + m_len, len, (size_t) n * 4);
    uint8_t *dm = (uint8_t *) ctx->meta.p;
    if (zwz_rt::memcpy_h2d(dm, hp, meta_bytes, st)) return fail(ctx, ZWZ_E_CUDA, "descriptor upload failed");
    {
        ProfSpan ps(ctx, ZWZ_PROF_ADLER, st);
        ZWZ_LAUNCH(zwz::adler32_kernel, (n + 7) / 8, 256, 0, st, d_data, (const uint64_t *) dm, (const uint32_t *) (dm + m_len),
                   (uint32_t *) (dm + r_base), n);
    }
    if ((rc = check_launch(ctx, "adler32_kernel"))) return rc;
    if (zwz_rt::memcpy_d2h(hp, dm + r_base, (size_t) n * 4, st) || zwz_rt::stream_sync(st)) return fail(ctx, ZWZ_E_CUDA, "adler kernel failed");
    memcpy(adler, hp, (size_t) n * 4);
    return ZWZ_OK;
}

} // extern "C"
