// simt_emu.h — a small SIMT emulator so the CUDA kernel SOURCES under csrc/ can be executed on a CPU-only box.
//
// TEST INFRASTRUCTURE ONLY. The build container has nvcc but no GPU and a `gpurun` round-trip costs minutes of a
// small budget, so the kernels are first run here (g++ -DZWZ_EMU, tests/simt/Makefile -> tests/simt/libzwz_emu.so)
// to shake out indexing / protocol bugs. The product package never loads this library: it dlopens libzwz_cuda.so
// only and raises if that is missing. Nothing measured or shipped goes through this file.
//
// Model: one CTA at a time; every CUDA thread is a fiber (own stack, hand-written x86-64 context switch); warp
// collectives (__shfl*_sync, __ballot_sync, __match_any_sync, __syncwarp) and __syncthreads() are rendezvous points
// at which a fiber parks until all participants arrived. Single OS thread => atomics are plain ops and data races are
// NOT detected; what is detected: wrong answers, out-of-bounds (when built with the guard allocator), mismatched
// collectives, deadlocks.
#pragma once
#if !defined(__x86_64__)
#error "simt_emu.h: x86-64 only"
#endif
#include <algorithm>
#include <cassert>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <vector>

namespace simt {

struct uint3_ {
    unsigned x, y, z;
};

extern "C" void zwz_simt_switch(void **save_sp, void *load_sp);

struct Fiber {
    void *sp = nullptr;
    char *stack = nullptr;
    unsigned tid = 0;
    bool done = false;
};

struct WarpSlot {
    uint32_t mask = 0, arrived = 0, departed = 0;
    int op = 0;
    bool ready = false;
    uint64_t val[32];
    uint64_t aux[32];
};

struct State {
    std::vector<Fiber> fibers;
    std::vector<WarpSlot> warps;
    std::vector<unsigned> cursor; // per warp: lane to resume next when control comes in from another warp
    std::function<void()> body;
    void *sched_sp = nullptr;
    unsigned cur = 0, nthreads = 0;
    unsigned exited = 0, bar_count = 0;
    uint64_t bar_gen = 0;
    uint64_t stall = 0; // consecutive blocked yields without a progress event
    unsigned char *dyn = nullptr;
    size_t dyn_bytes = 0;
    static constexpr size_t kStack = 96 * 1024;
};

inline State &st() {
    static State s;
    return s;
}

// CUDA built-in variables
inline uint3_ threadIdx, blockIdx, blockDim, gridDim;

inline void progress() { st().stall = 0; }

[[noreturn]] inline void die(const char *msg) {
    State &s = st();
    fprintf(stderr, "simt_emu: %s (block %u thread %u)\n", msg, blockIdx.x, s.cur);
    abort();
}

inline void switch_to(unsigned next) {
    State &s = st();
    unsigned me = s.cur;
    if (next == me) return;
    s.cur = next;
    threadIdx.x = s.fibers[next].tid;
    zwz_simt_switch(&s.fibers[me].sp, s.fibers[next].sp);
    // resumed
    threadIdx.x = s.fibers[s.cur].tid;
}

inline void note_stall() {
    State &s = st();
    if (++s.stall > 400000000ull) die("deadlock: no progress event for 4e8 blocked yields");
}

// yield to the next live fiber of the same warp (cyclic); if none other is live, to the next warp
inline void yield_in_warp() {
    State &s = st();
    note_stall();
    unsigned me = s.cur, w0 = me & ~31u, n = s.nthreads;
    for (unsigned k = 1; k < 32; ++k) {
        unsigned c = w0 + ((me - w0 + k) & 31u);
        if (c < n && !s.fibers[c].done) {
            switch_to(c);
            return;
        }
    }
}
// yield to a live fiber of the next warp that has one (cyclic over the CTA); inside the target warp the entry lane
// rotates so that every lane gets to run
inline void yield_next_warp() {
    State &s = st();
    note_stall();
    unsigned me = s.cur, n = s.nthreads, nw = (n + 31u) / 32u;
    unsigned w0 = me >> 5;
    for (unsigned dw = 1; dw <= nw; ++dw) {
        unsigned w = (w0 + dw) % nw;
        unsigned cur = s.cursor[w];
        for (unsigned k = 0; k < 32; ++k) {
            unsigned l = (cur + k) & 31u;
            unsigned c = w * 32u + l;
            if (c >= n || c == me || s.fibers[c].done) continue;
            s.cursor[w] = (l + 1u) & 31u;
            switch_to(c);
            return;
        }
    }
}

inline void fiber_exit() {
    State &s = st();
    s.fibers[s.cur].done = true;
    s.exited++;
    progress();
    if (s.bar_count && s.bar_count + s.exited == s.nthreads) { // Volta+: exited threads count as arrived
        s.bar_count = 0;
        s.bar_gen++;
    }
    // switch away for good: next live fiber, else scheduler
    unsigned n = s.nthreads, me = s.cur;
    for (unsigned k = 1; k < n; ++k) {
        unsigned c = (me + k) % n;
        if (!s.fibers[c].done) {
            s.cur = c;
            threadIdx.x = s.fibers[c].tid;
            zwz_simt_switch(&s.fibers[me].sp, s.fibers[c].sp);
            die("resumed a finished fiber");
        }
    }
    zwz_simt_switch(&s.fibers[me].sp, s.sched_sp);
    die("resumed a finished fiber");
}

extern "C" inline void zwz_simt_entry() {
    st().body();
    fiber_exit();
}

inline unsigned char *dyn_smem() { return st().dyn; }

void launch(unsigned grid, unsigned block, size_t smem_bytes, std::function<void()> body);

#ifdef ZWZ_SIMT_IMPL
asm(R"(
.text
.globl zwz_simt_switch
.type zwz_simt_switch,@function
zwz_simt_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
.size zwz_simt_switch,.-zwz_simt_switch
)");

void launch(unsigned grid, unsigned block, size_t smem_bytes, std::function<void()> body) {
    // one kernel at a time: the emulator's state (fibers, threadIdx, `__shared__` statics) is process-wide, and a host with
    // several worker threads (each with its own zwz_ctx) may launch concurrently
    static std::mutex launch_mu;
    std::lock_guard<std::mutex> launch_lock(launch_mu);
    State &s = st();
    if (block == 0 || block > 1024) die("bad block size");
    if (s.fibers.size() < block) {
        size_t old = s.fibers.size();
        s.fibers.resize(block);
        for (size_t i = old; i < block; ++i) s.fibers[i].stack = (char *) aligned_alloc(64, State::kStack);
    }
    s.warps.assign((block + 31) / 32, WarpSlot());
    s.cursor.assign((block + 31) / 32, 0u);
    s.body = std::move(body);
    s.nthreads = block;
    // dynamic smem with a poisoned guard zone after it
    free(s.dyn);
    s.dyn_bytes = smem_bytes;
    s.dyn = (unsigned char *) aligned_alloc(1024, ((smem_bytes + 1023) / 1024 + 1) * 1024);
    blockDim = {block, 1, 1};
    gridDim = {grid, 1, 1};
    for (unsigned b = 0; b < grid; ++b) {
        blockIdx = {b, 0, 0};
        memset(s.dyn, 0xA5, smem_bytes + 1024); // smem is garbage at CTA start on real hardware too
        // ZWZ_EMU_POISON=<seed>: garbage that LOOKS like data (small 16-bit values, as the previous kernel's position lists
        // leave behind on the device) — 0xA5A5 is out of range for most indices and hides reads of unset state
        if (const char *ps = getenv("ZWZ_EMU_POISON")) {
            static uint64_t launches = 0;
            uint64_t x = (strtoull(ps, nullptr, 0) + 1000u * ++launches) * 0x9E3779B97F4A7C15ull + b + 1;
            uint16_t *w = (uint16_t *) s.dyn;
            for (size_t k = 0; k < smem_bytes / 2; ++k) {
                x ^= x << 13, x ^= x >> 7, x ^= x << 17;
                w[k] = (uint16_t) ((x >> 20) % 6000u);
            }
        }
        s.exited = 0;
        s.bar_count = 0;
        s.stall = 0;
        for (auto &w : s.warps) w = WarpSlot();
        for (unsigned t = 0; t < block; ++t) {
            Fiber &f = s.fibers[t];
            f.tid = t;
            f.done = false;
            uintptr_t top = ((uintptr_t) f.stack + State::kStack) & ~(uintptr_t) 15;
            void **sp = (void **) top;
            *--sp = nullptr;                    // fake return address of the entry function
            *--sp = (void *) &zwz_simt_entry;   // `ret` in zwz_simt_switch jumps here
            for (int k = 0; k < 6; ++k) *--sp = nullptr;
            f.sp = sp;
        }
        s.cur = 0;
        threadIdx = {0, 0, 0};
        zwz_simt_switch(&s.sched_sp, s.fibers[0].sp);
        if (s.exited != block) die("scheduler resumed with live fibers");
        for (unsigned k = 0; k < 1024; ++k)
            if (s.dyn[smem_bytes + k] != 0xA5) die("dynamic shared memory overrun detected (guard zone modified)");
    }
}
#endif // ZWZ_SIMT_IMPL

// ------------------------------------------------------------------------------------------------
// rendezvous
// ------------------------------------------------------------------------------------------------
inline void syncthreads_() {
    State &s = st();
    uint64_t gen = s.bar_gen;
    s.bar_count++;
    if (s.bar_count + s.exited == s.nthreads) {
        s.bar_count = 0;
        s.bar_gen++;
        progress();
        return;
    }
    while (s.bar_gen == gen) yield_next_warp();
}

enum { OP_SYNC = 1, OP_SHFL, OP_SHFL_UP, OP_SHFL_DOWN, OP_SHFL_XOR, OP_BALLOT, OP_MATCH, OP_ANY, OP_ALL };

template <class F> inline uint64_t collective(uint32_t mask, int op, uint64_t v, uint64_t a, F compute) {
    State &s = st();
    unsigned lane = s.cur & 31u;
    WarpSlot &W = s.warps[s.cur >> 5];
    uint32_t bit = 1u << lane;
    if (!(mask & bit)) die("collective: calling lane not in mask");
    // lanes of the CTA's last, partial warp that do not exist can never arrive
    unsigned wbase = s.cur & ~31u;
    uint32_t exist = (s.nthreads - wbase >= 32) ? 0xffffffffu : ((1u << (s.nthreads - wbase)) - 1u);
    mask &= exist;
    int spins = 0;
    while (W.ready || (W.arrived && (W.mask != mask || (W.arrived & bit)))) {
        if (++spins > 64) yield_next_warp(); else yield_in_warp();
    }
    if (!W.arrived) {
        W.mask = mask;
        W.op = op;
        W.departed = 0;
    } else if (W.op != op) {
        die("collective: lanes of one warp met in different primitives");
    }
    W.val[lane] = v;
    W.aux[lane] = a;
    W.arrived |= bit;
    if (W.arrived == mask) {
        W.ready = true;
        progress();
    } else {
        spins = 0;
        while (!W.ready) {
            // after one fruitless lap over the warp the missing lanes are parked elsewhere: let other warps run
            if (++spins > 40) yield_next_warp(); else yield_in_warp();
        }
    }
    uint64_t r = compute(W, lane);
    W.departed |= bit;
    if (W.departed == W.mask) {
        W.arrived = 0;
        W.departed = 0;
        W.ready = false;
        progress();
    }
    return r;
}

template <class T> inline uint64_t to_u64(T v) {
    static_assert(sizeof(T) <= 8, "shuffle payload too wide");
    uint64_t u = 0;
    memcpy(&u, &v, sizeof(T));
    return u;
}
template <class T> inline T from_u64(uint64_t u) {
    T v;
    memcpy(&v, &u, sizeof(T));
    return v;
}

} // namespace simt

// ------------------------------------------------------------------------------------------------
// CUDA surface
// ------------------------------------------------------------------------------------------------
using simt::blockDim;
using simt::blockIdx;
using simt::gridDim;
using simt::threadIdx;

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(n) __attribute__((aligned(n)))
#define __constant__ static const
#define warpSize 32

inline void __syncthreads() { simt::syncthreads_(); }
inline void __syncwarp(unsigned mask = 0xffffffffu) {
    simt::collective(mask, simt::OP_SYNC, 0, 0, [](simt::WarpSlot &, unsigned) { return (uint64_t) 0; });
}
template <class T> inline T __shfl_sync(unsigned mask, T v, int src, int width = 32) {
    return simt::from_u64<T>(simt::collective(mask, simt::OP_SHFL, simt::to_u64(v), (uint64_t) (unsigned) src, [width](simt::WarpSlot &W, unsigned lane) {
        unsigned base = lane & ~(unsigned) (width - 1);
        unsigned s = base + ((unsigned) W.aux[lane] & (unsigned) (width - 1));
        if (!(W.mask >> s & 1u)) return W.val[lane]; // undefined on hardware; keep own value
        return W.val[s];
    }));
}
template <class T> inline T __shfl_up_sync(unsigned mask, T v, unsigned delta, int width = 32) {
    return simt::from_u64<T>(simt::collective(mask, simt::OP_SHFL_UP, simt::to_u64(v), delta, [width](simt::WarpSlot &W, unsigned lane) {
        unsigned base = lane & ~(unsigned) (width - 1);
        int s = (int) lane - (int) W.aux[lane];
        if (s < (int) base || !(W.mask >> s & 1u)) return W.val[lane];
        return W.val[s];
    }));
}
template <class T> inline T __shfl_down_sync(unsigned mask, T v, unsigned delta, int width = 32) {
    return simt::from_u64<T>(simt::collective(mask, simt::OP_SHFL_DOWN, simt::to_u64(v), delta, [width](simt::WarpSlot &W, unsigned lane) {
        unsigned base = lane & ~(unsigned) (width - 1);
        unsigned s = lane + (unsigned) W.aux[lane];
        if (s >= base + (unsigned) width || !(W.mask >> s & 1u)) return W.val[lane];
        return W.val[s];
    }));
}
template <class T> inline T __shfl_xor_sync(unsigned mask, T v, int lanemask, int width = 32) {
    return simt::from_u64<T>(simt::collective(mask, simt::OP_SHFL_XOR, simt::to_u64(v), (uint64_t) (unsigned) lanemask, [width](simt::WarpSlot &W, unsigned lane) {
        unsigned s = lane ^ (unsigned) W.aux[lane];
        if ((s & ~(unsigned) (width - 1)) != (lane & ~(unsigned) (width - 1)) || !(W.mask >> s & 1u)) return W.val[lane];
        return W.val[s];
    }));
}
inline unsigned __ballot_sync(unsigned mask, int pred) {
    return (unsigned) simt::collective(mask, simt::OP_BALLOT, pred ? 1 : 0, 0, [](simt::WarpSlot &W, unsigned) {
        uint64_t r = 0;
        for (unsigned l = 0; l < 32; ++l)
            if ((W.mask >> l & 1u) && W.val[l]) r |= 1ull << l;
        return r;
    });
}
inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0; }
inline int __all_sync(unsigned mask, int pred) {
    return (int) simt::collective(mask, simt::OP_ALL, pred ? 1 : 0, 0, [](simt::WarpSlot &W, unsigned) {
        for (unsigned l = 0; l < 32; ++l)
            if ((W.mask >> l & 1u) && !W.val[l]) return (uint64_t) 0;
        return (uint64_t) 1;
    });
}
template <class T> inline unsigned __match_any_sync(unsigned mask, T v) {
    return (unsigned) simt::collective(mask, simt::OP_MATCH, simt::to_u64(v), 0, [](simt::WarpSlot &W, unsigned lane) {
        uint64_t r = 0;
        for (unsigned l = 0; l < 32; ++l)
            if ((W.mask >> l & 1u) && W.val[l] == W.val[lane]) r |= 1ull << l;
        return r;
    });
}
inline unsigned __activemask() { simt::die("__activemask is not supported by the emulator: pass explicit masks"); }

// spin-wait hint used by kernels that poll shared memory written by another warp
inline void zwz_emu_spin_pause() { simt::yield_next_warp(); }
inline void __nanosleep(unsigned) { simt::yield_next_warp(); }
inline void __threadfence() {}
inline void __threadfence_block() {}
inline void __trap() { simt::die("__trap()"); }

// bit intrinsics
inline int __popc(unsigned x) { return __builtin_popcount(x); }
inline int __popcll(unsigned long long x) { return __builtin_popcountll(x); }
inline int __clz(int x) { return x ? __builtin_clz((unsigned) x) : 32; }
inline int __clzll(long long x) { return x ? __builtin_clzll((unsigned long long) x) : 64; }
inline int __ffs(int x) { return __builtin_ffs(x); }
inline int __ffsll(long long x) { return __builtin_ffsll(x); }
inline unsigned __brev(unsigned x) {
    x = ((x >> 1) & 0x55555555u) | ((x & 0x55555555u) << 1);
    x = ((x >> 2) & 0x33333333u) | ((x & 0x33333333u) << 2);
    x = ((x >> 4) & 0x0f0f0f0fu) | ((x & 0x0f0f0f0fu) << 4);
    return __builtin_bswap32(x);
}
inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned sh) {
    uint64_t v = ((uint64_t) hi << 32) | lo;
    return (unsigned) (v >> (sh & 31u));
}
inline unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned sh) {
    uint64_t v = ((uint64_t) hi << 32) | lo;
    return (unsigned) ((v << (sh & 31u)) >> 32);
}
inline unsigned __byte_perm(unsigned a, unsigned b, unsigned sel) {
    uint64_t v = ((uint64_t) b << 32) | a;
    unsigned r = 0;
    for (int i = 0; i < 4; ++i) {
        unsigned s = (sel >> (4 * i)) & 0xf;
        unsigned byte = (unsigned) (v >> (8 * (s & 7))) & 0xff;
        if (s & 8) byte = (byte & 0x80) ? 0xff : 0x00;
        r |= byte << (8 * i);
    }
    return r;
}
inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned) (((uint64_t) a * b) >> 32); }
template <class T> inline T __ldg(const T *p) { return *p; }
template <class T> inline T __ldcg(const T *p) { return *p; }
template <class T> inline T __ldcs(const T *p) { return *p; }
template <class T> inline void __stcg(T *p, T v) { *p = v; }
template <class T> inline void __stcs(T *p, T v) { *p = v; }

// atomics (single OS thread: plain read-modify-write)
template <class T> inline T atomicAdd(T *p, T v) { T o = *p; *p = o + v; return o; }
template <class T> inline T atomicSub(T *p, T v) { T o = *p; *p = o - v; return o; }
template <class T> inline T atomicMax(T *p, T v) { T o = *p; if (v > o) *p = v; return o; }
template <class T> inline T atomicMin(T *p, T v) { T o = *p; if (v < o) *p = v; return o; }
template <class T> inline T atomicOr(T *p, T v) { T o = *p; *p = o | v; return o; }
template <class T> inline T atomicAnd(T *p, T v) { T o = *p; *p = o & v; return o; }
template <class T> inline T atomicExch(T *p, T v) { T o = *p; *p = v; return o; }
template <class T> inline T atomicCAS(T *p, T c, T v) { T o = *p; if (o == c) *p = v; return o; }

using std::max;
using std::min;

struct uint4 { unsigned x, y, z, w; };
struct uint2 { unsigned x, y; };
inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
