// deflate_match.cuh — LZ77 match finding for <= 65 535-byte chunks, everything in shared memory.
//
// First half of the replacement for zlib's deflate() at compression.cpp:119-134 (the second half is deflate_encode.cuh).
//
// Per chunk (persistent CTAs, chunks handed out by an atomic counter):
//   1. the chunk is staged HBM -> shared memory with one bulk asynchronous copy (cp.async.bulk + mbarrier, SASS UBLKCP),
//      a 16-byte aligned superset of the chunk; the chunk itself starts at byte `skew` of the staging buffer;
//   2. all threads hash every position (3-byte multiplicative hash) into prev[] in parallel;
//   3. warp 0 turns the hashes into EXACT hash chains, 32 positions per step: every lane links to the old head and
//      becomes the new head; lanes sharing a hash inside the step are found by reading the head back and repaired with
//      ballot + shuffle (lowest lane keeps the old head, the others link to the nearest earlier lane, the highest stays
//      head). Chains are final behind the build front, which is published through a shared-memory counter;
//   4. the other warps (and warp 0 once it is done) pull 32-position tiles and, one lane per position, walk the chain:
//      byte checks around the current best length reject most candidates, survivors are compared 4 bytes at a time on
//      funnel-shifted aligned words; `depth` candidates at most, stop at `nice` bytes;
//   5. the best (length, distance) of EVERY position goes to the chunk's scratch in HBM (4 bytes per input byte,
//      coalesced); the parse in deflate_encode.cuh picks the path through them;
//   6. Adler-32 of the chunk is reduced from shared memory while it is there.
//
// Three size classes so that small chunks do not leave an SM to one serial chain builder (config C2 is 370 000 chunks of
// ~6.7 KB): the class fixes the shared-memory footprint and with it how many CTAs — each with its own builder warp — an SM
// holds.
//      class   chunk bytes   threads   smem/CTA   CTAs/SM   hash bits
//        S       <=  8 192      256     ~33 KB       6         12
//        M       <= 32 768      512    ~113 KB       2         13
//        L       <= 65 535     1024    ~225 KB       1         14
// Algorithmic HBM bytes per chunk for the roofline: N_raw read (the 4 N_raw scratch write is traffic of this design, not of
// the algorithm — reported separately in DESIGN.md).
#pragma once
#include "zwz_common.cuh"

namespace zwz {

#define ZWZ_DM_NIL 0xffffu

template <int CLS> struct MatchClass;
template <> struct MatchClass<0> {
    static constexpr uint32_t kCap = 8192, kThreads = 256, kHBits = 12, kData = 8192 + 64;
    static constexpr uint32_t kPrevBytes = ((kCap + 31u) & ~31u) * 2u, kSmem = kData + kPrevBytes + (2u << kHBits) + 512u;
};
template <> struct MatchClass<1> {
    static constexpr uint32_t kCap = 32768, kThreads = 512, kHBits = 13, kData = 32768 + 64;
    static constexpr uint32_t kPrevBytes = ((kCap + 31u) & ~31u) * 2u, kSmem = kData + kPrevBytes + (2u << kHBits) + 512u;
};
template <> struct MatchClass<2> {
    static constexpr uint32_t kCap = 65535, kThreads = 1024, kHBits = 14, kData = 65600;
    static constexpr uint32_t kPrevBytes = ((kCap + 31u) & ~31u) * 2u, kSmem = kData + kPrevBytes + (2u << kHBits) + 512u;
};

struct MatchCtl {
    unsigned long long mbar;   // mbarrier for the bulk copy
    volatile uint32_t front;   // positions < front have final chains
    uint32_t next_tile;
    uint32_t cur_work;
    uint32_t pad;
    uint32_t adler_a[32], adler_b[32], adler_len[32];
};

template <int HBITS> ZWZ_DEV uint32_t dm_hash3(uint32_t b0, uint32_t b1, uint32_t b2) {
    uint32_t v = b0 | (b1 << 8) | (b2 << 16);
    return (v * 0x9E3779B1u) >> (32 - HBITS);
}

#ifndef ZWZ_EMU
ZWZ_DEV uint32_t dm_smem_addr(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }
ZWZ_DEV void dm_mbar_init(unsigned long long *bar) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(dm_smem_addr(bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
ZWZ_DEV void dm_bulk_g2s(void *dst, const void *src, uint32_t bytes, unsigned long long *bar) {
    // generic-proxy accesses to this smem (previous chunk) are ordered before the async-proxy write
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(dm_smem_addr(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dm_smem_addr(dst)),
                 "l"(src), "r"(bytes), "r"(dm_smem_addr(bar))
                 : "memory");
}
ZWZ_DEV void dm_mbar_wait(unsigned long long *bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done)
                     : "r"(dm_smem_addr(bar)), "r"(parity)
                     : "memory");
    }
}
#endif

template <int CLS>
ZWZ_KERNEL __launch_bounds__(MatchClass<CLS>::kThreads) lz_match_kernel(DeflateJob job, const uint32_t *__restrict__ order, uint32_t n_work,
                                                                         uint32_t *work_counter) {
    constexpr uint32_t T = MatchClass<CLS>::kThreads, HB = MatchClass<CLS>::kHBits, DATA = MatchClass<CLS>::kData;
    constexpr uint32_t NW = T / 32u;
    ZWZ_DYN_SMEM(smem);
    const uint32_t *dataw = (const uint32_t *) smem;                                  // chunk bytes as aligned words
    uint16_t *prev = (uint16_t *) (smem + DATA);
    uint16_t *head = (uint16_t *) (smem + DATA + MatchClass<CLS>::kPrevBytes);
    MatchCtl *ctl = (MatchCtl *) (smem + DATA + MatchClass<CLS>::kPrevBytes + (2u << HB));
    const uint8_t *datab = (const uint8_t *) smem;
    const unsigned tid = threadIdx.x, lane = lane_id(), wid = warp_id();
    uint32_t parity = 0;

#ifndef ZWZ_EMU
    if (tid == 0) dm_mbar_init(&ctl->mbar);
#endif
    __syncthreads();

    for (;;) {
        if (tid == 0) ctl->cur_work = atomicAdd(work_counter, 1u);
        __syncthreads();
        const uint32_t w = ctl->cur_work;
        if (w >= n_work) break;
        const uint32_t c = order[w];
        const uint32_t n = job.raw_len[c];
        const uint8_t *src = job.raw + job.raw_off[c];
        const uint32_t skew = (uint32_t) ((uintptr_t) src & 15u);
        const uint32_t stage_bytes = (skew + n + 15u) & ~15u;

        // ---- 1. stage the chunk ----
#ifndef ZWZ_EMU
        if (tid == 0 && stage_bytes) dm_bulk_g2s(smem, src - skew, stage_bytes, &ctl->mbar);
#else
        for (uint32_t i = tid; i < skew + n; i += T) smem[i] = i < skew ? 0 : src[i - skew];
#endif
        for (uint32_t i = tid; i < (1u << HB) / 2u; i += T) ((uint32_t *) head)[i] = 0xffffffffu;
        if (tid == 0) {
            ctl->front = 0;
            ctl->next_tile = 0;
        }
#ifndef ZWZ_EMU
        if (stage_bytes) dm_mbar_wait(&ctl->mbar, parity);
        parity ^= (stage_bytes != 0u);
#endif
        __syncthreads();
        // bytes past the end take part in 4-byte compares: make them deterministic (the final clamp to `maxlen` makes
        // their value irrelevant for the result)
        if (tid < 32u) smem[skew + n + tid] = 0;
        __syncthreads();

        uint32_t *mout = job.scratch + job.scr_off[c];
        const uint32_t nhash = n >= 3u ? n - 2u : 0u; // positions that own a 3-byte hash
        const uint32_t ntiles = (n + 31u) >> 5;

        // ---- 2. hash every position (parallel); prev[p] holds hash(p) until the builder replaces it by the link ----
        for (uint32_t p = tid; p < nhash; p += T) {
            uint32_t v = ld32u(dataw, skew + p);
            prev[p] = (uint16_t) dm_hash3<HB>(v & 0xffu, (v >> 8) & 0xffu, (v >> 16) & 0xffu);
        }
        __syncthreads();

        // ---- 3. warp 0: exact chain build ----
        // 32 positions per step. Common case (all 32 hashes distinct): read the old head, store it as the link, store the
        // position as the new head. Lanes that share a hash inside the step are detected by reading the head back (one of
        // them won the store race, the others see a foreign position) and repaired group by group with ballot + shuffle,
        // which costs per COLLIDING group — __match_any_sync would cost per DISTINCT value (measured ~600 cycles per step
        // on sm_100a, which made this warp the bottleneck of the whole kernel).
        if (wid == 0) {
            const unsigned lt = (1u << lane) - 1u;
            uint32_t hn = lane < nhash ? (uint32_t) prev[lane] : 0u;
            for (uint32_t p0 = 0; p0 < nhash; p0 += 32u) {
                const uint32_t p = p0 + lane;
                const bool valid = p < nhash;
                const uint32_t h = hn;
                const uint32_t pn = p + 32u;
                hn = pn < nhash ? (uint32_t) prev[pn] : 0u; // next step's hash, fetched ahead of this step's stores
                uint32_t old = ZWZ_DM_NIL;
                if (valid) old = head[h];
                __syncwarp(); // all reads of head[] precede the writes of this step
                if (valid) {
                    prev[p] = (uint16_t) old;
                    head[h] = (uint16_t) p;
                }
                __syncwarp();
                uint32_t rb = valid ? (uint32_t) head[h] : p;
                unsigned rem = __ballot_sync(ZWZ_FULL, rb != p); // lanes that lost a store race
                while (rem) {                                    // one round per colliding group (warp-uniform loop)
                    const int leader = __ffs((int) rem) - 1;
                    const uint32_t hl = __shfl_sync(ZWZ_FULL, h, leader);
                    const unsigned same = __ballot_sync(ZWZ_FULL, valid && h == hl);
                    if (valid && h == hl) {
                        const unsigned lower = same & lt;
                        if (lower) prev[p] = (uint16_t) (p0 + 31u - (uint32_t) __clz((int) lower)); // nearest earlier lane
                        if ((same >> lane) == 1u) head[h] = (uint16_t) p;                               // highest lane is the new head
                    }
                    rem &= ~same;
                }
                if ((p0 & 224u) == 224u) { // publish the front every 8 steps
                    __threadfence_block();
                    __syncwarp();
                    if (lane == 0) ctl->front = p0 + 32u;
                } else {
                    __syncwarp();
                }
            }
            __threadfence_block();
            __syncwarp();
            if (lane == 0) ctl->front = 0x7fffffffu;
        }

        // ---- 4. search: one lane per position, 32-position tiles ----
        for (;;) {
            uint32_t tile = 0;
            if (lane == 0) tile = atomicAdd(&ctl->next_tile, 1u);
            tile = __shfl_sync(ZWZ_FULL, tile, 0);
            if (tile >= ntiles) break;
            const uint32_t need = (tile + 1u) * 32u < nhash ? (tile + 1u) * 32u : nhash;
            // Searchers that caught up with the build front must get out of its way: a polling warp burns issue slots of the
            // scheduler it may share with the builder (measured: 8 polling warps slowed the builder 6x), so back off
            // exponentially — the front moves 256 positions every ~0.5 us.
            for (uint32_t ns = 200u; ctl->front < need; ns = ns < 3200u ? ns * 2u : ns) ZWZ_SPIN_SLEEP(ns);
            __threadfence_block();

            const uint32_t p = tile * 32u + lane;
            uint32_t best_len = 2u, best_dist = 0u;
            if (p < nhash) {
                const uint32_t maxlen = (n - p) < ZWZ_MAX_MATCH ? (n - p) : ZWZ_MAX_MATCH;
                const uint32_t limit = p > ZWZ_MAX_DIST ? p - ZWZ_MAX_DIST : 0u;
                const uint32_t ap = skew + p;
                const uint32_t first3 = ld32u(dataw, ap) & 0x00ffffffu;
                uint32_t scan_end = datab[ap + 2u], scan_end1 = datab[ap + 1u]; // bytes at best_len and best_len - 1
                uint32_t cand = prev[p];
                uint32_t budget = job.depth;
                while (cand != ZWZ_DM_NIL && cand >= limit && budget-- != 0u) {
                    const uint32_t ac = skew + cand;
                    const uint32_t nxt = prev[cand]; // next link is fetched while this candidate is examined
                    // cheap rejects first (zlib's longest_match order): the byte that would extend the best match, its
                    // predecessor, then the 3-byte prefix (hash collisions)
                    if (datab[ac + best_len] == scan_end && datab[ac + best_len - 1u] == scan_end1 &&
                        (ld32u(dataw, ac) & 0x00ffffffu) == first3) {
                        uint32_t len = 3u;
                        while (len < maxlen) {
                            uint32_t y = ld32u(dataw, ac + len) ^ ld32u(dataw, ap + len);
                            if (y) {
                                len += ((uint32_t) __ffs((int) y) - 1u) >> 3;
                                break;
                            }
                            len += 4u;
                        }
                        if (len > maxlen) len = maxlen;
                        if (len > best_len) {
                            best_len = len;
                            best_dist = p - cand;
                            if (len >= job.nice || len >= maxlen) break;
                            scan_end = datab[ap + len];
                            scan_end1 = datab[ap + len - 1u];
                        }
                    }
                    cand = nxt;
                }
                if (best_len == 3u && best_dist > 4096u) best_len = 2u; // zlib's TOO_FAR: such a match costs more than 3 literals
            }
            if (p < n) mout[p] = best_len >= 3u ? ((best_len << 16) | best_dist) : 0u;
        }
        __syncthreads();

        // ---- 6. Adler-32 of the chunk ----
        {
            uint32_t seg = (n + T - 1u) / T;
            uint32_t lo = tid * seg, hi = lo + seg;
            if (lo > n) lo = n;
            if (hi > n) hi = n;
            AdlerPart part;
            part.a = 0;
            part.b = 0;
            part.len = hi - lo;
            for (uint32_t i = lo; i < hi; ++i) { // seg <= 64: no overflow
                part.a += datab[skew + i];
                part.b += part.a;
            }
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                AdlerPart o;
                o.a = __shfl_down_sync(ZWZ_FULL, part.a, d);
                o.b = __shfl_down_sync(ZWZ_FULL, part.b, d);
                o.len = __shfl_down_sync(ZWZ_FULL, part.len, d);
                if ((lane & (2u * d - 1u)) == 0u) part = adler_combine(part, o);
            }
            if (lane == 0) {
                ctl->adler_a[wid] = part.a;
                ctl->adler_b[wid] = part.b;
                ctl->adler_len[wid] = part.len;
            }
            __syncthreads();
            if (wid == 0) {
                part.a = lane < NW ? ctl->adler_a[lane] : 0u;
                part.b = lane < NW ? ctl->adler_b[lane] : 0u;
                part.len = lane < NW ? ctl->adler_len[lane] : 0u;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    AdlerPart o;
                    o.a = __shfl_down_sync(ZWZ_FULL, part.a, d);
                    o.b = __shfl_down_sync(ZWZ_FULL, part.b, d);
                    o.len = __shfl_down_sync(ZWZ_FULL, part.len, d);
                    if ((lane & (2u * d - 1u)) == 0u) part = adler_combine(part, o);
                }
                if (lane == 0) {
                    uint32_t a = (1u + part.a) % 65521u;
                    uint32_t b = (n % 65521u + part.b) % 65521u;
                    job.adler[c] = (b << 16) | a;
                }
            }
        }
        __syncthreads(); // smem is reused by the next chunk
    }
}

} // namespace zwz
