// deflate_match.cuh — LZ77 match finding for one <= 65 535-byte chunk per CTA, everything in shared memory.
//
// First half of the replacement for zlib's deflate() at compression.cpp:119-134 (the second half is deflate_encode.cuh).
//
// Per chunk (one persistent CTA of 1024 threads per SM, chunks handed out by an atomic counter):
//   1. the chunk is staged HBM -> shared memory with one bulk asynchronous copy (cp.async.bulk + mbarrier, SASS UBLKCP),
//      16-byte aligned superset of the chunk, the chunk itself starts at byte `skew` of the staging buffer;
//   2. warp 0 builds EXACT hash chains (3-byte hash, 14-bit head table, u16 prev[] per position) 32 positions per step:
//      lanes with the same hash inside the step are linked with __match_any_sync, the lowest of a group links to the
//      head table, the highest becomes the new head. The chain of position p is final as soon as the build front has
//      passed p, which it publishes through a shared-memory counter;
//   3. the other 31 warps (and warp 0 once it is done) pull 32-position tiles and, one lane per position, walk the
//      chain: 4-byte compares on funnel-shifted aligned words, early reject on the bytes around the current best
//      length, `depth` candidates at most, stop at `nice` bytes;
//   4. the best (length, distance) of EVERY position goes to the chunk's scratch in HBM (4 bytes per input byte,
//      coalesced); the parse in deflate_encode.cuh picks the path through them;
//   5. Adler-32 of the chunk is reduced from shared memory while it is there.
//
// Shared memory: 65 600 (chunk) + 131 072 (prev) + 32 768 (head) + control = 229 952 bytes -> one CTA per SM.
// Algorithmic HBM bytes per chunk for the roofline: N_raw read (+ the 4 N_raw scratch write, which is traffic of this
// design, not of the algorithm — reported separately in DESIGN.md).
#pragma once
#include "zwz_common.cuh"

namespace zwz {

#define ZWZ_DM_THREADS 1024
#define ZWZ_DM_HBITS 14
#define ZWZ_DM_NIL 0xffffu
#define ZWZ_DM_DATA_BYTES 65600u
#define ZWZ_DM_SMEM_BYTES (ZWZ_DM_DATA_BYTES + 131072u + (2u << ZWZ_DM_HBITS) + 512u)

struct MatchCtl {
    unsigned long long mbar;   // mbarrier for the bulk copy
    volatile uint32_t front;   // positions < front have final chains
    uint32_t next_tile;
    uint32_t cur_chunk;
    uint32_t pad;
    uint32_t adler_a[32], adler_b[32], adler_len[32];
};

ZWZ_DEV uint32_t dm_hash3(uint32_t b0, uint32_t b1, uint32_t b2) {
    uint32_t v = b0 | (b1 << 8) | (b2 << 16);
    return (v * 0x9E3779B1u) >> (32 - ZWZ_DM_HBITS);
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

ZWZ_KERNEL __launch_bounds__(ZWZ_DM_THREADS, 1) lz_match_kernel(DeflateJob job) {
    ZWZ_DYN_SMEM(smem);
    uint32_t *dataw = (uint32_t *) smem;                                                    // chunk bytes as aligned words
    uint16_t *prev = (uint16_t *) (smem + ZWZ_DM_DATA_BYTES);                               // [65536]
    uint16_t *head = (uint16_t *) (smem + ZWZ_DM_DATA_BYTES + 131072u);                     // [1 << HBITS]
    MatchCtl *ctl = (MatchCtl *) (smem + ZWZ_DM_DATA_BYTES + 131072u + (2u << ZWZ_DM_HBITS));
    const uint8_t *datab = (const uint8_t *) smem;
    const unsigned tid = threadIdx.x, lane = lane_id(), wid = warp_id();
    uint32_t parity = 0;

#ifndef ZWZ_EMU
    if (tid == 0) dm_mbar_init(&ctl->mbar);
#endif
    __syncthreads();

    for (;;) {
        if (tid == 0) ctl->cur_chunk = atomicAdd(job.work_counter, 1u);
        __syncthreads();
        const uint32_t c = ctl->cur_chunk;
        if (c >= job.n) break;
        const uint32_t n = job.raw_len[c];
        const uint8_t *src = job.raw + job.raw_off[c];
        const uint32_t skew = (uint32_t) ((uintptr_t) src & 15u);
        const uint32_t stage_bytes = (skew + n + 15u) & ~15u;

        // ---- 1. stage the chunk ----
#ifndef ZWZ_EMU
        if (tid == 0 && stage_bytes) dm_bulk_g2s(smem, src - skew, stage_bytes, &ctl->mbar);
#else
        for (uint32_t i = tid; i < skew + n; i += ZWZ_DM_THREADS) smem[i] = i < skew ? 0 : src[i - skew];
#endif
        for (uint32_t i = tid; i < (1u << ZWZ_DM_HBITS) / 2u; i += ZWZ_DM_THREADS) ((uint32_t *) head)[i] = 0xffffffffu;
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

        // ---- 2. warp 0: exact chain build ----
        if (wid == 0) {
            for (uint32_t p0 = 0; p0 < nhash; p0 += 32u) {
                uint32_t p = p0 + lane;
                bool valid = p < nhash;
                uint32_t h = 0x10000u + lane; // unique => singleton group
                if (valid) {
                    uint32_t a = skew + p;
                    h = dm_hash3(datab[a], datab[a + 1], datab[a + 2]);
                }
                unsigned grp = __match_any_sync(ZWZ_FULL, h);
                if (valid) {
                    unsigned lower = grp & ((1u << lane) - 1u);
                    uint32_t pv = lower ? p0 + (31u - (uint32_t) __clz((int) lower)) : (uint32_t) head[h];
                    prev[p] = (uint16_t) pv;
                }
                __syncwarp();
                if (valid && (grp >> lane) == 1u) head[h] = (uint16_t) p;
                __threadfence_block();
                __syncwarp();
                if (lane == 0) ctl->front = p0 + 32u;
            }
            __threadfence_block();
            if (lane == 0) ctl->front = 0x7fffffffu;
        }

        // ---- 3. search: one lane per position, 32-position tiles ----
        for (;;) {
            uint32_t tile = 0;
            if (lane == 0) tile = atomicAdd(&ctl->next_tile, 1u);
            tile = __shfl_sync(ZWZ_FULL, tile, 0);
            if (tile >= ntiles) break;
            const uint32_t need = (tile + 1u) * 32u < nhash ? (tile + 1u) * 32u : nhash;
            while (ctl->front < need) ZWZ_SPIN_PAUSE();
            __threadfence_block();

            const uint32_t p = tile * 32u + lane;
            uint32_t best_len = 2u, best_dist = 0u;
            if (p < nhash) {
                const uint32_t maxlen = (n - p) < ZWZ_MAX_MATCH ? (n - p) : ZWZ_MAX_MATCH;
                const uint32_t limit = p > ZWZ_MAX_DIST ? p - ZWZ_MAX_DIST : 0u;
                const uint32_t ap = skew + p;
                uint32_t cand = prev[p];
                uint32_t budget = job.depth;
                while (cand != ZWZ_DM_NIL && cand >= limit && budget-- != 0u) {
                    const uint32_t ac = skew + cand;
                    // early reject: the 4 bytes ending at index best_len must match for the candidate to be longer
                    uint32_t o = (best_len < 3u ? 3u : best_len) - 3u;
                    uint32_t x = ld32u(dataw, ac + o) ^ ld32u(dataw, ap + o);
                    if (best_len < 3u) x &= 0x00ffffffu;
                    if (x == 0u) {
                        uint32_t len = 0;
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
                        }
                    }
                    cand = prev[cand];
                }
                if (best_len == 3u && best_dist > 4096u) best_len = 2u; // zlib's TOO_FAR: such a match costs more than 3 literals
            }
            if (p < n) mout[p] = best_len >= 3u ? ((best_len << 16) | best_dist) : 0u; // dist 32768 needs all 16 low bits
        }
        __syncthreads();

        // ---- 5. Adler-32 of the chunk ----
        {
            uint32_t seg = (n + ZWZ_DM_THREADS - 1u) / ZWZ_DM_THREADS;
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
                part.a = ctl->adler_a[lane];
                part.b = ctl->adler_b[lane];
                part.len = ctl->adler_len[lane];
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
