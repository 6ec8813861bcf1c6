// inflate.cuh — batched zlib-wrapped DEFLATE decoder, one warp per stream.
//
// Replaces decompression.cpp:11-37 (inflateInit / inflate(Z_NO_FLUSH) loop / inflateEnd, return codes ignored).
// Output contract = the bytes zlib 1.3 would have written for the same input (see oracle/zwz_oracle.c for the rules):
// a literal needs its whole code inside the input, a match needs length+extra+distance+extra, a stored block copies
// min(LEN, bytes left), and whatever was produced before an error or the end of the input stays.
//
// Layout per warp
//   * the compressed stream is read as aligned 32-bit words, 32 words at a time (one coalesced 128-byte load, the next
//     128 bytes already in flight); the word feeding the bit buffer comes from a register of the owning lane by shuffle;
//   * the bit buffer and every decode decision are computed redundantly by all 32 lanes (warp-uniform control flow, no
//     divergence, no broadcast needed), so the lanes are all there when a match has to be copied;
//   * literal/length and distance codes resolve through shared-memory lookup tables (2^10 and 2^8 entries, one LDS per
//     symbol); longer codes fall back to a canonical walk;
//   * literals are parked in the lane `position & 31` and stored 32 at a time; back-references are copied by the whole
//     warp, 32 bytes per step (pattern-replicated when distance < 32).
// Output bytes are written straight to global memory; back-references re-read them through L1/L2 (the window of one
// 65 535-byte record fits L1+L2 trivially), so shared memory only holds the ~6 KB of tables per warp and 36 warps fit an SM.
//
// Algorithmic bytes per stream: N_comp read + N_raw written.
#pragma once
#include "zwz_common.cuh"

namespace zwz {

#define ZWZ_INF_WARPS 1 // one warp per CTA: ~14 KB of shared memory each, so the CTA count per SM is what shared memory allows
#define ZWZ_INF_LBITS 10
#define ZWZ_INF_DBITS 8

enum { INF_KIND_LIT = 0, INF_KIND_BASE = 1, INF_KIND_EOB = 2, INF_KIND_SPECIAL = 3 };
// entry: [3:0] code length, [7:4] extra bits, [9:8] kind, [31:16] literal byte / base length / base distance.
//        kind SPECIAL: extra-bits field 1 = code longer than the table index (canonical walk needed);
//                      extra-bits field 0 = invalid code whose length is in [3:0] (zlib's op=64 entries)
#define ZWZ_INF_ENTRY(nb, eb, kind, val) ((uint32_t) (nb) | ((uint32_t) (eb) << 4) | ((uint32_t) (kind) << 8) | ((uint32_t) (val) << 16))
#define ZWZ_INF_INVALID(nb) ZWZ_INF_ENTRY((nb), 0, INF_KIND_SPECIAL, 0)
#define ZWZ_INF_LONG ZWZ_INF_ENTRY(0, 1, INF_KIND_SPECIAL, 0)

struct InflateWarpSmem {
    uint32_t lit[1 << ZWZ_INF_LBITS];
    uint32_t dst[1 << ZWZ_INF_DBITS];
    uint16_t sorted_ll[288];
    uint16_t sorted_d[32];
    uint32_t cnt_ll[16];
    uint32_t cnt_d[16];
    uint32_t run[16]; // running offsets while sorting
    uint32_t cw[32u * 15u + 8u];  // inflate_fast.cuh: staged window of compressed words (ZWZ_IF_CW), then the resolve staging area
    union {
        struct { // block header parsing only: dead once the tables are built, when the token regions come to life
            uint8_t lens[352]; // [0,19) code-length-code lengths; [32, 32+316) literal/length then distance lengths
            uint32_t clt[128]; // code-length-code table (7 bits)
            uint16_t sorted_cl[20];
            uint32_t cnt_cl[16];
        };
        uint16_t tok[32u * 98u];  // inflate_fast.cuh: one token region per lane (32 x ZWZ_IF_RS)
    };
};
#define ZWZ_INF_SMEM (ZWZ_INF_WARPS * (uint32_t) sizeof(zwz::InflateWarpSmem))

ZWZ_DEV uint32_t inf_litlen_entry(uint32_t sym, uint32_t nb) {
    if (sym < 256u) return ZWZ_INF_ENTRY(nb, 0, INF_KIND_LIT, sym);
    if (sym == 256u) return ZWZ_INF_ENTRY(nb, 0, INF_KIND_EOB, 0);
    if (sym > 285u) return ZWZ_INF_INVALID(nb);
    uint32_t k = sym - 257u, eb, base;
    if (k < 8u) {
        eb = 0;
        base = 3u + k;
    } else if (k == 28u) {
        eb = 0;
        base = 258u;
    } else {
        eb = (k - 4u) >> 2;
        base = 3u + ((4u + (k & 3u)) << eb);
    }
    return ZWZ_INF_ENTRY(nb, eb, INF_KIND_BASE, base);
}
ZWZ_DEV uint32_t inf_dist_entry(uint32_t sym, uint32_t nb) {
    if (sym > 29u) return ZWZ_INF_INVALID(nb);
    uint32_t eb, base;
    if (sym < 4u) {
        eb = 0;
        base = 1u + sym;
    } else {
        eb = (sym - 2u) >> 1;
        base = 1u + ((2u + (sym & 1u)) << eb);
    }
    return ZWZ_INF_ENTRY(nb, eb, INF_KIND_BASE, base);
}
ZWZ_DEV uint32_t inf_cl_entry(uint32_t sym, uint32_t nb) { return ZWZ_INF_ENTRY(nb, 0, INF_KIND_LIT, sym); }

// Canonical-code walk over `bits` (LSB first), lengths 1..maxwalk. Returns entry-maker input (sym,len) or ~0u.
ZWZ_DEV uint32_t inf_canon_walk(uint32_t bits, const uint32_t *cnt, const uint16_t *sorted, uint32_t maxwalk, uint32_t &len_out) {
    int code = 0, first = 0, index = 0;
    for (uint32_t len = 1; len <= maxwalk; ++len) {
        code |= (int) (bits & 1u);
        bits >>= 1;
        int c = (int) cnt[len];
        if (code - c < first) {
            len_out = len;
            return sorted[index + (code - first)];
        }
        index += c;
        first += c;
        first <<= 1;
        code <<= 1;
    }
    return 0xffffffffu;
}

// Warp-cooperative table build. kind: 0 = code-length code (must be complete), 1 = literal/length, 2 = distance.
// Returns (warp-uniform) 0 ok / 1 invalid set. max_len_out = longest code (0 = empty set).
template <int KIND>
ZWZ_DEV int inf_build(const uint8_t *lens, uint32_t n, uint32_t *cnt, uint16_t *sorted, uint32_t *run, uint32_t *table, uint32_t tbits,
                      uint32_t &max_len_out) {
    const unsigned lane = lane_id();
    if (lane < 16u) cnt[lane] = 0;
    __syncwarp();
    for (uint32_t s = lane; s < n; s += 32u) atomicAdd(&cnt[lens[s]], 1u);
    __syncwarp();
    // validity (inftrees.c rules), computed redundantly by every lane
    uint32_t max_len = 0;
    int left = 1, bad = 0;
    uint32_t acc = 0;
    for (uint32_t l = 1; l <= 15u; ++l) {
        uint32_t c = cnt[l];
        if (c) max_len = l;
        left = (left << 1) - (int) c;
        if (left < 0) bad = 1;
        acc += c;
    }
    __syncwarp();
    // running offsets per length: run[l] = start of length-l symbols in `sorted`
    if (lane == 0) {
        uint32_t o = 0;
        for (uint32_t l = 1; l <= 15u; ++l) {
            run[l] = o;
            o += cnt[l];
        }
        run[0] = 0;
    }
    __syncwarp();
    max_len_out = max_len;
    if (bad) return 1;
    if (max_len != 0 && left > 0 && (KIND == 0 || max_len != 1u)) return 1; // incomplete set
    // stable sort of symbols by (length, symbol): 32 symbols per step
    for (uint32_t base = 0; base < n; base += 32u) {
        uint32_t s = base + lane;
        uint32_t L = s < n ? lens[s] : 0u;
        unsigned grp = __match_any_sync(ZWZ_FULL, L);
        if (L) {
            uint32_t r = (uint32_t) __popc(grp & ((1u << lane) - 1u));
            sorted[run[L] + r] = (uint16_t) s;
        }
        __syncwarp();
        if (L && (grp >> lane) == 1u) run[L] += (uint32_t) __popc(grp); // highest lane of each group
        __syncwarp();
    }
    // fill: every table slot resolves its own prefix
    uint32_t walk = max_len < tbits ? max_len : tbits;
    for (uint32_t e = lane; e < (1u << tbits); e += 32u) {
        uint32_t len = 0;
        uint32_t sym = inf_canon_walk(e, cnt, sorted, walk, len);
        uint32_t ent;
        if (sym == 0xffffffffu)
            ent = (max_len > tbits) ? ZWZ_INF_LONG : ZWZ_INF_INVALID(max_len ? max_len : 1u);
        else
            ent = KIND == 0 ? inf_cl_entry(sym, len) : (KIND == 1 ? inf_litlen_entry(sym, len) : inf_dist_entry(sym, len));
        table[e] = ent;
    }
    __syncwarp();
    return 0;
}

// Warp-uniform bit reader over aligned words with a 32-word register cache per lane.
struct InfBits {
    const uint32_t *wbase;  // aligned word holding the stream's first byte
    uint32_t skew;          // byte offset of the stream inside that word
    uint32_t nbytes;        // stream length
    uint32_t nwords;        // words covering [skew, skew + nbytes)
    uint32_t nfull;         // leading words that lie entirely inside the stream: while widx <= nfull every fed bit is real
    uint32_t widx;          // next word to feed
    uint32_t cache, cache_next;
    uint32_t cache_blk;     // block (32 words) held in `cache`; cache_next holds cache_blk + 1
    uint64_t hold;
    uint32_t cnt;           // valid bits in hold
    uint64_t fed;           // stream bits fed into hold so far (may run past 8 * nbytes: zero padding)
};

ZWZ_DEV uint32_t infb_load(const InfBits &b, uint32_t k) {
    if (k >= b.nwords) return 0u;
    uint32_t w = __ldg(b.wbase + k);
    // bytes of the last word that lie past the end of the stream belong to somebody else: zero them
    uint32_t end = b.skew + b.nbytes; // byte index one past the stream, relative to wbase
    if ((k + 1u) * 4u > end) {
        uint32_t keep = end - k * 4u; // 1..3
        w &= (1u << (keep * 8u)) - 1u;
    }
    return w;
}
ZWZ_DEV uint32_t infb_next_word(InfBits &b) {
    uint32_t blk = b.widx >> 5;
    if (blk != b.cache_blk) { // warp-uniform
        if (blk == b.cache_blk + 1u) {
            b.cache = b.cache_next;
        } else {
            b.cache = infb_load(b, blk * 32u + lane_id());
        }
        b.cache_next = infb_load(b, (blk + 1u) * 32u + lane_id());
        b.cache_blk = blk;
    }
    uint32_t w = __shfl_sync(ZWZ_FULL, b.cache, (int) (b.widx & 31u));
    b.widx++;
    return w;
}
// position the reader at stream byte `byte_pos`
ZWZ_DEV void infb_seek(InfBits &b, uint32_t byte_pos) {
    uint32_t a = b.skew + byte_pos;
    b.widx = a >> 2;
    b.cache_blk = 0xfffffff0u;
    b.hold = 0;
    b.cnt = 0;
    uint32_t w = infb_next_word(b);
    uint32_t drop = (a & 3u) * 8u;
    b.hold = (uint64_t) (w >> drop);
    b.cnt = 32u - drop;
    b.fed = (uint64_t) byte_pos * 8u + b.cnt;
}
ZWZ_DEV void infb_refill(InfBits &b) {
    if (b.cnt <= 32u) {
        uint32_t w = infb_next_word(b);
        b.hold |= (uint64_t) w << b.cnt;
        b.cnt += 32u;
        b.fed += 32u;
    }
}
ZWZ_DEV void infb_drop(InfBits &b, uint32_t n) {
    b.hold >>= n;
    b.cnt -= n;
}
// bits consumed from the stream so far
ZWZ_DEV uint64_t infb_used(const InfBits &b) { return b.fed - b.cnt; }
// position the reader at stream bit `bit_pos`
ZWZ_DEV void infb_seek_bits(InfBits &b, uint64_t bit_pos) {
    infb_seek(b, (uint32_t) (bit_pos >> 3));
    infb_drop(b, (uint32_t) bit_pos & 7u); // <= 7 of the >= 8 bits the seek left in the buffer
}

} // namespace zwz
#include "inflate_fast.cuh"
namespace zwz {

// One warp decodes stream `sid`.
ZWZ_DEV void inflate_stream(InflateWarpSmem &S, const uint8_t *__restrict__ comp, uint32_t comp_len, uint8_t *out, uint32_t cap,
                            uint32_t flags, uint32_t &raw_len_out, uint32_t &status_out) {
    const unsigned lane = lane_id();
    const uint64_t total_bits = (uint64_t) comp_len * 8u;
    InfBits B;
    B.skew = (uint32_t) ((uintptr_t) comp & 3u);
    B.wbase = (const uint32_t *) (comp - B.skew);
    B.nbytes = comp_len;
    B.nwords = (B.skew + comp_len + 3u) >> 2;
    B.nfull = (B.skew + comp_len) >> 2;
    infb_seek(B, 0);

    uint32_t pos = 0;       // bytes produced (counted past cap too)
    uint32_t pend_lo = 0;   // literals [pend_lo, pos) are parked in lanes (q & 31)
    uint32_t mybyte = 0;
    uint32_t status = ZWZ_STREAM_END;
    bool overflow = false;

#define INF_FLUSH_LITS()                                                         \
    do {                                                                         \
        uint32_t q_ = (pend_lo & ~31u) + lane;                                   \
        if (q_ >= pend_lo && q_ < pos && q_ < cap) out[q_] = (uint8_t) mybyte;   \
        pend_lo = pos;                                                           \
    } while (0)
// Everything already in `hold` is real stream data while fed <= total_bits; only the tail of a stream (or a truncated
// one) needs the per-symbol availability checks that give zlib's exact stop position.
#define INF_FAST() (B.widx <= B.nfull)
#define INF_NEED(nbits_) (INF_FAST() || infb_used(B) + (uint64_t) (nbits_) <= total_bits)

    // ---- RFC 1950 header (inflate.c HEAD) ----
    infb_refill(B);
    if (!INF_NEED(16)) {
        status = ZWZ_STREAM_TRUNCATED;
        goto done;
    }
    {
        uint32_t cmf = (uint32_t) B.hold & 0xffu, flg = ((uint32_t) B.hold >> 8) & 0xffu;
        if (((cmf << 8) + flg) % 31u != 0u || (cmf & 15u) != 8u || (cmf >> 4) + 8u > 15u || (flg & 0x20u)) {
            status = ZWZ_STREAM_BAD;
            goto done;
        }
        infb_drop(B, 16);
    }

    for (;;) {
        infb_refill(B);
        if (!INF_NEED(3)) {
            status = ZWZ_STREAM_TRUNCATED;
            goto done;
        }
        uint32_t last = (uint32_t) B.hold & 1u;
        uint32_t type = ((uint32_t) B.hold >> 1) & 3u;
        infb_drop(B, 3);

        if (type == 0u) {
            // stored: skip to the byte boundary, LEN/NLEN, then a plain copy
            uint64_t used = infb_used(B);
            uint32_t bpos = (uint32_t) ((used + 7u) >> 3);
            if ((uint64_t) bpos + 4u > comp_len) {
                status = ZWZ_STREAM_TRUNCATED;
                goto done;
            }
            infb_seek(B, bpos);
            infb_refill(B);
            uint32_t v = (uint32_t) B.hold;
            if ((v & 0xffffu) != ((v >> 16) ^ 0xffffu)) {
                status = ZWZ_STREAM_BAD;
                goto done;
            }
            uint32_t len = v & 0xffffu;
            bpos += 4u;
            uint32_t avail = comp_len - bpos;
            uint32_t ncopy = len < avail ? len : avail;
            INF_FLUSH_LITS();
            {
                const uint32_t room = pos < cap ? cap - pos : 0u;
                inf_copy_plain(out + pos, comp + bpos, ncopy < room ? ncopy : room);
            }
            if (pos + ncopy > cap) overflow = true;
            pos += ncopy;
            pend_lo = pos;
            if (ncopy < len) {
                status = ZWZ_STREAM_TRUNCATED;
                goto done;
            }
            infb_seek(B, bpos + len);
            if (last) break;
            continue;
        }
        if (type == 3u) {
            status = ZWZ_STREAM_BAD;
            goto done;
        }

        uint32_t max_ll = 0, max_d = 0;
        uint32_t min_ll = 1; // shortest literal/length code of this block: bounds how many codes fit a 32-bit window
        if (type == 1u) {
            for (uint32_t s = lane; s < 288u; s += 32u) S.lens[32u + s] = (uint8_t) (s < 144u ? 8 : (s < 256u ? 9 : (s < 280u ? 7 : 8)));
            S.lens[320u + lane] = 5;
            __syncwarp();
            inf_build<1>(S.lens + 32, 288u, S.cnt_ll, S.sorted_ll, S.run, S.lit, ZWZ_INF_LBITS, max_ll);
            inf_build<2>(S.lens + 320, 32u, S.cnt_d, S.sorted_d, S.run, S.dst, ZWZ_INF_DBITS, max_d);
        } else {
            if (!INF_NEED(14)) {
                status = ZWZ_STREAM_TRUNCATED;
                goto done;
            }
            uint32_t nlen = ((uint32_t) B.hold & 31u) + 257u;
            uint32_t ndist = (((uint32_t) B.hold >> 5) & 31u) + 1u;
            uint32_t ncode = (((uint32_t) B.hold >> 10) & 15u) + 4u;
            infb_drop(B, 14);
            if (nlen > 286u || ndist > 30u) {
                status = ZWZ_STREAM_BAD;
                goto done;
            }
            if (lane < 19u) S.lens[lane] = 0;
            __syncwarp();
            for (uint32_t i = 0; i < ncode; ++i) {
                infb_refill(B);
                if (!INF_NEED(3)) {
                    status = ZWZ_STREAM_TRUNCATED;
                    goto done;
                }
                // RFC 1951 §3.2.7 permutation 16 17 18 0 8 7 9 6 10 5 11 4 12 3 13 2 14 1 15
                uint32_t slot = i < 3u ? 16u + i : (i == 3u ? 0u : ((i & 1u) ? 7u - ((i - 5u) >> 1) : 8u + ((i - 4u) >> 1)));
                if (lane == 0) S.lens[slot] = (uint8_t) ((uint32_t) B.hold & 7u);
                infb_drop(B, 3);
            }
            __syncwarp();
            uint32_t max_cl = 0;
            if (inf_build<0>(S.lens, 19u, S.cnt_cl, S.sorted_cl, S.run, S.clt, 7u, max_cl) != 0 || max_cl == 0u) {
                status = ZWZ_STREAM_BAD;
                goto done;
            }
            // code lengths for nlen + ndist symbols (warp-uniform decode, lane 0 stores)
            uint32_t have = 0, total = nlen + ndist, prev_len = 0;
            while (have < total) {
                infb_refill(B);
                uint32_t e = S.clt[(uint32_t) B.hold & 127u];
                uint32_t nb = e & 15u;
                if (nb == 0u) { // invalid pattern is impossible for a complete code; defensive
                    status = ZWZ_STREAM_BAD;
                    goto done;
                }
                uint32_t sym = e >> 16;
                uint32_t eb = sym < 16u ? 0u : (sym == 16u ? 2u : (sym == 17u ? 3u : 7u));
                if (!INF_NEED(nb + eb)) {
                    status = ZWZ_STREAM_TRUNCATED;
                    goto done;
                }
                infb_drop(B, nb);
                if (sym < 16u) {
                    if (lane == 0) S.lens[32u + have] = (uint8_t) sym;
                    prev_len = sym;
                    have++;
                    continue;
                }
                uint32_t rep, val = 0;
                uint32_t x = (uint32_t) B.hold & ((1u << eb) - 1u);
                infb_drop(B, eb);
                if (sym == 16u) {
                    if (have == 0u) {
                        status = ZWZ_STREAM_BAD;
                        goto done;
                    }
                    val = prev_len;
                    rep = 3u + x;
                } else if (sym == 17u) {
                    rep = 3u + x;
                } else {
                    rep = 11u + x;
                }
                if (have + rep > total) {
                    status = ZWZ_STREAM_BAD;
                    goto done;
                }
                if (lane < rep) S.lens[32u + have + lane] = (uint8_t) val;
                if (lane + 32u < rep) S.lens[32u + have + lane + 32u] = (uint8_t) val;
                if (lane + 64u < rep) S.lens[32u + have + lane + 64u] = (uint8_t) val;
                if (lane + 96u < rep) S.lens[32u + have + lane + 96u] = (uint8_t) val;
                if (lane + 128u < rep) S.lens[32u + have + lane + 128u] = (uint8_t) val;
                prev_len = val;
                have += rep;
            }
            __syncwarp();
            if (S.lens[32u + 256u] == 0) { // missing end-of-block code
                status = ZWZ_STREAM_BAD;
                goto done;
            }
            // S.lens[32 .. 32+nlen) literal/length, then ndist distance lengths. inf_build reads them in place; the two
            // builds use disjoint outputs, and the distance lengths are read before anything overwrites them.
            if (inf_build<1>(S.lens + 32, nlen, S.cnt_ll, S.sorted_ll, S.run, S.lit, ZWZ_INF_LBITS, max_ll) != 0 ||
                inf_build<2>(S.lens + 32 + nlen, ndist, S.cnt_d, S.sorted_d, S.run, S.dst, ZWZ_INF_DBITS, max_d) != 0) {
                status = ZWZ_STREAM_BAD;
                goto done;
            }
        }

        for (min_ll = 1; min_ll < 15u && S.cnt_ll[min_ll] == 0u; ++min_ll) {}
        // ---- lane-parallel decode of the block (inflate_fast.cuh); the careful loop below only sees what it hands over:
        // the first token that does not lie completely inside the input, or an invalid code ----
        if (!(flags & ZWZ_INFLATE_CAREFUL)) {
            INF_FLUSH_LITS();
            __syncwarp();
            uint64_t bp = infb_used(B);
            const int fr = inf_fast_block(S, B, bp, max_ll, max_d, out, cap, pos, overflow);
            pend_lo = pos;
            if (fr == IFR_BAD) {
                status = ZWZ_STREAM_BAD;
                goto done;
            }
            infb_seek_bits(B, bp);
            if (fr == IFR_EOB) {
                if (last) break;
                continue;
            }
        }
        // ---- symbol loop ----
        for (;;) {
            infb_refill(B); // >= 33 valid bits
            uint32_t e;
            if (INF_FAST()) {
                // Literal runs, 32 bit offsets at a time ("warp-ballot symbol decode"): lane l decodes the code that WOULD
                // start at bit l of the buffer; the ballot says which offsets hold a complete literal; the offsets that
                // really are code starts form the chain 0 -> nb(0) -> ..., recovered with 5 rounds of pointer doubling on
                // (reach mask, jump) pairs. All literals on the chain are stored by their lanes in one go and the whole
                // run is dropped from the bit buffer at once. The chain ends at the first non-literal (handled below) or
                // past offset 31.
                const uint32_t el = S.lit[(uint32_t) (B.hold >> lane) & ((1u << ZWZ_INF_LBITS) - 1u)];
                const bool ok = (lane + 15u <= B.cnt) && ((el >> 8) & 3u) == INF_KIND_LIT;
                const unsigned litmask = __ballot_sync(ZWZ_FULL, ok);
                if (litmask & 1u) {
                    uint32_t J = lane + (el & 15u);
                    uint32_t R = ok ? (1u << lane) : 0u;
                    // a window of 32 bits holds at most ceil(32 / min_ll) code starts and r doubling rounds reach 2^r of them:
                    // 2 rounds when no code is shorter than 8 bits (near-uniform bytes: the JPEG-like class), 3 down to 4 bits
#define INF_ROUND()                                                                  \
    do {                                                                             \
        const bool go = ok && J < 32u && ((litmask >> J) & 1u);                      \
        const int from = go ? (int) J : (int) lane;                                  \
        const uint32_t Rj = __shfl_sync(ZWZ_FULL, R, from);                          \
        const uint32_t Jj = __shfl_sync(ZWZ_FULL, J, from);                          \
        if (go) {                                                                    \
            R |= Rj;                                                                 \
            J = Jj;                                                                  \
        }                                                                            \
    } while (0)
                    INF_ROUND();
                    INF_ROUND();
                    if (min_ll < 8u) {
                        INF_ROUND();
                        if (min_ll < 4u) {
                            INF_ROUND();
                            if (min_ll < 2u) INF_ROUND();
                        }
                    }
#undef INF_ROUND
                    const uint32_t R0 = __shfl_sync(ZWZ_FULL, R, 0);
                    const uint32_t J0 = __shfl_sync(ZWZ_FULL, J, 0);
                    const uint32_t nlit = (uint32_t) __popc(R0);
                    if (pend_lo != pos) INF_FLUSH_LITS(); // only after a step of the scalar path
                    if ((R0 >> lane) & 1u) {
                        const uint32_t q = pos + (uint32_t) __popc(R0 & ((1u << lane) - 1u));
                        if (q < cap) out[q] = (uint8_t) (el >> 16);
                    }
                    if (pos + nlit > cap) overflow = true;
                    pos += nlit;
                    pend_lo = pos;
                    infb_drop(B, J0);
                    continue;
                }
                e = __shfl_sync(ZWZ_FULL, el, 0);
            } else {
                e = S.lit[(uint32_t) B.hold & ((1u << ZWZ_INF_LBITS) - 1u)];
            }
            uint32_t kind = (e >> 8) & 3u;
            uint32_t nb = e & 15u;
            if (kind == INF_KIND_SPECIAL) {
                if ((e >> 4) & 15u) { // code longer than the table index: canonical walk (complete sets only get here)
                    uint32_t len = 0;
                    uint32_t sym = inf_canon_walk((uint32_t) B.hold, S.cnt_ll, S.sorted_ll, max_ll, len);
                    e = sym == 0xffffffffu ? ZWZ_INF_INVALID(max_ll) : inf_litlen_entry(sym, len);
                    kind = (e >> 8) & 3u;
                    nb = e & 15u;
                }
                if (kind == INF_KIND_SPECIAL) { // zlib's invalid-code entry: it still has to see the code's bits first
                    status = INF_NEED(nb) ? ZWZ_STREAM_BAD : ZWZ_STREAM_TRUNCATED;
                    goto done;
                }
            }
            if (!INF_NEED(nb)) {
                status = ZWZ_STREAM_TRUNCATED;
                goto done;
            }
            infb_drop(B, nb);
            if (kind == INF_KIND_LIT) {
                if (lane == (pos & 31u)) mybyte = e >> 16;
                pos++;
                if ((pos & 31u) == 0u) INF_FLUSH_LITS();
                continue;
            }
            if (kind == INF_KIND_EOB) break;
            // length (the code may have been the second one since the last refill: top up before the extra bits)
            infb_refill(B);
            uint32_t eb = (e >> 4) & 15u;
            uint32_t mlen = (e >> 16) + ((uint32_t) B.hold & ((1u << eb) - 1u));
            if (!INF_NEED(eb)) {
                status = ZWZ_STREAM_TRUNCATED;
                goto done;
            }
            infb_drop(B, eb);
            infb_refill(B);
            uint32_t d = S.dst[(uint32_t) B.hold & ((1u << ZWZ_INF_DBITS) - 1u)];
            uint32_t dnb = d & 15u;
            if (((d >> 8) & 3u) == INF_KIND_SPECIAL) {
                if ((d >> 4) & 15u) {
                    uint32_t len = 0;
                    uint32_t sym = inf_canon_walk((uint32_t) B.hold, S.cnt_d, S.sorted_d, max_d, len);
                    d = sym == 0xffffffffu ? ZWZ_INF_INVALID(max_d) : inf_dist_entry(sym, len);
                    dnb = d & 15u;
                }
                if (((d >> 8) & 3u) == INF_KIND_SPECIAL) {
                    status = INF_NEED(dnb) ? ZWZ_STREAM_BAD : ZWZ_STREAM_TRUNCATED;
                    goto done;
                }
            }
            uint32_t deb = (d >> 4) & 15u;
            if (!INF_NEED(dnb + deb)) {
                status = ZWZ_STREAM_TRUNCATED;
                goto done;
            }
            infb_drop(B, dnb);
            uint32_t dist = (d >> 16) + ((uint32_t) B.hold & ((1u << deb) - 1u));
            infb_drop(B, deb);
            if (dist > pos) { // "invalid distance too far back"
                status = ZWZ_STREAM_BAD;
                goto done;
            }
            // copy: everything before `pos` must be in memory first. Loads go through L2 (ld.global.cg): the bytes were
            // written by other lanes of this warp moments ago and L1 is not coherent for that.
            INF_FLUSH_LITS();
            __syncwarp();
            if (pos + mlen > cap) overflow = true;
            inf_copy_match(out, pos, mlen, dist, cap);
            pos += mlen;
            pend_lo = pos;
        }
        if (last) break;
    }
    // ---- RFC 1950 trailer (inflate.c CHECK) ----
    {
        INF_FLUSH_LITS();
        __syncwarp();
        uint64_t used = infb_used(B);
        uint32_t bpos = (uint32_t) ((used + 7u) >> 3);
        if ((uint64_t) bpos + 4u > comp_len) {
            status = ZWZ_STREAM_TRUNCATED;
            goto done;
        }
        if (!(flags & 1u) && !overflow) {
            uint32_t want = ((uint32_t) comp[bpos] << 24) | ((uint32_t) comp[bpos + 1] << 16) | ((uint32_t) comp[bpos + 2] << 8) | comp[bpos + 3];
            const uint32_t have = inf_adler32(out, pos); // a = 1 + sum(byte), b = n + sum((n - j) * byte_j)  (mod 65521)
            if (have != want) status = ZWZ_STREAM_BAD;
        }
    }
done:
    INF_FLUSH_LITS();
    __syncwarp();
    if (status == ZWZ_STREAM_END && overflow) status = ZWZ_STREAM_OUTPUT_FULL;
    raw_len_out = pos;
    status_out = status;
#undef INF_FLUSH_LITS
#undef INF_NEED
}

// Persistent warps: every warp pulls the next stream from a global counter, so a CTA is never kept alive by one long stream
// while its other warps idle (streams of one batch differ in length by 100x in config C2).
ZWZ_KERNEL __launch_bounds__(ZWZ_INF_WARPS * 32) inflate_kernel(const uint8_t *__restrict__ comp, const uint64_t *__restrict__ off,
                                                              const uint32_t *__restrict__ len, uint8_t *raw_out,
                                                              const uint64_t *__restrict__ raw_off, uint32_t *raw_len, uint32_t *status,
                                                              uint32_t n, uint32_t flags, uint32_t *work_counter) {
    ZWZ_DYN_SMEM(inf_raw);
    InflateWarpSmem *smem = (InflateWarpSmem *) inf_raw;
    for (;;) {
        uint32_t sid = 0;
        if (lane_id() == 0) sid = atomicAdd(work_counter, 1u);
        sid = __shfl_sync(ZWZ_FULL, sid, 0);
        if (sid >= n) break;
        uint64_t cap64 = raw_off[sid + 1] - raw_off[sid];
        uint32_t cap = cap64 > 0xfffffff0ull ? 0xfffffff0u : (uint32_t) cap64;
        uint32_t rl = 0, st = 0;
        inflate_stream(smem[warp_id()], comp + off[sid], len[sid], raw_out + raw_off[sid], cap, flags, rl, st);
        if (lane_id() == 0) {
            raw_len[sid] = rl;
            status[sid] = st;
        }
        __syncwarp();
    }
}

} // namespace zwz
