// inflate_lanes.cuh — batched zlib-wrapped DEFLATE decoder, one LANE per stream (32 streams per warp).
//
// Same contract as inflate.cuh (decompression.cpp:11-37; output = the bytes zlib 1.3 would have written, status as in
// include/zwz_cuda.h), different mapping. inflate.cuh gives a whole warp to one stream and decodes every symbol redundantly
// in all 32 lanes: ~25 warp-instructions per output byte, and the kernel is issue-bound (profiles/round1). With thousands
// of independent streams per batch (370 000 in config C2, 16 384 in C3) the lanes can each run their own stream instead:
// the decode work per byte is the same per lane, but a warp-instruction now advances 32 streams.
//
// SIMT shape
//   * every lane keeps its own 64-bit bit buffer (fed from aligned 32-bit words, the next word always prefetched), its own
//     output cursor and its own state; the streams of a warp are of similar compressed size (the host passes them sorted),
//     so the lanes finish together;
//   * a round lets every lane take ONE step of its state: decode a literal/length symbol, decode a distance, copy up to
//     8 bytes of a match, or copy up to 8 bytes of a stored block. Lanes in different states run one after the other (that
//     is the SIMT price), lanes in the same state run together;
//   * block headers and table builds are long lane-serial code. A lane that reaches one parks; the warp builds when a
//     quarter of its live lanes are parked (or nobody else can move), so builds run many lanes wide.
//
// Shared memory: every lane owns one 4-byte bank. Word i of lane l lives at smem word i*32 + l, so any per-lane table
// access — same index or 32 different ones — is conflict-free. Per lane (words): literal/length LUT 2^9 x u16 (256),
// distance LUT 2^7 x u16 (64), symbols sorted by code (144 + 16), counts per length (8 + 8), running offsets (8), code
// lengths (80) = 584 words, 73 KB per warp, three warps per SM.
//
// Algorithmic bytes per stream: N_comp read + N_raw written.
#pragma once
#include "zwz_common.cuh"

namespace zwz {

#define ZWZ_IL_LBITS 9u
#define ZWZ_IL_DBITS 7u
#define ZWZ_IL_LUT_L 0u
#define ZWZ_IL_LUT_D 256u
#define ZWZ_IL_SORT_L 320u
#define ZWZ_IL_SORT_D 464u
#define ZWZ_IL_CNT_L 480u
#define ZWZ_IL_CNT_D 488u
#define ZWZ_IL_OFF 496u
#define ZWZ_IL_LENS 504u
#define ZWZ_IL_WORDS 584u
#define ZWZ_IL_SMEM (ZWZ_IL_WORDS * 32u * 4u)
// the code-length code lives in regions that are dead until the literal/length build
#define ZWZ_IL_CL_LENS ZWZ_IL_SORT_L
#define ZWZ_IL_CL_SORT ZWZ_IL_SORT_D
#define ZWZ_IL_CL_CNT ZWZ_IL_CNT_D
#define ZWZ_IL_CL_LUT ZWZ_IL_LUT_L

// per-lane views: W points at this lane's word 0
#define IL_U16(W, word0, idx) (((uint16_t *) ((W) + ((word0) + ((uint32_t) (idx) >> 1)) * 32u))[(idx) & 1u])
#define IL_U8(W, word0, idx) (((uint8_t *) ((W) + ((word0) + ((uint32_t) (idx) >> 2)) * 32u))[(idx) & 3u])

// LUT entry (u16): [3:0] code length, [15:4] symbol. LONG: the code is longer than the table index. INVALID: zlib's op=64
// entry of an incomplete/empty set (1 bit).
#define ZWZ_IL_LONG 0xfff0u
#define ZWZ_IL_INVALID 0xffe1u

enum { IL_BLOCK = 0, IL_TRAILER = 1, IL_SYM = 2, IL_DIST = 3, IL_COPY = 4, IL_STORED = 5, IL_DONE = 6 };

struct ILBits {
    const uint32_t *wbase; // aligned word holding the stream's first byte
    uint32_t skew;         // byte offset of the stream inside that word
    uint32_t end;          // skew + nbytes
    uint32_t nwords;       // words covering [skew, end)
    uint32_t widx;         // index of `nextw`
    uint32_t nextw;        // prefetched word widx
    uint64_t hold;
    uint32_t cnt;          // valid bits in hold (zero padding past the end of the stream included)
    uint32_t rem;          // stream bits not yet consumed: a step that needs more than this stops the stream (truncated)
};

ZWZ_DEV uint32_t ilb_load(const ILBits &b, uint32_t k) {
    if (k >= b.nwords) return 0u;
    uint32_t w = __ldg(b.wbase + k);
    if ((k + 1u) * 4u > b.end) { // bytes of the last word past the end of the stream belong to somebody else
        uint32_t keep = b.end - k * 4u; // 1..3
        w &= (1u << (keep * 8u)) - 1u;
    }
    return w;
}
ZWZ_DEV void ilb_feed(ILBits &b) {
    b.hold |= (uint64_t) b.nextw << b.cnt;
    b.cnt += 32u;
    b.widx++;
    b.nextw = ilb_load(b, b.widx);
}
ZWZ_DEV void ilb_refill(ILBits &b) {
    if (b.cnt <= 32u) ilb_feed(b);
}
// position the reader at stream byte `byte_pos` (<= nbytes)
ZWZ_DEV void ilb_seek(ILBits &b, uint32_t byte_pos) {
    uint32_t a = b.skew + byte_pos;
    uint32_t k = a >> 2;
    uint32_t w0 = ilb_load(b, k);
    b.widx = k + 1u;
    b.nextw = ilb_load(b, b.widx);
    uint32_t drop = (a & 3u) * 8u;
    b.hold = (uint64_t) (w0 >> drop);
    b.cnt = 32u - drop;
    b.rem = (b.end - a) * 8u;
}
ZWZ_DEV void ilb_drop(ILBits &b, uint32_t n) {
    b.hold >>= n;
    b.cnt -= n;
    b.rem -= n;
}

// Lane-serial canonical table build (inftrees.c rules). kind: 0 = code-length code (must be complete), 1 = literal/length,
// 2 = distance. Returns 0 ok / 1 invalid set. walk_first/walk_index: state of the canonical walk after `tbits` lengths.
ZWZ_DEV int il_build(uint32_t *W, uint32_t lens0, uint32_t lens_idx0, uint32_t n, uint32_t cnt0, uint32_t sort0, uint32_t lut0, uint32_t tbits,
                     int kind, uint32_t &max_len_out, uint32_t &walk_first, uint32_t &walk_index) {
    for (uint32_t l = 0; l < 8u; ++l) W[(cnt0 + l) * 32u] = 0u;
    for (uint32_t s = 0; s < n; ++s) {
        uint32_t L = IL_U8(W, lens0, lens_idx0 + s);
        IL_U16(W, cnt0, L) = (uint16_t) (IL_U16(W, cnt0, L) + 1u);
    }
    uint32_t max_len = 0;
    int left = 1, bad = 0;
    uint32_t o = 0;
    for (uint32_t l = 1; l <= 15u; ++l) {
        uint32_t c = IL_U16(W, cnt0, l);
        if (c) max_len = l;
        left = (left << 1) - (int) c;
        if (left < 0) bad = 1;
        IL_U16(W, ZWZ_IL_OFF, l) = (uint16_t) o;
        o += c;
    }
    max_len_out = max_len;
    walk_first = 0;
    walk_index = 0;
    if (bad) return 1;
    if (max_len != 0 && left > 0 && (kind == 0 || max_len != 1u)) return 1; // incomplete set
    for (uint32_t s = 0; s < n; ++s) {
        uint32_t L = IL_U8(W, lens0, lens_idx0 + s);
        if (L) {
            uint32_t at = IL_U16(W, ZWZ_IL_OFF, L);
            IL_U16(W, sort0, at) = (uint16_t) s;
            IL_U16(W, ZWZ_IL_OFF, L) = (uint16_t) (at + 1u);
        }
    }
    const uint32_t tsize = 1u << tbits;
    if (left > 0 || max_len == 0u) { // the single-code or empty set: every other slot is zlib's invalid-code entry
        for (uint32_t e = 0; e < tsize; ++e) IL_U16(W, lut0, e) = (uint16_t) ZWZ_IL_INVALID;
    }
    uint32_t code = 0, idx = 0;
    for (uint32_t len = 1; len <= max_len; ++len) {
        uint32_t c = IL_U16(W, cnt0, len);
        if (len == tbits + 1u) {
            walk_first = code;
            walk_index = idx;
        }
        for (uint32_t k = 0; k < c; ++k) {
            uint32_t sym = IL_U16(W, sort0, idx);
            ++idx;
            uint32_t rev = __brev(code) >> (32u - len);
            if (len <= tbits) {
                uint16_t ent = (uint16_t) ((sym << 4) | len);
                for (uint32_t j = rev; j < tsize; j += 1u << len) IL_U16(W, lut0, j) = ent;
            } else {
                IL_U16(W, lut0, rev & (tsize - 1u)) = (uint16_t) ZWZ_IL_LONG;
            }
            ++code;
        }
        code <<= 1;
    }
    return 0;
}

// code longer than the table index: continue the canonical walk from length tbits + 1. Returns the symbol or ~0u.
ZWZ_DEV uint32_t il_walk(const uint32_t *W, uint32_t bits, uint32_t cnt0, uint32_t sort0, uint32_t tbits, uint32_t max_len, uint32_t first,
                         uint32_t index, uint32_t &len_out) {
    uint32_t code = __brev(bits & ((1u << tbits) - 1u)) >> (32u - tbits);
    bits >>= tbits;
    for (uint32_t len = tbits + 1u; len <= max_len; ++len) {
        code = (code << 1) | (bits & 1u);
        bits >>= 1;
        uint32_t c = IL_U16(W, cnt0, len);
        if (code - first < c) {
            len_out = len;
            return IL_U16(W, sort0, index + (code - first));
        }
        index += c;
        first = (first + c) << 1;
    }
    return 0xffffffffu;
}

struct ILState {
    ILBits B;
    uint8_t *out;
    uint32_t cap, pos;
    uint32_t st, status;
    uint32_t last;            // BFINAL of the current block
    uint32_t mlen, dist;      // pending match: bytes left to copy, distance
    uint32_t sbpos, sleft;    // stored block: next source byte, bytes left to copy
    uint32_t slen_after;      // stored block: 1 = the block was cut short by the end of the input
    uint32_t max_ll, max_d, wf_ll, wi_ll, wf_d, wi_d;
    uint32_t a1;              // Adler-32: 1 + sum of bytes (no reduction needed below 2^24 bytes)
    uint64_t a2;              // Adler-32: sum of the running a1
    bool overflow;
};

#define IL_STOP(S, code)      \
    do {                      \
        (S).status = (code);  \
        (S).st = IL_DONE;     \
    } while (0)

// Block header (+ tables for a Huffman block). Lane-serial; runs with whatever lanes are parked here.
ZWZ_DEV void il_block(ILState &S, uint32_t *W, const uint8_t *comp, uint32_t comp_len) {
    ILBits &B = S.B;
    ilb_refill(B);
    if (B.rem < 3u) {
        IL_STOP(S, ZWZ_STREAM_TRUNCATED);
        return;
    }
    S.last = (uint32_t) B.hold & 1u;
    const uint32_t type = ((uint32_t) B.hold >> 1) & 3u;
    ilb_drop(B, 3);
    if (type == 0u) {
        // stored: skip to the byte boundary, LEN/NLEN, then a plain copy
        uint32_t bpos = comp_len - (B.rem >> 3); // rem counts from the end of the stream: floor = next whole byte
        if ((uint64_t) bpos + 4u > comp_len) {
            IL_STOP(S, ZWZ_STREAM_TRUNCATED);
            return;
        }
        ilb_seek(B, bpos);
        ilb_refill(B);
        uint32_t v = (uint32_t) B.hold;
        if ((v & 0xffffu) != ((v >> 16) ^ 0xffffu)) {
            IL_STOP(S, ZWZ_STREAM_BAD);
            return;
        }
        uint32_t len = v & 0xffffu;
        bpos += 4u;
        uint32_t avail = comp_len - bpos;
        uint32_t ncopy = len < avail ? len : avail;
        S.sbpos = bpos;
        S.sleft = ncopy;
        S.slen_after = ncopy < len ? 1u : 0u;
        if (S.pos + ncopy > S.cap) S.overflow = true;
        S.st = IL_STORED;
        return;
    }
    if (type == 3u) {
        IL_STOP(S, ZWZ_STREAM_BAD);
        return;
    }
    uint32_t nlen, ndist;
    if (type == 1u) {
        nlen = 288u;
        ndist = 32u;
        for (uint32_t s = 0; s < 288u; ++s) IL_U8(W, ZWZ_IL_LENS, s) = (uint8_t) (s < 144u ? 8 : (s < 256u ? 9 : (s < 280u ? 7 : 8)));
        for (uint32_t s = 0; s < 32u; ++s) IL_U8(W, ZWZ_IL_LENS, 288u + s) = 5;
    } else {
        if (B.rem < 14u) {
            IL_STOP(S, ZWZ_STREAM_TRUNCATED);
            return;
        }
        nlen = ((uint32_t) B.hold & 31u) + 257u;
        ndist = (((uint32_t) B.hold >> 5) & 31u) + 1u;
        const uint32_t ncode = (((uint32_t) B.hold >> 10) & 15u) + 4u;
        ilb_drop(B, 14);
        if (nlen > 286u || ndist > 30u) {
            IL_STOP(S, ZWZ_STREAM_BAD);
            return;
        }
        for (uint32_t i = 0; i < 5u; ++i) W[(ZWZ_IL_CL_LENS + i) * 32u] = 0u;
        for (uint32_t i = 0; i < ncode; ++i) {
            ilb_refill(B);
            if (B.rem < 3u) {
                IL_STOP(S, ZWZ_STREAM_TRUNCATED);
                return;
            }
            // RFC 1951 §3.2.7 permutation 16 17 18 0 8 7 9 6 10 5 11 4 12 3 13 2 14 1 15
            uint32_t slot = i < 3u ? 16u + i : (i == 3u ? 0u : ((i & 1u) ? 7u - ((i - 5u) >> 1) : 8u + ((i - 4u) >> 1)));
            IL_U8(W, ZWZ_IL_CL_LENS, slot) = (uint8_t) ((uint32_t) B.hold & 7u);
            ilb_drop(B, 3);
        }
        uint32_t max_cl = 0, f0, i0;
        if (il_build(W, ZWZ_IL_CL_LENS, 0, 19u, ZWZ_IL_CL_CNT, ZWZ_IL_CL_SORT, ZWZ_IL_CL_LUT, 7u, 0, max_cl, f0, i0) != 0 || max_cl == 0u) {
            IL_STOP(S, ZWZ_STREAM_BAD);
            return;
        }
        uint32_t have = 0, prev_len = 0;
        const uint32_t total = nlen + ndist;
        while (have < total) {
            ilb_refill(B);
            uint32_t e = IL_U16(W, ZWZ_IL_CL_LUT, (uint32_t) B.hold & 127u);
            uint32_t nb = e & 15u, sym = e >> 4;
            if (e == ZWZ_IL_INVALID || nb == 0u) { // impossible for a complete code; defensive
                IL_STOP(S, ZWZ_STREAM_BAD);
                return;
            }
            uint32_t eb = sym < 16u ? 0u : (sym == 16u ? 2u : (sym == 17u ? 3u : 7u));
            if (B.rem < nb + eb) {
                IL_STOP(S, ZWZ_STREAM_TRUNCATED);
                return;
            }
            ilb_drop(B, nb);
            if (sym < 16u) {
                IL_U8(W, ZWZ_IL_LENS, have) = (uint8_t) sym;
                prev_len = sym;
                have++;
                continue;
            }
            uint32_t rep, val = 0;
            uint32_t x = (uint32_t) B.hold & ((1u << eb) - 1u);
            ilb_drop(B, eb);
            if (sym == 16u) {
                if (have == 0u) {
                    IL_STOP(S, ZWZ_STREAM_BAD);
                    return;
                }
                val = prev_len;
                rep = 3u + x;
            } else if (sym == 17u) {
                rep = 3u + x;
            } else {
                rep = 11u + x;
            }
            if (have + rep > total) {
                IL_STOP(S, ZWZ_STREAM_BAD);
                return;
            }
            for (uint32_t k = 0; k < rep; ++k) IL_U8(W, ZWZ_IL_LENS, have + k) = (uint8_t) val;
            prev_len = val;
            have += rep;
        }
        if (IL_U8(W, ZWZ_IL_LENS, 256u) == 0) { // missing end-of-block code
            IL_STOP(S, ZWZ_STREAM_BAD);
            return;
        }
    }
    if (il_build(W, ZWZ_IL_LENS, 0, nlen, ZWZ_IL_CNT_L, ZWZ_IL_SORT_L, ZWZ_IL_LUT_L, ZWZ_IL_LBITS, 1, S.max_ll, S.wf_ll, S.wi_ll) != 0 ||
        il_build(W, ZWZ_IL_LENS, nlen, ndist, ZWZ_IL_CNT_D, ZWZ_IL_SORT_D, ZWZ_IL_LUT_D, ZWZ_IL_DBITS, 2, S.max_d, S.wf_d, S.wi_d) != 0) {
        IL_STOP(S, ZWZ_STREAM_BAD);
        return;
    }
    S.st = IL_SYM;
}

// RFC 1950 trailer (inflate.c CHECK)
ZWZ_DEV void il_trailer(ILState &S, const uint8_t *comp, uint32_t comp_len, uint32_t flags) {
    uint32_t bpos = comp_len - (S.B.rem >> 3);
    if ((uint64_t) bpos + 4u > comp_len) {
        IL_STOP(S, ZWZ_STREAM_TRUNCATED);
        return;
    }
    if (!(flags & 1u) && !S.overflow) {
        uint32_t want = ((uint32_t) comp[bpos] << 24) | ((uint32_t) comp[bpos + 1] << 16) | ((uint32_t) comp[bpos + 2] << 8) | comp[bpos + 3];
        uint32_t a = S.a1 % 65521u;
        uint32_t b = (uint32_t) (S.a2 % 65521u);
        if (((b << 16) | a) != want) {
            IL_STOP(S, ZWZ_STREAM_BAD);
            return;
        }
    }
    IL_STOP(S, ZWZ_STREAM_END);
}

#define IL_EMIT(S, byte_)                                        \
    do {                                                         \
        uint32_t v_ = (byte_);                                   \
        if ((S).pos < (S).cap) (S).out[(S).pos] = (uint8_t) v_;  \
        (S).a1 += v_;                                            \
        (S).a2 += (S).a1;                                        \
        (S).pos++;                                               \
    } while (0)

// One warp decodes streams order[32g .. 32g+32), one per lane. `order` lists the streams by decreasing compressed size.
ZWZ_KERNEL __launch_bounds__(32) inflate_lanes_kernel(const uint8_t *__restrict__ comp, const uint64_t *__restrict__ off,
                                                     const uint32_t *__restrict__ len, uint8_t *raw_out,
                                                     const uint64_t *__restrict__ raw_off, uint32_t *raw_len, uint32_t *status,
                                                     const uint32_t *__restrict__ order, uint32_t n, uint32_t flags,
                                                     uint32_t *work_counter) {
    ZWZ_DYN_SMEM(smem);
    const unsigned lane = lane_id();
    uint32_t *W = (uint32_t *) smem + lane;
    const uint32_t ngroups = (n + 31u) >> 5;
    for (;;) {
        uint32_t g = 0;
        if (lane == 0) g = atomicAdd(work_counter, 1u);
        g = __shfl_sync(ZWZ_FULL, g, 0);
        if (g >= ngroups) break;
        const uint32_t slot = g * 32u + lane;
        const bool have = slot < n;
        const uint32_t sid = have ? order[slot] : 0u;
        const uint8_t *cp = comp;
        uint32_t comp_len = 0;
        ILState S;
        S.out = raw_out;
        S.cap = 0;
        if (have) {
            cp = comp + off[sid];
            comp_len = len[sid];
            S.out = raw_out + raw_off[sid];
            uint64_t cap64 = raw_off[sid + 1] - raw_off[sid];
            S.cap = cap64 > 0xfffffff0ull ? 0xfffffff0u : (uint32_t) cap64;
        }
        S.pos = 0;
        S.status = ZWZ_STREAM_END;
        S.last = 0;
        S.mlen = S.dist = 0;
        S.sbpos = S.sleft = S.slen_after = 0;
        S.max_ll = S.max_d = S.wf_ll = S.wi_ll = S.wf_d = S.wi_d = 0;
        S.a1 = 1u;
        S.a2 = 0;
        S.overflow = false;
        S.B.skew = (uint32_t) ((uintptr_t) cp & 3u);
        S.B.wbase = (const uint32_t *) (cp - S.B.skew);
        S.B.end = S.B.skew + comp_len;
        S.B.nwords = (S.B.end + 3u) >> 2;
        ilb_seek(S.B, 0);
        S.st = have ? IL_BLOCK : IL_DONE;
        if (have) { // RFC 1950 header (inflate.c HEAD)
            ilb_refill(S.B);
            if (S.B.rem < 16u) {
                IL_STOP(S, ZWZ_STREAM_TRUNCATED);
            } else {
                uint32_t cmf = (uint32_t) S.B.hold & 0xffu, flg = ((uint32_t) S.B.hold >> 8) & 0xffu;
                if (((cmf << 8) + flg) % 31u != 0u || (cmf & 15u) != 8u || (cmf >> 4) + 8u > 15u || (flg & 0x20u))
                    IL_STOP(S, ZWZ_STREAM_BAD);
                else
                    ilb_drop(S.B, 16);
            }
        }
        uint32_t starve = 0;
        for (;;) {
            const unsigned m_slow = __ballot_sync(ZWZ_FULL, S.st <= IL_TRAILER);
            const unsigned m_fast = __ballot_sync(ZWZ_FULL, S.st >= IL_SYM && S.st <= IL_STORED);
            if (!(m_slow | m_fast)) break;
            if (m_slow && (!m_fast || __popc(m_slow) * 4 >= __popc(m_slow | m_fast) || starve >= 16u)) {
                starve = 0;
                if (S.st == IL_BLOCK)
                    il_block(S, W, cp, comp_len);
                else if (S.st == IL_TRAILER)
                    il_trailer(S, cp, comp_len, flags);
                __syncwarp();
                continue;
            }
            starve += m_slow ? 1u : 0u;
            // ---- fast rounds: every lane takes one step of its state per round
            for (int round = 0; round < 8; ++round) {
                ILBits &B = S.B;
                if (S.st == IL_SYM) {
                    ilb_refill(B);
                    uint32_t e = IL_U16(W, ZWZ_IL_LUT_L, (uint32_t) B.hold & ((1u << ZWZ_IL_LBITS) - 1u));
                    uint32_t nb = e & 15u, sym = e >> 4;
                    bool invalid = false;
                    if (e >= ZWZ_IL_INVALID) { // LONG or INVALID
                        if (e == ZWZ_IL_LONG) {
                            uint32_t l2 = 0;
                            sym = il_walk(W, (uint32_t) B.hold, ZWZ_IL_CNT_L, ZWZ_IL_SORT_L, ZWZ_IL_LBITS, S.max_ll, S.wf_ll, S.wi_ll, l2);
                            nb = l2;
                            if (sym == 0xffffffffu) {
                                invalid = true;
                                nb = S.max_ll;
                            }
                        } else {
                            invalid = true;
                            nb = 1u;
                        }
                    }
                    if (!invalid && sym > 285u) invalid = true; // 286/287 of the fixed code: zlib's invalid-code entries
                    if (invalid) {
                        IL_STOP(S, B.rem >= nb ? ZWZ_STREAM_BAD : ZWZ_STREAM_TRUNCATED); // zlib sees the code's bits first
                    } else if (B.rem < nb) {
                        IL_STOP(S, ZWZ_STREAM_TRUNCATED);
                    } else {
                        ilb_drop(B, nb);
                        if (sym < 256u) {
                            IL_EMIT(S, sym);
                        } else if (sym == 256u) {
                            S.st = S.last ? IL_TRAILER : IL_BLOCK;
                        } else {
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
                            if (B.rem < eb) {
                                IL_STOP(S, ZWZ_STREAM_TRUNCATED);
                            } else {
                                S.mlen = base + ((uint32_t) B.hold & ((1u << eb) - 1u));
                                ilb_drop(B, eb);
                                S.st = IL_DIST;
                            }
                        }
                    }
                } else if (S.st == IL_DIST) {
                    ilb_refill(B);
                    uint32_t e = IL_U16(W, ZWZ_IL_LUT_D, (uint32_t) B.hold & ((1u << ZWZ_IL_DBITS) - 1u));
                    uint32_t nb = e & 15u, sym = e >> 4;
                    bool invalid = false;
                    if (e >= ZWZ_IL_INVALID) {
                        if (e == ZWZ_IL_LONG) {
                            uint32_t l2 = 0;
                            sym = il_walk(W, (uint32_t) B.hold, ZWZ_IL_CNT_D, ZWZ_IL_SORT_D, ZWZ_IL_DBITS, S.max_d, S.wf_d, S.wi_d, l2);
                            nb = l2;
                            if (sym == 0xffffffffu) {
                                invalid = true;
                                nb = S.max_d;
                            }
                        } else {
                            invalid = true;
                            nb = 1u;
                        }
                    }
                    if (!invalid && sym > 29u) invalid = true;
                    if (invalid) {
                        IL_STOP(S, B.rem >= nb ? ZWZ_STREAM_BAD : ZWZ_STREAM_TRUNCATED);
                    } else {
                        uint32_t deb, dbase;
                        if (sym < 4u) {
                            deb = 0;
                            dbase = 1u + sym;
                        } else {
                            deb = (sym - 2u) >> 1;
                            dbase = 1u + ((2u + (sym & 1u)) << deb);
                        }
                        if (B.rem < nb + deb) {
                            IL_STOP(S, ZWZ_STREAM_TRUNCATED);
                        } else {
                            ilb_drop(B, nb);
                            S.dist = dbase + ((uint32_t) B.hold & ((1u << deb) - 1u));
                            ilb_drop(B, deb);
                            if (S.dist > S.pos) { // "invalid distance too far back"
                                IL_STOP(S, ZWZ_STREAM_BAD);
                            } else {
                                if (S.pos + S.mlen > S.cap) S.overflow = true;
                                S.st = IL_COPY;
                            }
                        }
                    }
                } else if (S.st == IL_COPY) {
                    const uint32_t nstep = S.mlen < 8u ? S.mlen : 8u;
                    if (S.pos + nstep <= S.cap) {
                        const uint8_t *src = S.out + S.pos - S.dist;
                        uint8_t *dst = S.out + S.pos;
                        if (S.dist >= 8u) { // the 8 loads are independent
                            uint32_t v[8];
#pragma unroll
                            for (uint32_t k = 0; k < 8u; ++k) v[k] = k < nstep ? (uint32_t) src[k] : 0u;
#pragma unroll
                            for (uint32_t k = 0; k < 8u; ++k) {
                                if (k < nstep) {
                                    dst[k] = (uint8_t) v[k];
                                    S.a1 += v[k];
                                    S.a2 += S.a1;
                                }
                            }
                        } else { // overlapping: byte k may be one this step wrote
                            for (uint32_t k = 0; k < nstep; ++k) {
                                uint32_t v = src[k];
                                dst[k] = (uint8_t) v;
                                S.a1 += v;
                                S.a2 += S.a1;
                            }
                        }
                        S.pos += nstep;
                    } else { // crossing the capacity: byte by byte, only what fits is written (and read)
                        for (uint32_t k = 0; k < nstep; ++k) {
                            if (S.pos < S.cap) {
                                uint32_t v = S.out[S.pos - S.dist];
                                S.out[S.pos] = (uint8_t) v;
                            }
                            S.pos++;
                        }
                    }
                    S.mlen -= nstep;
                    if (S.mlen == 0u) S.st = IL_SYM;
                } else if (S.st == IL_STORED) {
                    const uint32_t nstep = S.sleft < 8u ? S.sleft : 8u;
                    for (uint32_t k = 0; k < nstep; ++k) IL_EMIT(S, (uint32_t) cp[S.sbpos + k]);
                    S.sbpos += nstep;
                    S.sleft -= nstep;
                    if (S.sleft == 0u) {
                        if (S.slen_after) {
                            IL_STOP(S, ZWZ_STREAM_TRUNCATED);
                        } else {
                            ilb_seek(B, S.sbpos);
                            S.st = S.last ? IL_TRAILER : IL_BLOCK;
                        }
                    }
                }
                __syncwarp();
            }
        }
        if (have) {
            uint32_t stt = S.status;
            if (stt == ZWZ_STREAM_END && S.overflow) stt = ZWZ_STREAM_OUTPUT_FULL;
            raw_len[sid] = S.pos;
            status[sid] = stt;
        }
        __syncwarp();
    }
}

} // namespace zwz
