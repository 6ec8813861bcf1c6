/* zwz_oracle.c — CPU ORACLE. TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this. Nothing under parallel-data-compression-and-decompression_b200/ links, imports or
 * executes it; the product path is CUDA-only and fails loudly without its extension.
 *
 * What is restated here, and from where
 * -------------------------------------
 * The reference's hot path is three call sites whose arithmetic lives in two
 * third-party libraries that are NOT vendored in /root/reference and are NOT version-pinned by it:
 *
 *   zlib   (system 1.3 in this image)      compression.cpp:119-134   deflateInit/deflate(Z_FINISH)/deflateEnd
 *                                           decompression.cpp:16-36   inflateInit/inflate loop/inflateEnd
 *   OpenSSL libcrypto (3.0.13 here)        verification.cpp:13-27    MD5_Init/Update/Final + lowercase hex
 *
 * Two kinds of function live in this file:
 *
 *  (1) oracle_ref_*  — the reference's call sequence, verbatim in meaning, against the SAME system
 *      libraries the reference links (-lz -lcrypto). These are "the reference itself" at the seam.
 *  (2) oracle_*      — plain-C restatements of the published algorithms (RFC 1950 zlib wrapper,
 *      RFC 1951 inflate with zlib 1.3's truncated/erroneous-stream output semantics, Adler-32,
 *      RFC 1321 MD5). They have no library dependency and are pinned in tests/test_oracle.py against
 *      (1), against RFC known answers, and against golden vectors produced by oracle/_ref/main_ref
 *      (the unmodified reference built with oracle/stub/mpi.h) — see tests/golden/.
 *
 * Deflate (compression) has no bit-exact contract: BASELINE.json's criteria are that the reference's
 * zlib inflates our streams to the original bytes and that our size is within 3 % of zlib level 6.
 * So the deflate oracle is (1) for the size and {(1),(2)} inflate for the bytes.
 *
 * Parity status: PINNED (RFC vectors + differential against system zlib 1.3 / OpenSSL 3.0.13 + main_ref
 * golden archives). The reference ships no tests of its own.
 */
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#ifndef ZWZ_ORACLE_NO_SYSLIBS
#include <zlib.h>
#include <openssl/evp.h>
#endif

#define ORACLE_CHUNK_SIZE 65535u /* process.hpp:12 */

/* ---------------------------------------------------------------------------------------------
 * status codes shared with include/zwz_cuda.h (ZWZ_STREAM_*)
 * ------------------------------------------------------------------------------------------- */
enum {
    ORACLE_STREAM_END = 0,       /* Z_STREAM_END reached, Adler-32 trailer matched                   */
    ORACLE_STREAM_TRUNCATED = 1, /* input ran out before the end of the stream (zlib: Z_OK/Z_BUF_ERROR) */
    ORACLE_STREAM_BAD = 2,       /* zlib would have returned Z_DATA_ERROR / Z_NEED_DICT                */
    ORACLE_STREAM_OUTPUT_FULL = 3/* more output than out_cap; out_len = bytes that WOULD be produced  */
};

/* =============================================================================================
 * Adler-32 (RFC 1950 §8.2) — the trailer zlib appends at compression.cpp:130 and checks at
 * decompression.cpp:31.
 * =========================================================================================== */
uint32_t oracle_adler32(const uint8_t *p, size_t n) {
    uint32_t a = 1, b = 0;
    while (n) {
        size_t k = n < 5552 ? n : 5552; /* largest run that cannot overflow 32 bits */
        n -= k;
        while (k--) {
            a += *p++;
            b += a;
        }
        a %= 65521u;
        b %= 65521u;
    }
    return (b << 16) | a;
}

/* =============================================================================================
 * MD5 (RFC 1321) — verification.cpp:13-22 (MD5_Init / MD5_Update in 1024-byte reads / MD5_Final).
 * The update granularity does not affect the digest, so this takes the whole buffer.
 * =========================================================================================== */
static const uint32_t md5_k[64] = {
    0xd76aa478, 0xe8c7b756, 0x242070db, 0xc1bdceee, 0xf57c0faf, 0x4787c62a, 0xa8304613, 0xfd469501,
    0x698098d8, 0x8b44f7af, 0xffff5bb1, 0x895cd7be, 0x6b901122, 0xfd987193, 0xa679438e, 0x49b40821,
    0xf61e2562, 0xc040b340, 0x265e5a51, 0xe9b6c7aa, 0xd62f105d, 0x02441453, 0xd8a1e681, 0xe7d3fbc8,
    0x21e1cde6, 0xc33707d6, 0xf4d50d87, 0x455a14ed, 0xa9e3e905, 0xfcefa3f8, 0x676f02d9, 0x8d2a4c8a,
    0xfffa3942, 0x8771f681, 0x6d9d6122, 0xfde5380c, 0xa4beea44, 0x4bdecfa9, 0xf6bb4b60, 0xbebfbc70,
    0x289b7ec6, 0xeaa127fa, 0xd4ef3085, 0x04881d05, 0xd9d4d039, 0xe6db99e5, 0x1fa27cf8, 0xc4ac5665,
    0xf4292244, 0x432aff97, 0xab9423a7, 0xfc93a039, 0x655b59c3, 0x8f0ccc92, 0xffeff47d, 0x85845dd1,
    0x6fa87e4f, 0xfe2ce6e0, 0xa3014314, 0x4e0811a1, 0xf7537e82, 0xbd3af235, 0x2ad7d2bb, 0xeb86d391};
static const uint8_t md5_s[64] = {7, 12, 17, 22, 7, 12, 17, 22, 7, 12, 17, 22, 7, 12, 17, 22,
                                  5, 9,  14, 20, 5, 9,  14, 20, 5, 9,  14, 20, 5, 9,  14, 20,
                                  4, 11, 16, 23, 4, 11, 16, 23, 4, 11, 16, 23, 4, 11, 16, 23,
                                  6, 10, 15, 21, 6, 10, 15, 21, 6, 10, 15, 21, 6, 10, 15, 21};

static void md5_block(uint32_t st[4], const uint8_t blk[64]) {
    uint32_t m[16];
    for (int i = 0; i < 16; ++i)
        m[i] = (uint32_t) blk[4 * i] | ((uint32_t) blk[4 * i + 1] << 8) | ((uint32_t) blk[4 * i + 2] << 16) |
               ((uint32_t) blk[4 * i + 3] << 24);
    uint32_t a = st[0], b = st[1], c = st[2], d = st[3];
    for (int i = 0; i < 64; ++i) {
        uint32_t f;
        int g;
        if (i < 16) {
            f = (b & c) | (~b & d);
            g = i;
        } else if (i < 32) {
            f = (d & b) | (~d & c);
            g = (5 * i + 1) & 15;
        } else if (i < 48) {
            f = b ^ c ^ d;
            g = (3 * i + 5) & 15;
        } else {
            f = c ^ (b | ~d);
            g = (7 * i) & 15;
        }
        uint32_t t = a + f + md5_k[i] + m[g];
        a = d;
        d = c;
        c = b;
        b = b + ((t << md5_s[i]) | (t >> (32 - md5_s[i])));
    }
    st[0] += a;
    st[1] += b;
    st[2] += c;
    st[3] += d;
}

void oracle_md5(const uint8_t *p, uint64_t n, uint8_t digest[16]) {
    uint32_t st[4] = {0x67452301u, 0xefcdab89u, 0x98badcfeu, 0x10325476u};
    uint64_t full = n / 64;
    for (uint64_t i = 0; i < full; ++i) md5_block(st, p + 64 * i);
    uint8_t tail[128];
    size_t r = (size_t) (n - 64 * full);
    memset(tail, 0, sizeof tail);
    if (r) memcpy(tail, p + 64 * full, r);
    tail[r] = 0x80;
    size_t tl = (r < 56) ? 64 : 128;
    uint64_t bits = n * 8;
    for (int i = 0; i < 8; ++i) tail[tl - 8 + i] = (uint8_t) (bits >> (8 * i));
    md5_block(st, tail);
    if (tl == 128) md5_block(st, tail + 64);
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) digest[4 * i + j] = (uint8_t) (st[i] >> (8 * j));
}

/* verification.cpp:24-27 — 32 lowercase hex characters, no terminator written past hex[31]. */
void oracle_md5_hex(const uint8_t *p, uint64_t n, char hex[32]) {
    static const char d[] = "0123456789abcdef";
    uint8_t dg[16];
    oracle_md5(p, n, dg);
    for (int i = 0; i < 16; ++i) {
        hex[2 * i] = d[dg[i] >> 4];
        hex[2 * i + 1] = d[dg[i] & 15];
    }
}

/* =============================================================================================
 * Inflate — RFC 1950 wrapper + RFC 1951 body, with the output semantics of zlib 1.3's inflate()
 * as driven by decompression.cpp:24-34 (return codes ignored, everything produced is written):
 *
 *   - a literal is produced iff its whole Huffman code lies inside the available input bits;
 *   - a match is produced iff length code + length extra + distance code + distance extra are ALL
 *     available (no partial copy) — zlib's LEN/LENEXT/DIST/DISTEXT/MATCH state chain;
 *   - a stored block copies min(LEN, bytes left in input) — zlib's COPY state;
 *   - a dynamic-block header yields nothing until it is complete and valid;
 *   - on a data error everything produced BEFORE the offending symbol stays.
 *
 * Error rules follow zlib's inflate.c / inftrees.c (over-subscribed sets, incomplete sets unless the
 * set is a single 1-bit code, missing end-of-block code, repeat with no previous length, nlen > 286,
 * ndist > 30, distance further back than the bytes produced so far, literal/length symbols 286/287
 * and distance symbols 30/31, stored LEN != ~NLEN, block type 3, header checks, FDICT => no output).
 *
 * out_cap: bytes available at `out`. Decoding continues past it WITHOUT storing (so *out_len is the
 * full size and the caller can retry), status ORACLE_STREAM_OUTPUT_FULL — unless a later error or
 * truncation changes the status; then *out_len is still the full count and the status says why it ended.
 * Back-references past out_cap are resolved correctly only while they fall inside `out`; beyond that the
 * byte VALUES are not needed because nothing is stored, only counted.
 * =========================================================================================== */
typedef struct {
    const uint8_t *in;
    size_t in_len, in_pos;
    uint64_t hold;
    unsigned bits;
} bitsrc;

/* make up to `need` bits available; returns 0 if the input ran out first (hold keeps what it has) */
static int bs_need(bitsrc *s, unsigned need) {
    while (s->bits < need) {
        if (s->in_pos >= s->in_len) return 0;
        s->hold |= (uint64_t) s->in[s->in_pos++] << s->bits;
        s->bits += 8;
    }
    return 1;
}
static inline uint32_t bs_peek(const bitsrc *s, unsigned n) { return (uint32_t) (s->hold & ((1ull << n) - 1)); }
static inline void bs_drop(bitsrc *s, unsigned n) {
    s->hold >>= n;
    s->bits -= n;
}

typedef struct {
    uint16_t count[16];  /* codes per length */
    uint16_t symbol[288];/* symbols ordered by code */
    int max_len;         /* 0 = no codes */
    int incomplete;      /* single 1-bit code (allowed for litlen/dist) */
} huff;

/* returns 0 ok, -1 invalid (mirrors inflate_table's accept/reject decisions).
 * kind: 0 = code-length codes (must be complete), 1 = litlen, 2 = dist. */
static int huff_build(huff *h, const uint8_t *lens, int n, int kind) {
    uint16_t offs[16];
    memset(h->count, 0, sizeof h->count);
    for (int i = 0; i < n; ++i) h->count[lens[i]]++;
    h->max_len = 0;
    for (int l = 15; l >= 1; --l)
        if (h->count[l]) {
            h->max_len = l;
            break;
        }
    h->incomplete = 0;
    if (h->max_len == 0) return 0; /* "no symbols to code at all" is accepted by inflate_table */
    int left = 1;
    for (int l = 1; l <= 15; ++l) {
        left <<= 1;
        left -= h->count[l];
        if (left < 0) return -1; /* over-subscribed */
    }
    if (left > 0) {
        if (kind == 0 || h->max_len != 1) return -1; /* incomplete set */
        h->incomplete = 1;
    }
    offs[1] = 0;
    for (int l = 1; l < 15; ++l) offs[l + 1] = offs[l] + h->count[l];
    for (int i = 0; i < n; ++i)
        if (lens[i]) h->symbol[offs[lens[i]]++] = (uint16_t) i;
    return 0;
}

/* Decode one symbol. Returns symbol >= 0; -1 = not enough input (nothing consumed);
 * -2 = bit pattern that maps to no code (possible only for empty / single-code sets). */
static int huff_decode(bitsrc *s, const huff *h) {
    if (h->max_len == 0) {
        /* zlib's table for an empty set is all "invalid code" entries of length 1 */
        if (!bs_need(s, 1)) return -1;
        return -2;
    }
    int code = 0, first = 0, index = 0;
    for (int len = 1; len <= h->max_len; ++len) {
        if (!bs_need(s, (unsigned) len)) return -1;
        code |= (int) ((s->hold >> (len - 1)) & 1);
        int cnt = h->count[len];
        if (code - cnt < first) {
            bs_drop(s, (unsigned) len);
            return h->symbol[index + (code - first)];
        }
        index += cnt;
        first += cnt;
        first <<= 1;
        code <<= 1;
    }
    return -2; /* incomplete single-code set, the unused pattern */
}

static const uint16_t len_base[29] = {3,  4,  5,  6,  7,  8,  9,  10, 11,  13,  15,  17,  19,  23, 27,
                                      31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
static const uint8_t len_extra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
static const uint16_t dist_base[30] = {1,   2,   3,   4,   5,   7,    9,    13,   17,   25,   33,   49,   65,    97,    129,
                                       193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
static const uint8_t dist_extra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
static const uint8_t clc_order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

int oracle_inflate(const uint8_t *in, size_t in_len, uint8_t *out, size_t out_cap, uint64_t *out_len) {
    bitsrc s = {in, in_len, 0, 0, 0};
    uint64_t n = 0; /* bytes produced (counted even past out_cap) */
    int overflow = 0;
#define RET(st)                                                   \
    do {                                                          \
        *out_len = n;                                             \
        return (overflow && (st) == ORACLE_STREAM_END) ? ORACLE_STREAM_OUTPUT_FULL : (st); \
    } while (0)
#define RETX(st)                                                  \
    do {                                                          \
        *out_len = n;                                             \
        return (st);                                              \
    } while (0)

    /* RFC 1950 header — inflate.c HEAD state */
    if (!bs_need(&s, 16)) RETX(ORACLE_STREAM_TRUNCATED);
    {
        uint32_t cmf = bs_peek(&s, 8), flg = (uint32_t) (s.hold >> 8) & 0xff;
        if (((cmf << 8) + flg) % 31) RETX(ORACLE_STREAM_BAD);
        if ((cmf & 15) != 8) RETX(ORACLE_STREAM_BAD);
        if ((cmf >> 4) + 8 > 15) RETX(ORACLE_STREAM_BAD);
        if (flg & 0x20) RETX(ORACLE_STREAM_BAD); /* FDICT: Z_NEED_DICT, no output */
        bs_drop(&s, 16);
    }

    static huff fixed_ll, fixed_d;
    static int fixed_ready = 0;
    if (!fixed_ready) {
        uint8_t l[288];
        int i = 0;
        for (; i < 144; ++i) l[i] = 8;
        for (; i < 256; ++i) l[i] = 9;
        for (; i < 280; ++i) l[i] = 7;
        for (; i < 288; ++i) l[i] = 8;
        huff_build(&fixed_ll, l, 288, 1);
        for (i = 0; i < 32; ++i) l[i] = 5;
        huff_build(&fixed_d, l, 32, 2);
        fixed_ready = 1;
    }

    for (;;) {
        if (!bs_need(&s, 3)) RET(ORACLE_STREAM_TRUNCATED);
        int last = (int) bs_peek(&s, 1);
        int type = (int) (bs_peek(&s, 3) >> 1);
        bs_drop(&s, 3);
        huff dyn_ll, dyn_d;
        const huff *ll, *dd;
        if (type == 0) {
            bs_drop(&s, s.bits & 7);
            if (!bs_need(&s, 32)) RET(ORACLE_STREAM_TRUNCATED);
            uint32_t v = bs_peek(&s, 32);
            if ((v & 0xffff) != ((v >> 16) ^ 0xffff)) RET(ORACLE_STREAM_BAD);
            bs_drop(&s, 32);
            uint32_t len = v & 0xffff;
            /* whole bytes still in hold (none in practice after the aligned 32-bit drop) come first */
            while (len) {
                uint8_t c;
                if (s.bits >= 8) {
                    c = (uint8_t) bs_peek(&s, 8);
                    bs_drop(&s, 8);
                } else {
                    if (s.in_pos >= s.in_len) RET(ORACLE_STREAM_TRUNCATED);
                    c = s.in[s.in_pos++];
                }
                if (n < out_cap) out[n] = c; else overflow = 1;
                ++n;
                --len;
            }
            if (last) break;
            continue;
        } else if (type == 1) {
            ll = &fixed_ll;
            dd = &fixed_d;
        } else if (type == 2) {
            if (!bs_need(&s, 14)) RET(ORACLE_STREAM_TRUNCATED);
            int nlen = (int) bs_peek(&s, 5) + 257;
            bs_drop(&s, 5);
            int ndist = (int) bs_peek(&s, 5) + 1;
            bs_drop(&s, 5);
            int ncode = (int) bs_peek(&s, 4) + 4;
            bs_drop(&s, 4);
            if (nlen > 286 || ndist > 30) RET(ORACLE_STREAM_BAD);
            uint8_t lens[320];
            memset(lens, 0, sizeof lens);
            for (int i = 0; i < ncode; ++i) {
                if (!bs_need(&s, 3)) RET(ORACLE_STREAM_TRUNCATED);
                lens[clc_order[i]] = (uint8_t) bs_peek(&s, 3);
                bs_drop(&s, 3);
            }
            huff clc;
            if (huff_build(&clc, lens, 19, 0) != 0) RET(ORACLE_STREAM_BAD);
            if (clc.max_len == 0) RET(ORACLE_STREAM_BAD); /* CODES with no codes: inflate_table's
                 empty table makes the very first code-length lookup an invalid code ... which zlib
                 reports as "invalid bit length repeat"-class data error: no output either way */
            memset(lens, 0, sizeof lens);
            int have = 0;
            while (have < nlen + ndist) {
                /* zlib peeks code+extra together (NEEDBITS(here.bits + n)) before consuming either */
                bitsrc save = s;
                int sym = huff_decode(&s, &clc);
                if (sym == -1) RET(ORACLE_STREAM_TRUNCATED);
                if (sym < 0) RET(ORACLE_STREAM_BAD);
                if (sym < 16) {
                    lens[have++] = (uint8_t) sym;
                    continue;
                }
                unsigned eb = sym == 16 ? 2 : (sym == 17 ? 3 : 7);
                if (!bs_need(&s, eb)) {
                    s = save;
                    RET(ORACLE_STREAM_TRUNCATED);
                }
                int rep, val = 0;
                if (sym == 16) {
                    if (have == 0) RET(ORACLE_STREAM_BAD);
                    val = lens[have - 1];
                    rep = 3 + (int) bs_peek(&s, 2);
                } else if (sym == 17) {
                    rep = 3 + (int) bs_peek(&s, 3);
                } else {
                    rep = 11 + (int) bs_peek(&s, 7);
                }
                bs_drop(&s, eb);
                if (have + rep > nlen + ndist) RET(ORACLE_STREAM_BAD);
                while (rep--) lens[have++] = (uint8_t) val;
            }
            if (lens[256] == 0) RET(ORACLE_STREAM_BAD); /* missing end-of-block */
            if (huff_build(&dyn_ll, lens, nlen, 1) != 0) RET(ORACLE_STREAM_BAD);
            if (huff_build(&dyn_d, lens + nlen, ndist, 2) != 0) RET(ORACLE_STREAM_BAD);
            ll = &dyn_ll;
            dd = &dyn_d;
        } else {
            RET(ORACLE_STREAM_BAD);
        }

        for (;;) {
            bitsrc save = s; /* a match is all-or-nothing: rewind if any later part is missing */
            int sym = huff_decode(&s, ll);
            if (sym == -1) RET(ORACLE_STREAM_TRUNCATED);
            if (sym < 0) RET(ORACLE_STREAM_BAD);
            if (sym < 256) {
                if (n < out_cap) out[n] = (uint8_t) sym; else overflow = 1;
                ++n;
                continue;
            }
            if (sym == 256) break;
            if (sym > 285) RET(ORACLE_STREAM_BAD);
            sym -= 257;
            unsigned len = len_base[sym];
            if (len_extra[sym]) {
                if (!bs_need(&s, len_extra[sym])) {
                    s = save;
                    RET(ORACLE_STREAM_TRUNCATED);
                }
                len += bs_peek(&s, len_extra[sym]);
                bs_drop(&s, len_extra[sym]);
            }
            int ds = huff_decode(&s, dd);
            if (ds == -1) {
                s = save;
                RET(ORACLE_STREAM_TRUNCATED);
            }
            if (ds < 0 || ds > 29) RET(ORACLE_STREAM_BAD);
            unsigned dist = dist_base[ds];
            if (dist_extra[ds]) {
                if (!bs_need(&s, dist_extra[ds])) {
                    s = save;
                    RET(ORACLE_STREAM_TRUNCATED);
                }
                dist += bs_peek(&s, dist_extra[ds]);
                bs_drop(&s, dist_extra[ds]);
            }
            if (dist > n) RET(ORACLE_STREAM_BAD); /* "invalid distance too far back" */
            for (unsigned k = 0; k < len; ++k) {
                if (n < out_cap) out[n] = out[n - dist]; else overflow = 1;
                ++n;
            }
        }
        if (last) break;
    }
    /* RFC 1950 trailer — inflate.c CHECK state: NEEDBITS(32) keeps whatever whole bytes are still in
     * `hold` after the final block (zlib only discards bits & 7 here). */
    bs_drop(&s, s.bits & 7);
    if (!bs_need(&s, 32)) RET(ORACLE_STREAM_TRUNCATED);
    {
        uint32_t v = bs_peek(&s, 32);
        uint32_t want = ((v & 0xff) << 24) | ((v & 0xff00) << 8) | ((v >> 8) & 0xff00) | (v >> 24);
        if (!overflow) {
            uint32_t got = oracle_adler32(out, (size_t) n);
            if (got != want) RET(ORACLE_STREAM_BAD);
        }
    }
    RET(ORACLE_STREAM_END);
#undef RET
#undef RETX
}

#ifndef ZWZ_ORACLE_NO_SYSLIBS
/* =============================================================================================
 * (1) The reference's call sequences against the reference's own libraries.
 * =========================================================================================== */

/* compression.cpp:119-134. `out` is the reference's 65 535-byte buffer; the return value is its
 * `compressed_size = CHUNK_SIZE - strm.avail_out`. Return codes are ignored exactly as there, so an
 * incompressible 65 535-byte chunk yields a stream truncated to 65 535 bytes. */
long oracle_ref_deflate_chunk(const uint8_t *chunk, size_t size, uint8_t out[ORACLE_CHUNK_SIZE]) {
    z_stream strm;
    strm.zalloc = Z_NULL;
    strm.zfree = Z_NULL;
    strm.opaque = Z_NULL;
    deflateInit(&strm, Z_DEFAULT_COMPRESSION);
    strm.avail_in = (uInt) size;
    strm.next_in = (Bytef *) chunk;
    strm.avail_out = ORACLE_CHUNK_SIZE;
    strm.next_out = out;
    deflate(&strm, Z_FINISH);
    long compressed_size = (long) ORACLE_CHUNK_SIZE - (long) strm.avail_out;
    deflateEnd(&strm);
    return compressed_size;
}

/* Same zlib parameters with room for the worst case: the size zlib level 6 WOULD need. This is the
 * denominator of the "<= 3 % of zlib" ratio criterion. */
long oracle_ref_deflate_bound_size(const uint8_t *chunk, size_t size, int level) {
    uint8_t buf[ORACLE_CHUNK_SIZE + 1024];
    z_stream strm;
    memset(&strm, 0, sizeof strm);
    deflateInit(&strm, level);
    strm.avail_in = (uInt) size;
    strm.next_in = (Bytef *) chunk;
    strm.avail_out = sizeof buf;
    strm.next_out = buf;
    int rc = deflate(&strm, Z_FINISH);
    long n = (long) sizeof buf - (long) strm.avail_out;
    deflateEnd(&strm);
    return rc == Z_STREAM_END ? n : -1;
}

/* decompression.cpp:11-37 with `avail_in` made explicit. The reference always passes 65 535 (the
 * std::array's size) and relies on zlib stopping at end-of-stream; for every record the reference
 * itself can write, passing the true payload length gives the same bytes (a truncated record is
 * exactly 65 535 long). Output is appended to `dest`; returns total bytes written, or -1 if dest_cap
 * would be exceeded. */
long long oracle_ref_inflate_chunk(const uint8_t *data, size_t avail_in, uint8_t *dest, size_t dest_cap) {
    z_stream strm;
    strm.zalloc = Z_NULL;
    strm.zfree = Z_NULL;
    strm.opaque = Z_NULL;
    if (inflateInit(&strm) != Z_OK) return 0;
    strm.avail_in = (uInt) avail_in;
    strm.next_in = (Bytef *) data;
    unsigned char out[ORACLE_CHUNK_SIZE];
    long long total = 0;
    do {
        strm.avail_out = sizeof(out);
        strm.next_out = out;
        inflate(&strm, Z_NO_FLUSH);
        size_t have = sizeof(out) - strm.avail_out;
        if ((size_t) total + have > dest_cap) {
            inflateEnd(&strm);
            return -1;
        }
        memcpy(dest + total, out, have);
        total += (long long) have;
    } while (strm.avail_out == 0);
    inflateEnd(&strm);
    return total;
}

/* verification.cpp:13-27 over an in-memory buffer (EVP because MD5_* is deprecated in OpenSSL 3;
 * same digest). */
void oracle_ref_md5_hex(const uint8_t *p, uint64_t n, char hex[32]) {
    static const char d[] = "0123456789abcdef";
    unsigned char dg[EVP_MAX_MD_SIZE];
    unsigned int dl = 0;
    EVP_MD_CTX *ctx = EVP_MD_CTX_new();
    EVP_DigestInit_ex(ctx, EVP_md5(), NULL);
    for (uint64_t o = 0; o < n; o += 1024) { /* verification.cpp:15-19 reads 1024 bytes at a time */
        size_t k = (n - o) < 1024 ? (size_t) (n - o) : 1024;
        EVP_DigestUpdate(ctx, p + o, k);
    }
    EVP_DigestFinal_ex(ctx, dg, &dl);
    EVP_MD_CTX_free(ctx);
    for (int i = 0; i < 16; ++i) {
        hex[2 * i] = d[dg[i] >> 4];
        hex[2 * i + 1] = d[dg[i] & 15];
    }
}

/* ---- batch drivers for bench.py's cpu_baseline / --impl reference leg (one host thread each;
 *      bench.py fans them out over processes). Return bytes produced. ---- */
uint64_t oracle_ref_deflate_batch(const uint8_t *raw, const uint64_t *off, const uint32_t *len, uint32_t n,
                                  uint8_t *out /* n slots of 65535 */, uint32_t *out_len) {
    uint64_t tot = 0;
    for (uint32_t i = 0; i < n; ++i) {
        long c = oracle_ref_deflate_chunk(raw + off[i], len[i], out + (size_t) i * ORACLE_CHUNK_SIZE);
        out_len[i] = (uint32_t) c;
        tot += (uint64_t) c;
    }
    return tot;
}
uint64_t oracle_ref_inflate_batch(const uint8_t *comp /* n slots of 65535 */, const uint32_t *comp_len, uint32_t n,
                                  uint8_t *raw_out, const uint64_t *raw_off, uint32_t *raw_len) {
    uint64_t tot = 0;
    for (uint32_t i = 0; i < n; ++i) {
        long long r = oracle_ref_inflate_chunk(comp + (size_t) i * ORACLE_CHUNK_SIZE, comp_len[i], raw_out + raw_off[i],
                                               (size_t) (raw_off[i + 1] - raw_off[i]));
        raw_len[i] = r < 0 ? 0xffffffffu : (uint32_t) r;
        if (r > 0) tot += (uint64_t) r;
    }
    return tot;
}
void oracle_ref_md5_batch(const uint8_t *data, const uint64_t *off, const uint64_t *len, uint32_t n, char *hex /* n*32 */) {
    for (uint32_t i = 0; i < n; ++i) oracle_ref_md5_hex(data + off[i], len[i], hex + 32 * (size_t) i);
}
#endif /* ZWZ_ORACLE_NO_SYSLIBS */
