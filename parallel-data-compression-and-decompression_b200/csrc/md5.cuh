// md5.cuh — RFC 1321 MD5, one lane per file, batched across files.
//
// Replaces verification.cpp:13-22 (MD5_Init / MD5_Update in 1 024-byte reads / MD5_Final). The digest of one file is
// a serial Merkle–Damgård chain (64 dependent steps per 64-byte block), so parallelism comes only from the number of
// files: config C2 has 370 000 of them. Each lane streams its own file from HBM in 64-byte blocks (two full 32-byte
// sectors per lane per block — uncoalesced across lanes but every fetched sector is fully used).
//
// Algorithmic bytes: N_raw read + 16 B written per file.
#pragma once
#include "zwz_common.cuh"

namespace zwz {

#define ZWZ_MD5_F(b, c, d) ((d) ^ ((b) & ((c) ^ (d))))
#define ZWZ_MD5_G(b, c, d) ((c) ^ ((d) & ((b) ^ (c))))
#define ZWZ_MD5_H(b, c, d) ((b) ^ (c) ^ (d))
#define ZWZ_MD5_I(b, c, d) ((c) ^ ((b) | ~(d)))
// The chain through the freshest word (b) is LOP3 -> add -> LEA.HI (rotate + add): a + m + k does not depend on b and is
// formed first, behind an optimisation barrier. (Left alone the compiler emits IADD3(f, a, m) followed by a separate add
// of k: four dependent ops per step instead of three. ptxas still places the two-input add on the FMA pipe as IMAD.IADD;
// attempts to force an ALU-pipe IADD3 there were re-associated away.)
#ifdef ZWZ_EMU
#define ZWZ_MD5_PIN(x) ((void) 0)
#else
#define ZWZ_MD5_PIN(x) asm("" : "+r"(x))
#endif
#define ZWZ_MD5_STEP(f, a, b, c, d, m, k, s)         \
    do {                                             \
        uint32_t t_ = (a) + (m) + (uint32_t) (k);    \
        ZWZ_MD5_PIN(t_);                             \
        t_ += f((b), (c), (d));                      \
        (a) = __funnelshift_l(t_, t_, (s)) + (b);    \
    } while (0)

ZWZ_DEV void md5_block(uint32_t st[4], const uint32_t m[16]) {
    uint32_t a = st[0], b = st[1], c = st[2], d = st[3];
    ZWZ_MD5_STEP(ZWZ_MD5_F, a, b, c, d, m[0], 0xd76aa478, 7);
    ZWZ_MD5_STEP(ZWZ_MD5_F, d, a, b, c, m[1], 0xe8c7b756, 12);
    ZWZ_MD5_STEP(ZWZ_MD5_F, c, d, a, b, m[2], 0x242070db, 17);
    ZWZ_MD5_STEP(ZWZ_MD5_F, b, c, d, a, m[3], 0xc1bdceee, 22);
    ZWZ_MD5_STEP(ZWZ_MD5_F, a, b, c, d, m[4], 0xf57c0faf, 7);
    ZWZ_MD5_STEP(ZWZ_MD5_F, d, a, b, c, m[5], 0x4787c62a, 12);
    ZWZ_MD5_STEP(ZWZ_MD5_F, c, d, a, b, m[6], 0xa8304613, 17);
    ZWZ_MD5_STEP(ZWZ_MD5_F, b, c, d, a, m[7], 0xfd469501, 22);
    ZWZ_MD5_STEP(ZWZ_MD5_F, a, b, c, d, m[8], 0x698098d8, 7);
    ZWZ_MD5_STEP(ZWZ_MD5_F, d, a, b, c, m[9], 0x8b44f7af, 12);
    ZWZ_MD5_STEP(ZWZ_MD5_F, c, d, a, b, m[10], 0xffff5bb1, 17);
    ZWZ_MD5_STEP(ZWZ_MD5_F, b, c, d, a, m[11], 0x895cd7be, 22);
    ZWZ_MD5_STEP(ZWZ_MD5_F, a, b, c, d, m[12], 0x6b901122, 7);
    ZWZ_MD5_STEP(ZWZ_MD5_F, d, a, b, c, m[13], 0xfd987193, 12);
    ZWZ_MD5_STEP(ZWZ_MD5_F, c, d, a, b, m[14], 0xa679438e, 17);
    ZWZ_MD5_STEP(ZWZ_MD5_F, b, c, d, a, m[15], 0x49b40821, 22);

    ZWZ_MD5_STEP(ZWZ_MD5_G, a, b, c, d, m[1], 0xf61e2562, 5);
    ZWZ_MD5_STEP(ZWZ_MD5_G, d, a, b, c, m[6], 0xc040b340, 9);
    ZWZ_MD5_STEP(ZWZ_MD5_G, c, d, a, b, m[11], 0x265e5a51, 14);
    ZWZ_MD5_STEP(ZWZ_MD5_G, b, c, d, a, m[0], 0xe9b6c7aa, 20);
    ZWZ_MD5_STEP(ZWZ_MD5_G, a, b, c, d, m[5], 0xd62f105d, 5);
    ZWZ_MD5_STEP(ZWZ_MD5_G, d, a, b, c, m[10], 0x02441453, 9);
    ZWZ_MD5_STEP(ZWZ_MD5_G, c, d, a, b, m[15], 0xd8a1e681, 14);
    ZWZ_MD5_STEP(ZWZ_MD5_G, b, c, d, a, m[4], 0xe7d3fbc8, 20);
    ZWZ_MD5_STEP(ZWZ_MD5_G, a, b, c, d, m[9], 0x21e1cde6, 5);
    ZWZ_MD5_STEP(ZWZ_MD5_G, d, a, b, c, m[14], 0xc33707d6, 9);
    ZWZ_MD5_STEP(ZWZ_MD5_G, c, d, a, b, m[3], 0xf4d50d87, 14);
    ZWZ_MD5_STEP(ZWZ_MD5_G, b, c, d, a, m[8], 0x455a14ed, 20);
    ZWZ_MD5_STEP(ZWZ_MD5_G, a, b, c, d, m[13], 0xa9e3e905, 5);
    ZWZ_MD5_STEP(ZWZ_MD5_G, d, a, b, c, m[2], 0xfcefa3f8, 9);
    ZWZ_MD5_STEP(ZWZ_MD5_G, c, d, a, b, m[7], 0x676f02d9, 14);
    ZWZ_MD5_STEP(ZWZ_MD5_G, b, c, d, a, m[12], 0x8d2a4c8a, 20);

    ZWZ_MD5_STEP(ZWZ_MD5_H, a, b, c, d, m[5], 0xfffa3942, 4);
    ZWZ_MD5_STEP(ZWZ_MD5_H, d, a, b, c, m[8], 0x8771f681, 11);
    ZWZ_MD5_STEP(ZWZ_MD5_H, c, d, a, b, m[11], 0x6d9d6122, 16);
    ZWZ_MD5_STEP(ZWZ_MD5_H, b, c, d, a, m[14], 0xfde5380c, 23);
    ZWZ_MD5_STEP(ZWZ_MD5_H, a, b, c, d, m[1], 0xa4beea44, 4);
    ZWZ_MD5_STEP(ZWZ_MD5_H, d, a, b, c, m[4], 0x4bdecfa9, 11);
    ZWZ_MD5_STEP(ZWZ_MD5_H, c, d, a, b, m[7], 0xf6bb4b60, 16);
    ZWZ_MD5_STEP(ZWZ_MD5_H, b, c, d, a, m[10], 0xbebfbc70, 23);
    ZWZ_MD5_STEP(ZWZ_MD5_H, a, b, c, d, m[13], 0x289b7ec6, 4);
    ZWZ_MD5_STEP(ZWZ_MD5_H, d, a, b, c, m[0], 0xeaa127fa, 11);
    ZWZ_MD5_STEP(ZWZ_MD5_H, c, d, a, b, m[3], 0xd4ef3085, 16);
    ZWZ_MD5_STEP(ZWZ_MD5_H, b, c, d, a, m[6], 0x04881d05, 23);
    ZWZ_MD5_STEP(ZWZ_MD5_H, a, b, c, d, m[9], 0xd9d4d039, 4);
    ZWZ_MD5_STEP(ZWZ_MD5_H, d, a, b, c, m[12], 0xe6db99e5, 11);
    ZWZ_MD5_STEP(ZWZ_MD5_H, c, d, a, b, m[15], 0x1fa27cf8, 16);
    ZWZ_MD5_STEP(ZWZ_MD5_H, b, c, d, a, m[2], 0xc4ac5665, 23);

    ZWZ_MD5_STEP(ZWZ_MD5_I, a, b, c, d, m[0], 0xf4292244, 6);
    ZWZ_MD5_STEP(ZWZ_MD5_I, d, a, b, c, m[7], 0x432aff97, 10);
    ZWZ_MD5_STEP(ZWZ_MD5_I, c, d, a, b, m[14], 0xab9423a7, 15);
    ZWZ_MD5_STEP(ZWZ_MD5_I, b, c, d, a, m[5], 0xfc93a039, 21);
    ZWZ_MD5_STEP(ZWZ_MD5_I, a, b, c, d, m[12], 0x655b59c3, 6);
    ZWZ_MD5_STEP(ZWZ_MD5_I, d, a, b, c, m[3], 0x8f0ccc92, 10);
    ZWZ_MD5_STEP(ZWZ_MD5_I, c, d, a, b, m[10], 0xffeff47d, 15);
    ZWZ_MD5_STEP(ZWZ_MD5_I, b, c, d, a, m[1], 0x85845dd1, 21);
    ZWZ_MD5_STEP(ZWZ_MD5_I, a, b, c, d, m[8], 0x6fa87e4f, 6);
    ZWZ_MD5_STEP(ZWZ_MD5_I, d, a, b, c, m[15], 0xfe2ce6e0, 10);
    ZWZ_MD5_STEP(ZWZ_MD5_I, c, d, a, b, m[6], 0xa3014314, 15);
    ZWZ_MD5_STEP(ZWZ_MD5_I, b, c, d, a, m[13], 0x4e0811a1, 21);
    ZWZ_MD5_STEP(ZWZ_MD5_I, a, b, c, d, m[4], 0xf7537e82, 6);
    ZWZ_MD5_STEP(ZWZ_MD5_I, d, a, b, c, m[11], 0xbd3af235, 10);
    ZWZ_MD5_STEP(ZWZ_MD5_I, c, d, a, b, m[2], 0x2ad7d2bb, 15);
    ZWZ_MD5_STEP(ZWZ_MD5_I, b, c, d, a, m[9], 0xeb86d391, 21);
    st[0] += a;
    st[1] += b;
    st[2] += c;
    st[3] += d;
}

// Thread i digests file i. `state` (n*4 words) carries the chaining value across update calls when non-null;
// `finalize` appends RFC 1321 padding using total_len[i] (or len[i] when total_len is null) and writes digest[i].
ZWZ_KERNEL md5_files_kernel(const uint8_t *__restrict__ data, const uint64_t *__restrict__ off, const uint64_t *__restrict__ len,
                            const uint64_t *__restrict__ total_len, uint32_t *state, uint8_t *digest, uint32_t n, int finalize) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t st[4];
    if (state) {
        st[0] = state[4 * i + 0];
        st[1] = state[4 * i + 1];
        st[2] = state[4 * i + 2];
        st[3] = state[4 * i + 3];
    } else {
        st[0] = 0x67452301u;
        st[1] = 0xefcdab89u;
        st[2] = 0x98badcfeu;
        st[3] = 0x10325476u;
    }
    const uint8_t *p = data + off[i];
    uint64_t L = len[i];
    uint64_t nblk = L >> 6;
    // aligned-word view + byte skew: m[k] = bytes [4k+skew, 4k+skew+4) of the aligned stream
    uint32_t skew = (uint32_t) ((uintptr_t) p & 3u);
    const uint32_t *w = (const uint32_t *) (p - skew);
    uint32_t m[16];
    // Software pipeline: the 64 bytes of block b+1 are loaded into registers before the 64 dependent steps of block b run,
    // so the ~1 us of HBM latency of a lane's private stream hides behind ~1 100 cycles of arithmetic. (Lanes of a warp read 32
    // different files: nothing coalesces, every block is two full 32-byte sectors per lane.)
    // One code path for every alignment (a warp's 32 files start at 32 unrelated addresses; a branch on the skew would run
    // the 64-step chain twice per block): word 16 is only fetched when the file is skewed, and a funnel shift by 0 is the
    // identity on its low operand.
    {
        const uint32_t sh = skew * 8u;
        uint32_t nx[17];
        nx[16] = 0;
        if (nblk) {
#pragma unroll
            for (int k = 0; k < 16; ++k) nx[k] = __ldg(w + k);
            if (skew) nx[16] = __ldg(w + 16); // holds >= 1 byte of this block because skew != 0
        }
        for (uint64_t b = 0; b < nblk; ++b) {
#pragma unroll
            for (int k = 0; k < 16; ++k) m[k] = __funnelshift_r(nx[k], nx[k + 1], sh);
            w += 16;
            if (b + 1 < nblk) {
#pragma unroll
                for (int k = 0; k < 16; ++k) nx[k] = __ldg(w + k);
                if (skew) nx[16] = __ldg(w + 16);
            }
            md5_block(st, m);
        }
    }
    if (finalize) {
        const uint8_t *t = p + (nblk << 6);
        uint32_t r = (uint32_t) (L & 63u);
        uint64_t bits = (total_len ? total_len[i] : L) * 8ull;
        // fully unrolled with constant indices so m[] stays in registers (no local-memory indexing)
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            uint32_t v = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                uint32_t idx = 4u * k + j;
                uint32_t byte = idx < r ? (uint32_t) t[idx] : (idx == r ? 0x80u : 0u);
                v |= byte << (8 * j);
            }
            m[k] = v;
        }
        if (r >= 56u) {
            md5_block(st, m);
#pragma unroll
            for (int k = 0; k < 16; ++k) m[k] = 0;
        }
        m[14] = (uint32_t) bits;
        m[15] = (uint32_t) (bits >> 32);
        md5_block(st, m);
        uint32_t *dg = (uint32_t *) (digest + 16ull * i);
        dg[0] = st[0];
        dg[1] = st[1];
        dg[2] = st[2];
        dg[3] = st[3];
    }
    if (state) {
        state[4 * i + 0] = st[0];
        state[4 * i + 1] = st[1];
        state[4 * i + 2] = st[2];
        state[4 * i + 3] = st[3];
    }
}

// ---- large files: same one-lane-per-file chain, but the bytes arrive through shared memory -----------------------------
// With few, long files the per-lane loads above are the problem: 32 lanes x 16 scattered words per block, every block a
// fresh trip to HBM in front of 64 dependent steps. Here a CTA is two warps over the same 32 files: the LOADER warp copies,
// for each file in turn, the next 512 bytes with cp.async (32 lanes x 4 B = one coalesced 128-byte line per instruction) into
// a per-file row of a double-buffered staging area, while the HASHER warp's lanes digest the previous 512 bytes of their own
// rows — nothing but the 64-step chain sits in the hasher's instruction stream. One __syncthreads per round hands the
// buffers over. Row stride = 129 words: lane l reads word k of its row from bank (l + k) mod 32 — conflict-free.
#define ZWZ_MD5S_THREADS 64
#define ZWZ_MD5S_ROUND 512u                     // bytes per file per round (8 MD5 blocks)
#define ZWZ_MD5S_ROW 129u                       // words per row (128 + 1 right neighbour for the funnel shift)
#define ZWZ_MD5S_SMEM (2u * 32u * ZWZ_MD5S_ROW * 4u + 32u * 16u)

#ifdef ZWZ_EMU
ZWZ_DEV void md5s_cp4(uint32_t *dst, const uint32_t *src) { *dst = *src; }
ZWZ_DEV void md5s_commit() {}
ZWZ_DEV void md5s_wait_all() {}
#else
ZWZ_DEV void md5s_cp4(uint32_t *dst, const uint32_t *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t) __cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
ZWZ_DEV void md5s_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
ZWZ_DEV void md5s_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
#endif

// Each row holds the aligned words covering bytes [round*512, round*512 + 512 + 3] of the file's aligned stream.
ZWZ_KERNEL __launch_bounds__(ZWZ_MD5S_THREADS) md5_files_staged_kernel(const uint8_t *__restrict__ data, const uint64_t *__restrict__ off,
                                                                      const uint64_t *__restrict__ len,
                                                                      const uint64_t *__restrict__ total_len, uint32_t *state,
                                                                      uint8_t *digest, uint32_t n, int finalize) {
    ZWZ_DYN_SMEM(smem);
    const unsigned lane = lane_id();
    const bool loader = warp_id() == 1u;
    constexpr uint32_t ROWW = ZWZ_MD5S_ROW; // odd stride: lane l reads word k of its row from bank (l + k) mod 32
    unsigned long long *dir = (unsigned long long *) smem; // the CTA's 32 files: aligned word pointer, readable word count
    uint32_t *stage = (uint32_t *) (smem + 32u * 16u);     // 2 x 32 rows
    const uint32_t i = blockIdx.x * 32u + lane;
    const bool have = i < n;
    uint32_t st[4] = {0x67452301u, 0xefcdab89u, 0x98badcfeu, 0x10325476u};
    if (have && state && !loader) {
        st[0] = state[4 * i + 0];
        st[1] = state[4 * i + 1];
        st[2] = state[4 * i + 2];
        st[3] = state[4 * i + 3];
    }
    const uint8_t *p = have ? data + off[i] : data;
    const uint64_t L = have ? len[i] : 0;
    const uint64_t nblk = L >> 6;
    const uint32_t skew = (uint32_t) ((uintptr_t) p & 3u);
    const uint32_t *w = (const uint32_t *) (p - skew);
    // words of the aligned stream that may be read: covers every full block (+1 neighbour word when skewed)
    const uint64_t nwords = nblk * 16u + (skew && nblk ? 1u : 0u);
    const uint64_t rounds = (nblk + 7u) >> 3;
    uint64_t max_rounds = rounds;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        uint64_t o = __shfl_xor_sync(ZWZ_FULL, max_rounds, d);
        max_rounds = o > max_rounds ? o : max_rounds;
    }
    if (loader) {
        dir[2u * lane] = (unsigned long long) (uintptr_t) w;
        dir[2u * lane + 1u] = nwords;
        __syncwarp();
        // fill buffer r & 1 with round r of all 32 files: per file 4 coalesced 128-byte lines (+ the right neighbour word)
        for (uint64_t r = 0; r < max_rounds; ++r) {
            const uint64_t base = r * 128u;
            const uint32_t buf = (uint32_t) (r & 1u);
#pragma unroll 4
            for (uint32_t f = 0; f < 32u; ++f) {
                const uint32_t *fw = (const uint32_t *) (uintptr_t) dir[2u * f];
                const uint64_t fnw = dir[2u * f + 1u];
                uint32_t *row = stage + ((size_t) buf * 32u + f) * ROWW;
                if (base + 129u <= fnw) { // warp-uniform fast case: the whole row
#pragma unroll
                    for (uint32_t c = 0; c < 4u; ++c) md5s_cp4(row + c * 32u + lane, fw + base + c * 32u + lane);
                    if (lane == 0) md5s_cp4(row + 128u, fw + base + 128u);
                } else if (base < fnw) {
#pragma unroll
                    for (uint32_t c = 0; c < 4u; ++c) {
                        uint64_t k = base + c * 32u + lane;
                        if (k < fnw) md5s_cp4(row + c * 32u + lane, fw + k);
                    }
                }
            }
            md5s_commit();
            md5s_wait_all();
            __syncthreads(); // round r is in shared memory; the hasher has finished with round r - 1's buffer... (see below)
        }
        return;
    }
    // hasher: barrier r releases round r. Buffer r & 1 is refilled (with round r + 2) only after barrier r + 1, which this
    // warp reaches after it has hashed round r.
    const uint32_t sh = skew * 8u; // a funnel shift by 0 returns its low operand: one path for every alignment, no divergence
    for (uint64_t r = 0; r < max_rounds; ++r) {
        __syncthreads();
        if (r < rounds) {
            const uint32_t *row = stage + ((size_t) (r & 1u) * 32u + lane) * ROWW;
            const uint64_t b0 = r * 8u;
            const uint32_t nb = (uint32_t) (nblk - b0 < 8u ? nblk - b0 : 8u);
            for (uint32_t b = 0; b < nb; ++b) {
                uint32_t m[16];
#pragma unroll
                for (int k = 0; k < 16; ++k) m[k] = __funnelshift_r(row[b * 16u + k], row[b * 16u + k + 1], sh);
                md5_block(st, m);
            }
        }
    }
    if (!have) return;
    if (finalize) {
        const uint8_t *t = p + (nblk << 6);
        uint32_t r = (uint32_t) (L & 63u);
        uint64_t bits = (total_len ? total_len[i] : L) * 8ull;
        uint32_t m[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            uint32_t v = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                uint32_t idx = 4u * k + j;
                uint32_t byte = idx < r ? (uint32_t) t[idx] : (idx == r ? 0x80u : 0u);
                v |= byte << (8 * j);
            }
            m[k] = v;
        }
        if (r >= 56u) {
            md5_block(st, m);
#pragma unroll
            for (int k = 0; k < 16; ++k) m[k] = 0;
        }
        m[14] = (uint32_t) bits;
        m[15] = (uint32_t) (bits >> 32);
        md5_block(st, m);
        uint32_t *dg = (uint32_t *) (digest + 16ull * i);
        dg[0] = st[0];
        dg[1] = st[1];
        dg[2] = st[2];
        dg[3] = st[3];
    }
    if (state) {
        state[4 * i + 0] = st[0];
        state[4 * i + 1] = st[1];
        state[4 * i + 2] = st[2];
        state[4 * i + 3] = st[3];
    }
}

} // namespace zwz
