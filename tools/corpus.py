"""Deterministic synthetic corpora for the five BASELINE.json configs.

Test/bench infrastructure — not part of the product package.

Counter-based RNG (numpy Philox) keyed by ``(seed, unit_index)`` so any unit (a file, or a
1 000-file directory of the C2 shape) can be regenerated on its own. Base seed 596.

Content classes:
  T  log/text        timestamp + 6–15 tokens from a 2 010-word vocabulary per line   (zlib ratio ≈ 3)
  S  structured bin  fixed-width records with low-entropy fields                      (≈ 2.5)
  I  image-like      70 % "JPEG-like" (600-B structured header + high-entropy body)   (≈ 1.0)
                     30 % "raw-bitmap-like" (smooth 2-D gradient + ±1 noise)          (≈ 1.8)
  R  uniform random                                                                    (1.0)

Shapes:
  C1  1 000 files, sizes log-uniform on [4 KiB, 16 MiB], mix T45/S25/I20/R10, 3-level tree
  C2  370 000 files in 370 dirs × 1 000, log-normal sizes (median 5 632 B, σ 0.6) clipped to
      [512 B, 60 KiB], all class I; every file is one sub-65 535-byte chunk
  C3  one 16 GiB class-T file
  C5  32 000 files from the C1 generator (seed 597)
All shapes take a ``scale`` so tests can use the same distribution at a few MB.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Iterator, List, Tuple

import numpy as np

BASE_SEED = 596
CHUNK_SIZE = 65535  # process.hpp:12


def _rng(seed: int, unit: int, stream: int = 0) -> np.random.Generator:
    return np.random.Generator(np.random.Philox(key=[(seed << 20) ^ stream, unit]))


# ----------------------------------------------------------------------------------------------
# class T — log text
# ----------------------------------------------------------------------------------------------
_VOCAB_CACHE = {}


def _vocab(seed: int) -> Tuple[np.ndarray, np.ndarray, np.ndarray, np.ndarray]:
    """2 010 pseudo-words. Returns (blob, starts, lengths, cumulative zipf probabilities)."""
    if seed in _VOCAB_CACHE:
        return _VOCAB_CACHE[seed]
    r = _rng(seed, 0xFFFF_0001)
    n = 2010
    lens = r.integers(2, 11, size=n)
    syll = np.frombuffer(b"etaoinshrdlucmfwypvbgkqjxz", dtype=np.uint8)
    # letters drawn with english-like skew
    p = 1.0 / np.arange(1, 27) ** 0.8
    p /= p.sum()
    letters = syll[r.choice(26, size=int(lens.sum()), p=p)]
    # each word is followed by one space in the blob so "word " can be copied in one go
    starts = np.zeros(n, dtype=np.int64)
    blob = np.empty(int(lens.sum()) + n, dtype=np.uint8)
    o = 0
    lo = 0
    for i in range(n):
        starts[i] = o
        blob[o:o + lens[i]] = letters[lo:lo + lens[i]]
        blob[o + lens[i]] = 0x20
        o += lens[i] + 1
        lo += lens[i]
    z = 1.0 / np.arange(1, n + 1) ** 1.05
    cdf = np.cumsum(z / z.sum())
    _VOCAB_CACHE[seed] = (blob, starts, lens + 1, cdf)
    return _VOCAB_CACHE[seed]


def gen_text(size: int, seed: int, unit: int) -> np.ndarray:
    if size == 0:
        return np.zeros(0, dtype=np.uint8)
    blob, starts, wlens, cdf = _vocab(BASE_SEED)
    r = _rng(seed, unit, 1)
    # mean line ≈ 24 (timestamp) + 10.5 tokens × ~7 bytes
    n_lines = size // 80 + 2
    out_parts: List[np.ndarray] = []
    total = 0
    t0 = int(r.integers(1_600_000_000, 1_800_000_000))
    while total < size:
        ntok = r.integers(6, 16, size=n_lines)
        tok = np.searchsorted(cdf, r.random(int(ntok.sum())))
        tok = np.minimum(tok, len(starts) - 1)
        tl = wlens[tok]
        # per-line header "1700000000.123 L " (17 bytes + level char + space)
        dt = np.cumsum(r.integers(0, 3, size=n_lines)) + t0
        t0 = int(dt[-1]) + 1
        ms = r.integers(0, 1000, size=n_lines)
        lvl = np.frombuffer(b"IIIIIIWWED", dtype=np.uint8)[r.integers(0, 10, size=n_lines)]
        hdr = np.empty((n_lines, 17), dtype=np.uint8)
        d = dt.copy()
        for k in range(9, -1, -1):
            hdr[:, k] = 0x30 + d % 10
            d //= 10
        hdr[:, 10] = 0x2E
        m = ms.copy()
        for k in range(13, 10, -1):
            hdr[:, k] = 0x30 + m % 10
            m //= 10
        hdr[:, 14] = 0x20
        hdr[:, 15] = lvl
        hdr[:, 16] = 0x20
        line_body = np.add.reduceat(tl, np.concatenate(([0], np.cumsum(ntok)[:-1])))
        line_len = 17 + line_body  # trailing space of the last token becomes '\n'
        ends = np.cumsum(line_len)
        buf = np.empty(int(ends[-1]), dtype=np.uint8)
        line_start = ends - line_len
        # headers
        hidx = (line_start[:, None] + np.arange(17)[None, :]).ravel()
        buf[hidx] = hdr.ravel()
        # tokens
        tok_line = np.repeat(np.arange(n_lines), ntok)
        tok_off_in_line = np.cumsum(tl) - tl
        first_tok_off = np.repeat(tok_off_in_line[np.concatenate(([0], np.cumsum(ntok)[:-1]))], ntok)
        dst0 = line_start[tok_line] + 17 + (tok_off_in_line - first_tok_off)
        rep = np.repeat(np.arange(len(tok)), tl)
        within = np.arange(int(tl.sum())) - np.repeat(np.cumsum(tl) - tl, tl)
        buf[dst0[rep] + within] = blob[starts[tok][rep] + within]
        buf[ends - 1] = 0x0A
        out_parts.append(buf)
        total += len(buf)
    return np.concatenate(out_parts)[:size]


# ----------------------------------------------------------------------------------------------
# class S — structured binary records
# ----------------------------------------------------------------------------------------------
def gen_struct(size: int, seed: int, unit: int) -> np.ndarray:
    if size == 0:
        return np.zeros(0, dtype=np.uint8)
    r = _rng(seed, unit, 2)
    rec = 40
    n = size // rec + 1
    a = np.zeros((n, rec), dtype=np.uint8)
    seq = np.arange(n, dtype=np.uint64) + np.uint64(r.integers(0, 1 << 30))
    a[:, 0:8] = seq.view(np.uint8).reshape(n, 8)
    ts = (np.cumsum(r.integers(1, 2000, size=n)).astype(np.uint64) + np.uint64(1_700_000_000_000))
    a[:, 8:16] = ts.view(np.uint8).reshape(n, 8)
    kinds = r.integers(0, 12, size=n).astype(np.uint16)
    a[:, 16:18] = kinds.view(np.uint8).reshape(n, 2)
    a[:, 18] = r.integers(0, 4, size=n)
    a[:, 19] = 0
    val = (r.normal(1000.0, 40.0, size=n)).astype(np.float32)
    a[:, 20:24] = val.view(np.uint8).reshape(n, 4)
    q = r.integers(0, 50000, size=n).astype(np.uint32)
    a[:, 24:28] = q.view(np.uint8).reshape(n, 4)
    tag = np.frombuffer(b"ALPHBETAGAMMDELTEPSIZETAETA_THET", dtype=np.uint8).reshape(8, 4)
    a[:, 28:32] = tag[r.integers(0, 8, size=n)]
    a[:, 32:36] = r.integers(0, 256, size=(n, 4))  # one high-entropy field
    a[:, 36:40] = 0
    return a.ravel()[:size].copy()


# ----------------------------------------------------------------------------------------------
# class I — image-like
# ----------------------------------------------------------------------------------------------
_JPEG_HDR = None


def _jpeg_header() -> np.ndarray:
    global _JPEG_HDR
    if _JPEG_HDR is None:
        r = _rng(BASE_SEED, 0xFFFF_0002)
        h = np.zeros(600, dtype=np.uint8)
        h[0:4] = [0xFF, 0xD8, 0xFF, 0xE0]
        h[4:20] = np.frombuffer(b"\x00\x10JFIF\x00\x01\x01\x00\x00\x01\x00\x01\x00\x00", dtype=np.uint8)
        # two quantisation-table-like ramps, huffman-table-like runs
        h[20:24] = [0xFF, 0xDB, 0x00, 0x43]
        h[24:89] = np.minimum(255, 2 + np.arange(65) * 3)
        h[89:93] = [0xFF, 0xDB, 0x00, 0x43]
        h[93:158] = np.minimum(255, 3 + np.arange(65) * 4)
        h[158:162] = [0xFF, 0xC4, 0x01, 0xA2]
        h[162:600] = np.sort(r.integers(0, 256, size=438)).astype(np.uint8)
        _JPEG_HDR = h
    return _JPEG_HDR


def gen_jpeg_like(size: int, seed: int, unit: int) -> np.ndarray:
    r = _rng(seed, unit, 3)
    out = r.integers(0, 256, size=size, dtype=np.uint8)
    k = min(size, 600)
    out[:k] = _jpeg_header()[:k]
    if size > 170:
        # per-file dimensions so headers are similar but not identical
        out[163:167] = r.integers(0, 256, size=4, dtype=np.uint8)
    return out


def gen_bitmap_like(size: int, seed: int, unit: int) -> np.ndarray:
    if size == 0:
        return np.zeros(0, dtype=np.uint8)
    r = _rng(seed, unit, 4)
    w = int(r.integers(48, 200))
    idx = np.arange(size, dtype=np.int64)
    x = (idx // 3) % w
    y = (idx // 3) // w
    c = idx % 3
    gx, gy = r.integers(1, 4), r.integers(1, 4)
    base = r.integers(0, 128)
    v = base + (x * gx) // 2 + (y * gy) // 3 + c * 7 + r.integers(-1, 2, size=size)
    return (v & 0xFF).astype(np.uint8)


def gen_random(size: int, seed: int, unit: int) -> np.ndarray:
    return _rng(seed, unit, 5).integers(0, 256, size=size, dtype=np.uint8)


def gen_file(cls: str, size: int, seed: int, unit: int) -> np.ndarray:
    if cls == "T":
        return gen_text(size, seed, unit)
    if cls == "S":
        return gen_struct(size, seed, unit)
    if cls == "J":
        return gen_jpeg_like(size, seed, unit)
    if cls == "B":
        return gen_bitmap_like(size, seed, unit)
    if cls == "R":
        return gen_random(size, seed, unit)
    if cls == "I":
        return gen_jpeg_like(size, seed, unit) if (_rng(seed, unit, 6).random() < 0.7) else gen_bitmap_like(size, seed, unit)
    raise ValueError(cls)


# ----------------------------------------------------------------------------------------------
# shapes
# ----------------------------------------------------------------------------------------------
@dataclass
class FileSpec:
    relpath: str
    size: int
    cls: str
    unit: int


def c1_specs(n_files: int = 1000, seed: int = BASE_SEED, min_size: int = 4096, max_size: int = 16 << 20) -> List[FileSpec]:
    r = _rng(seed, 0xFFFF_0010)
    sizes = np.exp(r.uniform(np.log(min_size), np.log(max_size), size=n_files)).astype(np.int64)
    u = r.random(n_files)
    cls = np.where(u < 0.45, "T", np.where(u < 0.70, "S", np.where(u < 0.90, "I", "R")))
    ext = {"T": "log", "S": "bin", "I": "img", "R": "rnd"}
    out = []
    for i in range(n_files):
        out.append(FileSpec(f"d{i % 7}/s{(i // 7) % 5}/t{(i // 35) % 3}/f{i:06d}.{ext[cls[i]]}", int(sizes[i]), str(cls[i]), i))
    return out


def c2_sizes(n_files: int, seed: int = BASE_SEED) -> np.ndarray:
    r = _rng(seed, 0xFFFF_0020)
    s = np.exp(r.normal(np.log(5632.0), 0.6, size=n_files))
    return np.clip(s, 512, 60 * 1024).astype(np.int64)


def c2_specs(n_files: int = 370_000, seed: int = BASE_SEED) -> List[FileSpec]:
    sizes = c2_sizes(n_files, seed)
    return [FileSpec(f"dir{i // 1000:03d}/img{i % 1000:04d}.dat", int(sizes[i]), "I", i) for i in range(n_files)]


def c2_buffer(n_files: int = 370_000, seed: int = BASE_SEED, sizes=None) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Whole C2 corpus as ONE contiguous buffer (fast path for bench.py): returns (bytes, offsets[n+1], sizes[n]).

    Keyed per 1 000-file directory: (seed, dir_index). 70 % of files JPEG-like, 30 % raw-bitmap-like. `sizes`: generate the
    files at these sizes (a rank's share of the size-sorted deal) instead of drawing them.
    """
    sizes = c2_sizes(n_files, seed) if sizes is None else np.asarray(sizes, dtype=np.int64)
    n_files = len(sizes)
    offs = np.zeros(n_files + 1, dtype=np.int64)
    np.cumsum(sizes, out=offs[1:])
    buf = np.empty(int(offs[-1]), dtype=np.uint8)
    hdr = _jpeg_header()
    for d0 in range(0, n_files, 1000):
        d1 = min(n_files, d0 + 1000)
        r = _rng(seed, d0 // 1000, 7)
        lo, hi = int(offs[d0]), int(offs[d1])
        seg = buf[lo:hi]
        seg[:] = r.integers(0, 256, size=hi - lo, dtype=np.uint8)
        is_bmp = r.random(d1 - d0) >= 0.7
        fo = offs[d0:d1] - lo
        fs = sizes[d0:d1]
        # JPEG-like headers
        j = np.nonzero(~is_bmp)[0]
        if len(j):
            k = np.minimum(fs[j], 600)
            rep = np.repeat(np.arange(len(j)), k)
            within = np.arange(int(k.sum())) - np.repeat(np.cumsum(k) - k, k)
            seg[fo[j][rep] + within] = hdr[within]
            dims = r.integers(0, 256, size=(len(j), 4), dtype=np.uint8)
            ok = fs[j] > 170
            didx = (fo[j][ok][:, None] + np.arange(163, 167)[None, :]).ravel()
            seg[didx] = dims[ok].ravel()
        # bitmap-like bodies
        b = np.nonzero(is_bmp)[0]
        if len(b):
            k = fs[b]
            rep = np.repeat(np.arange(len(b)), k)
            idx = np.arange(int(k.sum()), dtype=np.int64) - np.repeat(np.cumsum(k) - k, k)
            w = r.integers(48, 200, size=len(b))[rep]
            gx = r.integers(1, 4, size=len(b))[rep]
            gy = r.integers(1, 4, size=len(b))[rep]
            base = r.integers(0, 128, size=len(b))[rep]
            px = idx // 3
            v = base + ((px % w) * gx) // 2 + ((px // w) * gy) // 3 + (idx % 3) * 7 + r.integers(-1, 2, size=len(idx))
            seg[fo[b][rep] + idx] = (v & 0xFF).astype(np.uint8)
    return buf, offs, sizes


def c3_buffer(size: int = 16 << 30, seed: int = BASE_SEED, unit_bytes: int = 8 << 20) -> np.ndarray:
    """Class-T file of `size` bytes, generated in independent `unit_bytes` pieces keyed (seed, piece)."""
    out = np.empty(size, dtype=np.uint8)
    for i, o in enumerate(range(0, size, unit_bytes)):
        n = min(unit_bytes, size - o)
        out[o:o + n] = gen_text(n, seed, 0x10_0000 + i)
    return out


def mixed_specs(total_bytes: int, seed: int = BASE_SEED + 1, min_size: int = 4096, max_size: int = 16 << 20) -> List[FileSpec]:
    """File list of a C1/C5-shaped corpus (no contents): files drawn from the C1 generator until `total_bytes` is reached."""
    specs: List[FileSpec] = []
    tot = 0
    batch = 0
    while tot < total_bytes:
        for s in c1_specs(256, seed + 1000 * batch, min_size, max_size):
            s.unit += batch * 256
            s.relpath = f"b{batch:03d}/" + s.relpath
            specs.append(s)
            tot += s.size
            if tot >= total_bytes:
                break
        batch += 1
    return specs


def mixed_buffer(total_bytes: int, seed: int = BASE_SEED + 1, min_size: int = 4096, max_size: int = 16 << 20):
    """C1/C5-shaped corpus as one buffer: files drawn from the C1 generator until `total_bytes` is reached.
    Returns (bytes, offsets[n+1], specs)."""
    specs = mixed_specs(total_bytes, seed, min_size, max_size)
    offs = np.zeros(len(specs) + 1, dtype=np.int64)
    np.cumsum([s.size for s in specs], out=offs[1:])
    buf = np.empty(int(offs[-1]), dtype=np.uint8)
    for i, s in enumerate(specs):
        buf[offs[i]:offs[i + 1]] = gen_file(s.cls, s.size, seed, s.unit)
    return buf, offs, specs


def write_tree(root: str, specs: List[FileSpec], seed: int = BASE_SEED) -> int:
    total = 0
    for s in specs:
        p = os.path.join(root, s.relpath)
        os.makedirs(os.path.dirname(p), exist_ok=True)
        gen_file(s.cls, s.size, seed, s.unit).tofile(p)
        total += s.size
    return total


# ----------------------------------------------------------------------------------------------
# the reference's chunking rule (compression.cpp:52-64): floor(S/65535)+1 chunks, last may be empty
# ----------------------------------------------------------------------------------------------
def chunk_table(file_offs: np.ndarray) -> Tuple[np.ndarray, np.ndarray, np.ndarray, np.ndarray]:
    """file_offs[n+1] -> (chunk_off[u64], chunk_len[u32], chunk_file[i32], chunk_seq[i32])."""
    sizes = np.diff(file_offs).astype(np.int64)
    nch = sizes // CHUNK_SIZE + 1
    tot = int(nch.sum())
    cfile = np.repeat(np.arange(len(sizes), dtype=np.int64), nch)
    first = np.cumsum(nch) - nch
    seq = np.arange(tot, dtype=np.int64) - first[cfile]
    coff = file_offs[:-1][cfile] + seq * CHUNK_SIZE
    clen = np.minimum(CHUNK_SIZE, sizes[cfile] - seq * CHUNK_SIZE)
    return coff.astype(np.uint64), clen.astype(np.uint32), cfile.astype(np.int32), seq.astype(np.int32)


def edge_case_tree(root: str, seed: int = BASE_SEED) -> List[FileSpec]:
    """SURVEY.md §8(c) regression corpus: empty, 96 B text, exactly 65 535 / 65 536 bytes, 200 000 B text in
    nested dirs, 70 000 random (one full incompressible chunk: the reference truncates it), 60 000 random."""
    specs = [
        FileSpec("empty.bin", 0, "R", 1),
        FileSpec("small96.txt", 96, "T", 2),
        FileSpec("exact65535.txt", 65535, "T", 3),
        FileSpec("exact65536.bin", 65536, "S", 4),
        FileSpec("nest/a/b/text200k.log", 200_000, "T", 5),
        FileSpec("rand70k.bin", 70_000, "R", 6),
        FileSpec("rand60k.bin", 60_000, "R", 7),
        FileSpec("nest/bmp.raw", 100_000, "B", 8),
        FileSpec("one.byte", 1, "T", 9),
    ]
    write_tree(root, specs, seed)
    return specs
