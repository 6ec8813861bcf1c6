"""CPU fuzz of the inflate path under the SIMT emulator: zlib streams (all levels and strategies, small and large windows) and
our own streams of chunks assembled from mixed segments, whole and truncated at random points; bytes, count and status must
equal the oracle's (== zlib's behaviour as decompression.cpp drives it).   python tools/fuzz_inflate_emu.py [rounds] [seed]"""
import os, sys, zlib
import numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
os.environ.setdefault("ZWZ_EMU_POISON", "9")
import emu_lib
import oracle_lib as O
from tools import corpus

def run(rounds, seed, chunks_per_round=None):
    """returns True when every stream checked out"""
    rng = np.random.default_rng(seed)
    ctx = emu_lib.emu_context()
    text = corpus.gen_text(200000, 596, 7).tobytes()

    def chunk():
        c = b""
        target = int(rng.choice([rng.integers(0, 300), rng.integers(300, 5000), rng.integers(5000, 40000), 65535]))
        while len(c) < target:
            k = rng.integers(0, 6); n = int(rng.integers(1, 6000))
            if k == 0: c += rng.integers(0, 256, n, dtype=np.uint8).tobytes()
            elif k == 1: o = int(rng.integers(0, len(text) - n)); c += text[o:o + n]
            elif k == 2 and c: o = int(rng.integers(0, len(c))); d = c[o:o + n]; c += (d * (n // len(d) + 1))[:n]
            elif k == 3: c += bytes([int(rng.integers(0, 256))]) * n
            elif k == 4: c += (rng.integers(0, 4, n, dtype=np.uint8) * 60).tobytes()
            else: c += bytes((np.arange(n) & 255).astype(np.uint8))
        return c[:target]

    total = 0
    for r in range(rounds):
        chunks = [chunk() for _ in range(24)]
        streams = []
        for c in chunks:
            lvl = int(rng.integers(0, 10)); strat = int(rng.choice([zlib.Z_DEFAULT_STRATEGY, zlib.Z_FILTERED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE, zlib.Z_FIXED]))
            co = zlib.compressobj(lvl, zlib.DEFLATED, int(rng.choice([9, 12, 15])), int(rng.integers(1, 10)), strat)
            s = co.compress(c) + co.flush()
            streams.append(s)
            if len(s) > 8 and rng.random() < 0.4:
                streams.append(s[:int(rng.integers(2, len(s)))])          # truncated
            if len(s) > 12 and rng.random() < 0.2:
                b = bytearray(s); b[int(rng.integers(2, len(s)))] ^= 1 << int(rng.integers(0, 8)); streams.append(bytes(b))   # one bit flipped
        raw, off = b"".join(chunks), None
        lens = np.array([len(c) for c in chunks], dtype=np.uint32); o = np.zeros(len(chunks), dtype=np.uint64); o[1:] = np.cumsum(lens)[:-1]
        packed, poff, res = ctx.deflate_batch(np.frombuffer(raw, dtype=np.uint8) if raw else np.zeros(0, np.uint8), o, lens)
        for i in range(len(chunks)):
            if res["len1"][i] == 0:
                streams.append(packed[int(poff[i]):int(poff[i + 1])].tobytes())
        so = np.zeros(len(streams) + 1, dtype=np.uint64); np.cumsum([len(s) for s in streams], out=so[1:])
        ro = np.arange(len(streams) + 1, dtype=np.uint64) * np.uint64(70000)
        out, rl, st = ctx.inflate_batch(b"".join(streams), so[:-1], np.diff(so).astype(np.uint32), ro)
        for i, s in enumerate(streams):
            want, wst, wn = O.inflate(s, 70000)
            got = out[int(ro[i]):int(ro[i]) + int(rl[i])].tobytes()
            if (int(rl[i]), int(st[i])) != (wn, wst) or got != want[:wn]:
                open(f"/tmp/fuzz_inflate_bad_{seed}_{r}_{i}.bin", "wb").write(s)
                print("MISMATCH round", r, "stream", i, "len", len(s), "ours", int(rl[i]), int(st[i]), "oracle", wn, wst); return False
        total += len(streams)
        print(f"round {r}: {len(streams)} streams ok ({total} so far)", flush=True)
    return True


if __name__ == "__main__":
    ok = run(int(sys.argv[1]) if len(sys.argv) > 1 else 10, int(sys.argv[2]) if len(sys.argv) > 2 else 1)
    sys.exit(0 if ok else 1)
