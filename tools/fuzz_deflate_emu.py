"""CPU fuzz of the deflate path under the SIMT emulator: chunks assembled from segments of different kinds (noise, text, repeats
of earlier bytes at any distance, periodic tables, runs, ramps) with boundaries anywhere relative to the 1 024-byte stretches
and 32-position tiles. Every stream must inflate (zlib) to its input; sizes are compared with zlib level 6.
    python tools/fuzz_deflate_emu.py [rounds] [seed]"""
import os, sys, zlib
import numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
os.environ.setdefault("ZWZ_EMU_POISON", "9")
import emu_lib
from tools import corpus

def run(rounds, seed, chunks_per_round=None):
    """returns True when every stream checked out"""
    rng = np.random.default_rng(seed)
    ctx = emu_lib.emu_context()
    text = corpus.gen_text(200000, 596, 7).tobytes()
    struct = corpus.gen_struct(100000, 596, 8).tobytes()

    def segment(prev: bytes) -> bytes:
        kind = rng.integers(0, 9)
        n = int(rng.choice([rng.integers(1, 40), rng.integers(40, 600), rng.integers(600, 3000), rng.integers(3000, 20000)]))
        if kind == 0:
            return rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        if kind == 1:
            o = int(rng.integers(0, len(text) - n)); return text[o:o + n]
        if kind == 2 and len(prev) > 8:   # repeat of earlier bytes (any distance, maybe overlapping)
            o = int(rng.integers(0, len(prev))); d = prev[o:o + n]
            return (d * (n // max(len(d), 1) + 1))[:n]
        if kind == 3:
            p = rng.permutation(256).astype(np.uint8).tobytes(); return (p * (n // 256 + 1))[:n]
        if kind == 4:
            return bytes([int(rng.integers(0, 256))]) * n
        if kind == 5:
            return bytes((np.arange(n) * int(rng.integers(1, 7)) & 255).astype(np.uint8))
        if kind == 6:
            o = int(rng.integers(0, len(struct) - n)); return struct[o:o + n]
        if kind == 7:   # low-entropy noise (4 bits per byte)
            return (rng.integers(0, 16, n, dtype=np.uint8) * 17).tobytes()
        return rng.integers(0, 256, n, dtype=np.uint8).tobytes()

    tot_o = tot_z = 0
    worst = (0.0, None)
    for r in range(rounds):
        chunks = []
        for _ in range(48):
            c = b""
            target = int(rng.choice([rng.integers(1, 300), rng.integers(300, 5000), rng.integers(5000, 30000), 65535]))
            while len(c) < target:
                c += segment(c)
            chunks.append(c[:min(target, 65535)])
        raw = b"".join(chunks)
        lens = np.array([len(c) for c in chunks], dtype=np.uint32)
        off = np.zeros(len(chunks), dtype=np.uint64); off[1:] = np.cumsum(lens)[:-1]
        packed, poff, res = ctx.deflate_batch(np.frombuffer(raw, dtype=np.uint8), off, lens)
        for i, c in enumerate(chunks):
            p = packed[int(poff[i]):int(poff[i + 1])].tobytes()
            l0 = int(res["len0"][i])
            back = zlib.decompress(p[:l0]) + (zlib.decompress(p[l0:]) if res["len1"][i] else b"")
            if back != c:
                np.save(f"/tmp/fuzz_bad_{seed}_{r}_{i}.npy", np.frombuffer(c, dtype=np.uint8))
                print("MISMATCH round", r, "chunk", i, "len", len(c)); return False
            z = len(zlib.compress(c, 6))
            tot_o += len(p); tot_z += z
            if len(c) > 2000 and len(p) / z > worst[0]:
                worst = (len(p) / z, (r, i, len(c), len(p), z))
                np.save(f"/tmp/fuzz_worst_{seed}.npy", np.frombuffer(c, dtype=np.uint8))
        print(f"round {r}: ok, running ours/zlib6 = {tot_o / tot_z:.4f}, worst chunk so far {worst}", flush=True)
    return True


if __name__ == "__main__":
    ok = run(int(sys.argv[1]) if len(sys.argv) > 1 else 10, int(sys.argv[2]) if len(sys.argv) > 2 else 1)
    sys.exit(0 if ok else 1)
