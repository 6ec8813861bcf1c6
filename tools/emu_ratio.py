"""CPU check of a matcher/parser change BEFORE GPU time is spent on it: compressed size per content class under the SIMT emulator
(the kernel sources compiled for the CPU, tests/simt) against zlib level 6 on the same chunking.  tools/emu_ratio.py [level]"""
import sys, zlib, time
import numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import emu_lib
from tools import corpus

level = int(sys.argv[1]) if len(sys.argv) > 1 else 0
ctx = emu_lib.emu_context()
CH = 65535
classes = {
    "text":   [corpus.gen_text(CH, 596, 200 + i).tobytes() for i in range(6)],
    "struct": [corpus.gen_struct(CH, 596, 210 + i).tobytes() for i in range(4)],
    "bitmap": [corpus.gen_bitmap_like(CH, 596, 220 + i).tobytes() for i in range(4)],
    "c3log":  [corpus.c3_buffer(CH * 6, 596)[i * CH:(i + 1) * CH].tobytes() for i in range(6)],
    "small":  [corpus.gen_text(3000 + 700 * i, 596, 230 + i).tobytes() for i in range(12)] + [corpus.gen_struct(5000 + 900 * i, 596, 250 + i).tobytes() for i in range(8)],
}
_b, _o, _ = corpus.c2_buffer(150, 596)
classes["c2"] = [_b[int(_o[i]):int(_o[i + 1])].tobytes() for i in range(150)]
classes["jpeg"] = [corpus.gen_jpeg_like(3000 + 1777 * i, 596, 300 + i).tobytes() for i in range(12)]
tot_o = tot_z = 0
for name, chunks in classes.items():
    raw = b"".join(chunks)
    lens = np.array([len(c) for c in chunks], dtype=np.uint32)
    off = np.zeros(len(chunks), dtype=np.uint64); off[1:] = np.cumsum(lens)[:-1]
    t = time.time()
    packed, poff, res = ctx.deflate_batch(np.frombuffer(raw, dtype=np.uint8), off, lens, level)
    dt = time.time() - t
    ours = int(poff[-1]); zl = 0
    for i, c in enumerate(chunks):
        assert zlib.decompress(packed[int(poff[i]):int(poff[i + 1])].tobytes()) == c
        zl += len(zlib.compress(c, 6))
    tot_o += ours; tot_z += zl
    print(f"{name:8s} raw {len(raw):8d} ours {ours:8d} zlib6 {zl:8d} ours/zlib6 {ours / zl:.4f} ratio {len(raw) / ours:.3f}  ({dt:.1f} s)")
print(f"total ours/zlib6 {tot_o / tot_z:.4f}")
