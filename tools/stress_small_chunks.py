"""GPU stress of the shape that exposed the unset stored-region marks: 20 000 C2 files of ~4 KiB, deflated 8 times (which warp takes
which chunk differs from run to run), every stream checked with zlib."""
import sys, zlib, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import zwz_b200
from tools import corpus
nfiles = 20000
sizes = corpus.c2_sizes(50000, 596)
want = sizes[np.argsort(-sizes, kind="stable")][30000:30000 + nfiles]
buf, _, _ = corpus.c2_buffer(nfiles + 200, 596)
foffs = np.zeros(nfiles + 1, dtype=np.int64)
np.cumsum(want, out=foffs[1:])
buf = buf[:int(foffs[-1])]
off = foffs[:-1].astype(np.uint64)
ctx = zwz_b200.Context(0)
total_bad = 0
for rep in range(8):
    packed, poff, res = ctx.deflate_batch(buf, off, want.astype(np.uint32))
    bad = []
    for i in range(nfiles):
        try:
            ok = zlib.decompress(packed[int(poff[i]):int(poff[i + 1])].tobytes()) == buf[int(foffs[i]):int(foffs[i + 1])].tobytes()
        except Exception:
            ok = False
        if not ok:
            bad.append(i)
    total_bad += len(bad)
    print("rep", rep, "bad", len(bad), bad[:12], [int(want[i]) for i in bad[:12]], flush=True)
ctx.close()
print("TOTAL BAD", total_bad)
sys.exit(1 if total_bad else 0)
