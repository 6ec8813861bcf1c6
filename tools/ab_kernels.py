"""A/B of kernel variants on the GPU box: every library given on the command line (builds of csrc/ with different -D switches)
runs deflate + inflate over the same three device-resident corpora; per-kernel times come from the library's own CUDA-event
spans (zwz_profile_*). The inflate output is compared with the input, so a variant that is fast and wrong shows up as WRONG.
    python tools/ab_kernels.py [--mb 256] lib1.so lib2.so ..."""
import sys, os, time, json
import numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import zwz_b200
from tools import corpus

args = sys.argv[1:]
mb = 256
if args and args[0] == "--mb":
    mb = int(args[1]); args = args[2:]
libs = args or [zwz_b200.library_path()]
CH = 65535

def corpora():
    out = {}
    unit = corpus.c3_buffer(8 << 20, 596)
    out["c3"] = (np.tile(unit, mb // 8), None)
    nfiles = int(370000 * mb / 2500)
    buf, offs, _ = corpus.c2_buffer(nfiles, 596)
    out["c2"] = (buf, offs.astype(np.int64))
    buf, offs, _ = corpus.mixed_buffer(min(mb, 256) << 20)
    out["c1"] = (buf, offs)
    return out

t0 = time.time()
data = corpora()
print(f"corpora ready in {time.time() - t0:.1f} s", flush=True)
rows = []
for path in libs:
    lib = zwz_b200.load_library(os.path.abspath(path))
    ctx = zwz_b200.Context(0, library=lib)
    for name, (buf, offs) in data.items():
        U = len(buf)
        foffs = np.array([0, U], dtype=np.int64) if offs is None else offs
        coff, clen, cfile, cseq = zwz_b200.chunk_table(foffs)
        n = len(coff)
        slot = zwz_b200.deflate_bound(clen)
        slot_off = np.zeros(n + 1, dtype=np.uint64); np.cumsum(slot, out=slot_off[1:])
        d_raw = ctx.malloc_device(U + 64); d_slots = ctx.malloc_device(int(slot_off[-1]) + 64)
        d_packed = ctx.malloc_device(int(slot_off[-1]) + 64); d_back = ctx.malloc_device(U + 64)
        ctx.h2d(d_raw, buf)
        raw_off = np.concatenate([coff, [np.uint64(U)]]).astype(np.uint64)
        reps = 3
        for rep in range(reps + 1):
            if rep == 1:
                ctx.profile_enable(True); ctx.profile_read(True)
            res = ctx.deflate_batch_device(d_raw, coff, clen, d_slots, slot_off[:-1])
            poff = ctx.pack_streams_device(d_slots, slot_off[:-1], res, d_packed)
            split = res["len1"] > 0
            if split.any():
                k = np.nonzero(split)[0]
                r_off = np.insert(poff[:-1], k + 1, poff[:-1][k] + res["len0"][k].astype(np.uint64))
                r_len = np.insert(res["len0"], k + 1, res["len1"][k])
                r_raw = np.concatenate([np.insert(raw_off[:-1], k + 1, raw_off[:-1][k] + res["raw0"][k].astype(np.uint64)), raw_off[-1:]])
            else:
                r_off, r_len, r_raw = poff[:-1], res["len0"], raw_off
            rl, st = ctx.inflate_batch_device(d_packed, r_off, r_len, d_back, r_raw)
        prof = ctx.profile_read(True); ctx.profile_enable(False)
        back = np.empty(U, dtype=np.uint8); ctx.d2h(back, d_back)
        ok = bool((st == 0).all()) and np.array_equal(back, buf)
        C = int(poff[-1])
        k = {kk: v[0] / reps for kk, v in prof.items() if v[1]}
        row = {"lib": os.path.basename(path), "corpus": name, "MB": U / 1e6, "ok": ok, "ratio": U / C, **{kk: round(v, 2) for kk, v in k.items()}}
        rows.append(row)
        print(json.dumps(row), flush=True)
        for d in (d_raw, d_slots, d_packed, d_back):
            ctx.free_device(d)
    ctx.close()
if not all(r["ok"] for r in rows):
    print("WRONG RESULTS in", [(r["lib"], r["corpus"]) for r in rows if not r["ok"]])
    sys.exit(1)
