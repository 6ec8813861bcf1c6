"""End-to-end leg of bench.py only (host buffers -> zwz_compress_files -> zwz_decompress_records -> host buffers), for several
(workers x parts) settings on ONE build of the C2 shard: tools/e2e_sweep.py [--files 370000] 6x32 6x16 4x16 ...
Same part cut, same sizing pass and same per-part work as bench.py; prints one JSON line per setting."""
import json, os, sys, threading, time
from concurrent.futures import ThreadPoolExecutor
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import torch  # noqa: E402
import zwz_b200  # noqa: E402

args = sys.argv[1:]
files = 370000
if args and args[0] == "--files":
    files = int(args[1]); args = args[2:]
EMU = os.environ.get("ZWZ_SWEEP_EMU") == "1"   # dry run of this script on the CPU emulator build (tests/simt), no timing value
PIN = not EMU
LIB = zwz_b200.load_library(os.path.join(ROOT, "tests", "simt", "libzwz_emu.so")) if EMU else None
settings = [tuple(int(x) for x in a.split("x")) for a in args] or [(6, 32)]
sh = bench.build_shard("c2", files, 0, 1)
U, foffs = sh.U, sh.foffs
nf = len(foffs) - 1
coff, clen, cfile, cseq = zwz_b200.chunk_table(foffs)
n = len(coff)
slot = zwz_b200.deflate_bound(clen)
slot_off = np.zeros(n + 1, dtype=np.uint64); np.cumsum(slot, out=slot_off[1:])
raw_off = np.concatenate([coff, [np.uint64(U)]]).astype(np.uint64)
h_raw = torch.empty(U, dtype=torch.uint8, pin_memory=PIN); hr = h_raw.numpy(); hr[:] = sh.unit
h_back = torch.empty(U, dtype=torch.uint8, pin_memory=PIN)
h_comp = torch.empty(int(slot_off[-1]) + 64, dtype=torch.uint8, pin_memory=PIN)
hc_ptr, hr_ptr, hb_ptr = h_comp.data_ptr(), h_raw.data_ptr(), h_back.data_ptr()
first_chunk_of_file = np.concatenate([[0], np.cumsum(np.bincount(cfile, minlength=nf))]).astype(np.int64)
for W, P in settings:
    part_file = np.unique(np.searchsorted(foffs, np.linspace(0, U, P + 1))); part_file[0], part_file[-1] = 0, nf
    part_file = np.unique(part_file)
    part_chunk = first_chunk_of_file[part_file]
    nparts = len(part_chunk) - 1
    desc = bench.part_descriptors(sh, coff, clen, cfile, raw_off, slot_off, part_chunk, part_file, None)
    workers = [zwz_b200.Context(0, library=LIB) if EMU else zwz_b200.Context(0) for _ in range(W)]
    for w in workers:
        w.tune(w.TUNE_DEFLATE_SUBBATCH_BYTES, int(os.environ.get("ZWZ_SWEEP_SUBBATCH_MB", "64")) << 20)
    lock = threading.Lock(); free = list(range(W))

    def run_part(i, wi):
        d = desc[i]
        cap = int(slot_off[d["c1"]] - slot_off[d["c0"]]) + 64
        return bench.host_roundtrip(workers[wi], d, clen, hr_ptr + d["y0"], hc_ptr + d["hc"], cap, hb_ptr + d["y0"], 0, True)

    def part(i):
        with lock:
            wi = free.pop()
        try:
            return run_part(i, wi)
        finally:
            with lock:
                free.append(wi)

    sizing = sorted({max(range(nparts), key=lambda i: part_chunk[i + 1] - part_chunk[i]), max(range(nparts), key=lambda i: int(raw_off[part_chunk[i + 1]] - raw_off[part_chunk[i]])), nparts - 1})
    with ThreadPoolExecutor(W) as sp:
        list(sp.map(lambda wi: [run_part(i, wi) for i in sizing], range(W)))
    pool = ThreadPoolExecutor(W)
    for _ in range(2):
        list(pool.map(part, range(nparts)))
    if not EMU:
        torch.cuda.synchronize()
    times = []
    for _ in range(4):
        t0 = time.perf_counter(); list(pool.map(part, range(nparts)))
        if not EMU:
            torch.cuda.synchronize()
        times.append(time.perf_counter() - t0)
    assert np.array_equal(h_back.numpy(), hr)
    pool.shutdown()
    for w in workers:
        w.close()
    print(json.dumps({"workers": W, "parts": nparts, "e2e_gbs_best": U / min(times) / 1e9, "e2e_gbs_mean": U * len(times) / sum(times) / 1e9, "ms": [round(1e3 * t, 1) for t in times]}), flush=True)
