#!/usr/bin/env python
"""bench.py — the hot path of BASELINE.json on synthetic corpora of the named shapes.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c3|c1] [--files F]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
  python bench.py --impl reference ...        # the reference's CPU path (zlib L6 / zlib inflate / OpenSSL MD5) on host cores

A "step" is one pass of the hot path over the rank's shard of the workload:
  compress side   deflate of every 65 535-byte chunk (compression.cpp:119-134) + MD5 of every source file (:95-103)
  decompress side inflate of every record (decompression.cpp:11-37) + MD5 of every output file (:136)
`value` is uncompressed bytes of the whole job / device time of that step with the inputs already resident in HBM
(CUDA events on the launching stream, max over ranks); `e2e` is the same step from PINNED HOST buffers through the same
C ABI with the H2D/D2H copies inside the timed region. Default workload = BASELINE.json configs[1] (C2: 370 000 image-like
files, ~2.5 GB, every file one sub-65 535-byte chunk). Multi-GPU: the reference's policy — sort by size descending, rank r
takes sorted files r, r+N, ... (file_sort.cpp:30-31, compression.cpp:31-41); no data-path collective, one NCCL
all-gather of per-rank counters; per-rank work is fixed as N grows => "scaling": "weak".
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from tools import corpus  # noqa: E402

CHUNK = 65535
METRIC = "deflate/inflate GB/s (uncompressed)"
UNIT = "GB/s"


# ---------------------------------------------------------------------------------------------------------------------
# workload: every rank builds ITS shard only. File sizes come from the global list (N x the single-GPU corpus) sorted by
# size descending and dealt round-robin, exactly the reference's distribution rule; contents are generated per rank.
# ---------------------------------------------------------------------------------------------------------------------
def build_shard(workload: str, files: int, rank: int, world: int):
    seed = corpus.BASE_SEED
    if workload == "c2":
        # content keyed by (seed, rank, directory); sizes from the global size-sorted deal
        all_sizes = corpus.c2_sizes(files * world, seed)
        order = np.argsort(-all_sizes, kind="stable")
        mine = order[rank::world]
        buf, offs, sizes = corpus.c2_buffer(len(mine), seed + 7919 * rank)
        # c2_buffer draws its own sizes; re-cut the buffer to the dealt sizes (same generator, same class mix): regenerate
        # with the dealt sizes so the shard is size-descending like the reference's record file
        want = all_sizes[mine]
        tot = int(want.sum())
        if tot > len(buf):
            extra, _, _ = corpus.c2_buffer(int((tot - len(buf)) // 6000 + 1000), seed + 7919 * rank + 1)
            buf = np.concatenate([buf, extra])
        buf = buf[:tot]
        foffs = np.zeros(len(want) + 1, dtype=np.int64)
        np.cumsum(want, out=foffs[1:])
        desc = f"C2: {files} image-like files/GPU (70% JPEG-like, 30% bitmap-like), one chunk each"
        return buf, foffs, desc
    if workload == "c3":
        size = files  # bytes
        unit = corpus.gen_text(min(size, 32 << 20), seed, 0x100000 + rank)
        reps = (size + len(unit) - 1) // len(unit)
        buf = np.tile(unit, reps)[:size]
        return buf, np.array([0, size], dtype=np.int64), f"C3: one {size}-byte log/text file per GPU, 65 535-byte chunks (32 MiB period)"
    if workload == "c1":
        buf, offs, specs = corpus.mixed_buffer(files, seed + 1 + rank, 4096, 16 << 20)
        sizes = np.diff(offs)
        order = np.argsort(-sizes, kind="stable")
        parts = [buf[offs[i]:offs[i + 1]] for i in order]
        foffs = np.zeros(len(order) + 1, dtype=np.int64)
        np.cumsum(sizes[order], out=foffs[1:])
        return np.concatenate(parts), foffs, f"C1-shaped: {files} bytes of mixed T/S/I/R files per GPU, size-descending"
    raise SystemExit(f"unknown workload {workload}")


# ---------------------------------------------------------------------------------------------------------------------
# clocks (nvidia-smi sampled DURING the timed region)
# ---------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index: int):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for k, nme in enumerate(names):
                if f[3 + k].lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(max(mx)) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------------------
# CPU baseline: the reference's call sequences (oracle_ref_* in oracle/zwz_oracle.c: system zlib L6 deflate, zlib inflate,
# OpenSSL MD5 in 1024-byte updates), fanned out over host threads (ctypes releases the GIL).
# ---------------------------------------------------------------------------------------------------------------------
def cpu_step(buf, foffs, coff, clen, threads, do_md5=True):
    """One full step (compress side + decompress side) on the CPU. Returns (seconds, compressed_bytes)."""
    import ctypes as C
    from concurrent.futures import ThreadPoolExecutor

    import oracle_lib as O
    L = O.lib()
    n = len(coff)
    nf = len(foffs) - 1
    out = np.empty((n, CHUNK), dtype=np.uint8)
    out_len = np.zeros(n, dtype=np.uint32)
    back = np.empty(int(foffs[-1]) + 1, dtype=np.uint8)
    raw_off = np.concatenate([coff, [np.uint64(foffs[-1])]]).astype(np.uint64)
    raw_len = np.zeros(n, dtype=np.uint32)
    hex1 = np.zeros(nf * 32, dtype=np.uint8)
    hex2 = np.zeros(nf * 32, dtype=np.uint8)
    fo = foffs[:-1].astype(np.uint64)
    fl = np.diff(foffs).astype(np.uint64)
    cs = np.linspace(0, n, threads + 1).astype(np.int64)
    fs = np.linspace(0, nf, threads + 1).astype(np.int64)

    def run(fn):
        with ThreadPoolExecutor(threads) as ex:
            list(ex.map(fn, range(threads)))

    def deflate(t):
        a, b = int(cs[t]), int(cs[t + 1])
        if b > a:
            L.oracle_ref_deflate_batch(buf.ctypes.data, coff[a:].ctypes.data, clen[a:].ctypes.data, b - a, out[a:].ctypes.data, out_len[a:].ctypes.data)

    def inflate(t):
        a, b = int(cs[t]), int(cs[t + 1])
        if b > a:
            L.oracle_ref_inflate_batch(out[a:].ctypes.data, out_len[a:].ctypes.data, b - a, back.ctypes.data, raw_off[a:].ctypes.data, raw_len[a:].ctypes.data)

    def md5(src, dst):
        def f(t):
            a, b = int(fs[t]), int(fs[t + 1])
            if b > a:
                L.oracle_ref_md5_batch(src.ctypes.data, fo[a:].ctypes.data, fl[a:].ctypes.data, b - a, dst[32 * a:].ctypes.data)
        return f

    t0 = time.perf_counter()
    run(deflate)
    if do_md5:
        run(md5(buf, hex1))
    run(inflate)
    if do_md5:
        run(md5(back, hex2))
    dt = time.perf_counter() - t0
    ok = bool((raw_len == clen).all()) and np.array_equal(hex1, hex2)
    return dt, int(out_len.sum()), ok


def reference_binary_step(main_ref, sb, so, cores, workload):
    """Materialise the sample as a directory tree in RAM and return (step(), P): step() runs `main_ref compress` with P
    concurrently running emulated ranks (the MPI stub reads ZWZ_STUB_RANK/SIZE) and then `main_ref decompress`, and returns
    (seconds, archive bytes). The reference's per-file stdout goes to /dev/null."""
    import shutil
    import tempfile
    root = tempfile.mkdtemp(prefix="zwz_ref_", dir="/dev/shm")
    import atexit
    atexit.register(lambda: shutil.rmtree(root, ignore_errors=True))
    src = os.path.join(root, "w", "src")
    nf = len(so) - 1
    for d in range((nf + 999) // 1000):
        os.makedirs(os.path.join(src, f"dir{d:03d}"), exist_ok=True)
    for i in range(nf):
        sb[so[i]:so[i + 1]].tofile(os.path.join(src, f"dir{i // 1000:03d}", f"f{i % 1000:04d}.dat"))
    P = max(1, min(cores, nf))
    state = {"k": 0}

    def step():
        state["k"] += 1
        arch = os.path.join(root, f"arch{state['k']}")
        out = os.path.join(root, f"out{state['k']}")
        bc = os.path.join(root, f"bc{state['k']}")
        os.makedirs(bc)
        dn = open(os.devnull, "w")
        t0 = time.perf_counter()
        env0 = dict(os.environ, ZWZ_STUB_SIZE=str(P), ZWZ_STUB_RANK="0", ZWZ_STUB_DIR=bc, OMP_NUM_THREADS="2")
        p0 = subprocess.Popen([main_ref, "compress", src, arch], env=env0, stdout=dn, stderr=dn)
        while not os.path.exists(os.path.join(bc, "bcast_1")) and p0.poll() is None:
            time.sleep(0.001)  # rank 0 publishes the record file path; then the others may start
        procs = [p0] + [subprocess.Popen([main_ref, "compress", src, arch],
                                         env=dict(os.environ, ZWZ_STUB_SIZE=str(P), ZWZ_STUB_RANK=str(r), ZWZ_STUB_DIR=bc), stdout=dn, stderr=dn)
                        for r in range(1, P)]
        for p in procs:
            p.wait()
        subprocess.run([main_ref, "decompress", arch, out], stdout=dn, stderr=dn, env=dict(os.environ, ZWZ_STUB_SIZE="1"))
        dt = time.perf_counter() - t0
        comp = sum(os.path.getsize(os.path.join(arch, f)) for f in os.listdir(arch))
        shutil.rmtree(arch, ignore_errors=True)
        shutil.rmtree(out, ignore_errors=True)
        shutil.rmtree(bc, ignore_errors=True)
        return dt, comp

    return step, P


def sample_of(buf, foffs, target_bytes):
    """Bounded sample of the same workload: whole files taken evenly across the size-sorted shard."""
    nf = len(foffs) - 1
    tot = int(foffs[-1])
    if tot <= target_bytes:
        return buf, foffs, "whole shard"
    if nf == 1:
        n = (target_bytes // CHUNK) * CHUNK
        return buf[:n], np.array([0, n], dtype=np.int64), f"first {n} bytes of the file"
    stride = max(1, int(tot // target_bytes))
    idx = np.arange(0, nf, stride)
    sizes = np.diff(foffs)[idx]
    so = np.zeros(len(idx) + 1, dtype=np.int64)
    np.cumsum(sizes, out=so[1:])
    sb = np.empty(int(so[-1]), dtype=np.uint8)
    for k, i in enumerate(idx):
        sb[so[k]:so[k + 1]] = buf[foffs[i]:foffs[i + 1]]
    return sb, so, f"every {stride}-th file of the size-sorted shard ({len(idx)} files, {int(so[-1])} bytes)"


# ---------------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=["c2", "c3", "c1"])
    ap.add_argument("--files", type=int, default=0, help="c2: files per GPU (default 370000); c3/c1: bytes per GPU")
    ap.add_argument("--level", type=int, default=0)
    ap.add_argument("--cpu-sample-mb", type=float, default=0.0, help="CPU baseline sample size (default: ~15 s of work)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-workers", type=int, default=0, help="end-to-end leg: worker contexts (stream + buffers each) per rank; 0 = min(6, host cores / ranks), at least 2")
    ap.add_argument("--e2e-parts", type=int, default=32, help="end-to-end leg: parts the shard is cut into")
    ap.add_argument("--md5", default="auto", choices=["auto", "on", "off"],
                    help="auto: on for c2/c1 (compress+decompress+MD5 verify), off for c3 (BASELINE.json config 3 is deflate+inflate only: "
                         "the MD5 of ONE file is a single serial chain, one lane, ~0.1 GB/s)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3  # timing rule: W >= 3
    if args.files == 0:
        args.files = {"c2": 370_000, "c3": 2 << 30, "c1": 2 << 30}[args.workload]

    do_md5 = args.md5 == "on" or (args.md5 == "auto" and args.workload != "c3")
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)

    # ------------------------------------------------------------------------------------------------ reference arm
    if args.impl == "reference":
        if rank != 0:
            return
        buf, foffs, desc = build_shard(args.workload, args.files, 0, 1)
        target = int(args.cpu_sample_mb * 1e6) if args.cpu_sample_mb else int(min(8e6 * cores, 400e6))
        sb, so, what = sample_of(buf, foffs, target)
        main_ref = os.path.join(ROOT, "oracle", "_ref", "main_ref")
        if os.path.exists(main_ref) and os.path.isdir("/dev/shm"):
            kind = "reference"
            step = "UNMODIFIED reference binary (oracle/_ref/main_ref): compress with P emulated MPI ranks + decompress (1 process, OpenMP over archives) on a RAM-backed tree"
            run_step, P = reference_binary_step(main_ref, sb, so, cores, args.workload)
            what += f"; written as files under /dev/shm; compress ranks = {P}"
        else:
            kind = "port"
            step = "deflate(zlib L6)+MD5(src)+inflate+MD5(out) per chunk/file on host cores (oracle_ref_* call sequences)"
            coff, clen, _, _ = corpus.chunk_table(so)
            run_step = lambda: cpu_step(sb, so, coff, clen, cores, do_md5)[:2]
        times = []
        comp = 0
        for i in range(args.warmup + args.steps):
            dt, comp = run_step()
            if i >= args.warmup:
                times.append(dt)
        t = sum(times)
        val = len(sb) * len(times) / t / 1e9
        line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": 1e3 * t / len(times), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
                "data": "synthetic", "config": {"workload": desc, "step": step},
                "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": kind, "sample": what,
                                 "ratio": len(sb) / max(comp, 1)},
                "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    # ------------------------------------------------------------------------------------------------ our arm
    import torch
    import torch.distributed as dist

    import zwz_b200

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = zwz_b200.Context(local_rank)  # raises when libzwz_cuda.so or the GPU is missing: no fallback exists
    main_stream = torch.cuda.Stream()
    torch.cuda.set_stream(main_stream)       # everything below (events, copies, our kernels) is ordered on this stream
    stream = main_stream.cuda_stream

    buf, foffs, desc = build_shard(args.workload, args.files, rank, world)
    U = int(foffs[-1])
    nf = len(foffs) - 1
    coff, clen, cfile, cseq = zwz_b200.chunk_table(foffs)
    n = len(coff)
    slot = zwz_b200.deflate_bound(clen)
    slot_off = np.zeros(n + 1, dtype=np.uint64)
    np.cumsum(slot, out=slot_off[1:])
    raw_off = np.concatenate([coff, [np.uint64(U)]]).astype(np.uint64)
    f_off = foffs[:-1].astype(np.uint64)
    f_len = np.diff(foffs).astype(np.uint64)

    h_raw = torch.empty(U, dtype=torch.uint8, pin_memory=True)
    h_raw.numpy()[:] = buf
    h_back = torch.empty(U, dtype=torch.uint8, pin_memory=True)
    h_comp = torch.empty(int(slot_off[-1]), dtype=torch.uint8, pin_memory=True)
    d_raw = torch.empty(U + 64, dtype=torch.uint8, device="cuda")
    d_slots = torch.empty(int(slot_off[-1]) + 64, dtype=torch.uint8, device="cuda")
    d_packed = torch.empty(int(slot_off[-1]) + 64, dtype=torch.uint8, device="cuda")
    d_back = torch.empty(U + 64, dtype=torch.uint8, device="cuda")
    d_raw[:U].copy_(h_raw, non_blocking=True)
    torch.cuda.synchronize()

    state = {}

    def device_step():
        """inputs resident in HBM"""
        res = ctx.deflate_batch_device(d_raw.data_ptr(), coff, clen, d_slots.data_ptr(), slot_off[:-1], args.level, stream)
        dg1 = ctx.md5_batch_device(d_raw.data_ptr(), f_off, f_len, stream) if do_md5 else None
        poff = ctx.pack_streams_device(d_slots.data_ptr(), slot_off[:-1], res, d_packed.data_ptr(), stream)
        # records: one per stream (split chunks give two)
        r_off, r_len, r_raw_off = records_of(res, poff, raw_off)
        rl, st = ctx.inflate_batch_device(d_packed.data_ptr(), r_off, r_len, d_back.data_ptr(), r_raw_off, 0, stream)
        dg2 = ctx.md5_batch_device(d_back.data_ptr(), f_off, f_len, stream) if do_md5 else None
        state.update(res=res, dg1=dg1, dg2=dg2, rl=rl, st=st, poff=poff, r_raw_off=r_raw_off)

    def records_of(res, poff, raw_off):
        split = res["len1"] > 0
        if not split.any():
            return poff[:-1], res["len0"], raw_off
        k = np.nonzero(split)[0]
        r_off = np.insert(poff[:-1], k + 1, poff[:-1][k] + res["len0"][k].astype(np.uint64))
        r_len = np.insert(res["len0"], k + 1, res["len1"][k])
        r_raw = np.insert(raw_off[:-1], k + 1, raw_off[:-1][k] + res["raw0"][k].astype(np.uint64))
        return r_off, r_len, np.concatenate([r_raw, raw_off[-1:]])

    # ---- end-to-end: W workers (own zwz ctx + CUDA stream + device buffers each) take parts of the shard in turn, so the
    # H2D of one part, the kernels of another and the D2H of a third overlap (PCIe is full duplex). Every byte still starts in
    # pinned host memory and ends in pinned host memory inside the timed region.
    import threading
    from concurrent.futures import ThreadPoolExecutor
    W = args.e2e_workers if args.e2e_workers > 0 else max(2, min(6, (os.cpu_count() or 16) // max(world, 1)))
    first_chunk_of_file = np.concatenate([[0], np.cumsum(np.bincount(cfile, minlength=nf))]).astype(np.int64)
    if nf > 1:   # cut on file boundaries (MD5 needs whole files), parts of about equal bytes
        want = min(args.e2e_parts, max(1, nf // 2000))
        part_file = np.unique(np.searchsorted(foffs, np.linspace(0, U, want + 1)))
        part_file[0], part_file[-1] = 0, nf
        part_file = np.unique(part_file)
        part_chunk = first_chunk_of_file[part_file]
    elif do_md5:  # one file and its MD5 wanted: a single part (the digest is one serial chain anyway)
        part_file = np.array([0, 1])
        part_chunk = np.array([0, n], dtype=np.int64)
    else:
        part_file = None
        part_chunk = np.unique(np.linspace(0, n, min(args.e2e_parts, max(1, n // 512)) + 1).astype(np.int64))
    nparts = len(part_chunk) - 1
    max_raw = max(int(raw_off[part_chunk[i + 1]] - raw_off[part_chunk[i]]) for i in range(nparts))
    max_slot = max(int(slot_off[part_chunk[i + 1]] - slot_off[part_chunk[i]]) for i in range(nparts))
    workers = []
    for wi in range(W):
        wctx = zwz_b200.Context(local_rank)
        workers.append(dict(ctx=wctx, s=torch.cuda.Stream(),
                            raw=torch.empty(max_raw + 64, dtype=torch.uint8, device="cuda"),
                            slots=torch.empty(max_slot + 64, dtype=torch.uint8, device="cuda"),
                            packed=torch.empty(max_slot + 64, dtype=torch.uint8, device="cuda"),
                            back=torch.empty(max_raw + 64, dtype=torch.uint8, device="cuda")))
    wlocal = threading.local()
    wlock = threading.Lock()
    wfree = list(range(W))
    comp_bytes = [0] * nparts

    def e2e_part(i):
        with wlock:
            wi = wfree.pop()
        try:
            w = workers[wi]
            c0, c1 = int(part_chunk[i]), int(part_chunk[i + 1])
            b0, b1 = int(raw_off[c0]), int(raw_off[c1])
            so0 = slot_off[c0]
            with torch.cuda.stream(w["s"]):
                sp = w["s"].cuda_stream
                w["raw"][:b1 - b0].copy_(h_raw[b0:b1], non_blocking=True)
                pc = coff[c0:c1] - np.uint64(b0)
                res = w["ctx"].deflate_batch_device(w["raw"].data_ptr(), pc, clen[c0:c1], w["slots"].data_ptr(), slot_off[c0:c1] - so0, args.level, sp)
                if do_md5:
                    f0, f1 = int(part_file[i]), int(part_file[i + 1])
                    dg1 = w["ctx"].md5_batch_device(w["raw"].data_ptr(), f_off[f0:f1] - np.uint64(b0), f_len[f0:f1], sp)
                poff = w["ctx"].pack_streams_device(w["slots"].data_ptr(), slot_off[c0:c1] - so0, res, w["packed"].data_ptr(), sp)
                C = int(poff[-1])
                hc0 = int(so0)
                h_comp[hc0:hc0 + C].copy_(w["packed"][:C], non_blocking=True)   # payloads -> host (what a .zwz holds)
                w["s"].synchronize()
                w["packed"][:C].copy_(h_comp[hc0:hc0 + C], non_blocking=True)   # decompress side starts from host bytes
                pr = np.concatenate([pc, [np.uint64(b1 - b0)]]).astype(np.uint64)
                r_off, r_len, r_raw_off = records_of(res, poff, pr)
                rl, st = w["ctx"].inflate_batch_device(w["packed"].data_ptr(), r_off, r_len, w["back"].data_ptr(), r_raw_off, 0, sp)
                if do_md5:
                    dg2 = w["ctx"].md5_batch_device(w["back"].data_ptr(), f_off[f0:f1] - np.uint64(b0), f_len[f0:f1], sp)
                    assert np.array_equal(dg1, dg2)
                h_back[b0:b1].copy_(w["back"][:b1 - b0], non_blocking=True)
                w["s"].synchronize()
            assert (st == 0).all()
            comp_bytes[i] = C
        finally:
            with wlock:
                wfree.append(wi)

    pool = ThreadPoolExecutor(W)

    def e2e_step():
        list(pool.map(e2e_part, range(nparts)))
        state.update(C=sum(comp_bytes))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # warm-up (also sizes every arena)
    for _ in range(args.warmup):
        device_step()
    barrier()
    # correctness of what we are about to time
    res = state["res"]
    Cbytes = int(res["len0"].sum() + res["len1"].sum())
    assert (state["st"] == 0).all(), "inflate status"
    assert (not do_md5) or np.array_equal(state["dg1"], state["dg2"]), "MD5 verify failed"
    back = d_back[:U].cpu().numpy()
    assert np.array_equal(back, buf), "round trip mismatch"

    sampler = ClockSampler(local_rank)
    sampler.start()
    ctx.profile_enable(True)
    ctx.profile_read(True)
    launches0 = ctx.launches
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        device_step()
    e1.record()
    barrier()
    dev_ms = e0.elapsed_time(e1)
    launches = ctx.launches - launches0
    prof = ctx.profile_read(True)
    ctx.profile_enable(False)

    # e2e
    for _ in range(2):
        e2e_step()
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    for _ in range(args.steps):
        e2e_step()
    e3.record()
    barrier()
    e2e_ms = e2.elapsed_time(e3)
    clocks = sampler.stop()
    assert np.array_equal(h_back.numpy(), buf), "e2e round trip mismatch"

    # max over ranks, totals over ranks (the ONLY collective: a few counters per GPU)
    t = torch.tensor([dev_ms, e2e_ms], dtype=torch.float64, device="cuda")
    cnt = torch.tensor([U, Cbytes, n, nf, launches], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        allc = [torch.zeros_like(cnt) for _ in range(world)]
        dist.all_gather(allc, cnt)
        cnt = torch.stack(allc).sum(0)
    dev_ms, e2e_ms = float(t[0]), float(t[1])
    U_all, C_all, n_all, nf_all, launches_all = [float(x) for x in cnt]
    # ---- the chunk-offset exchange: when ONE file is cut into contiguous chunk ranges over the
    # GPUs, all its records must land in ONE archive (decompression.cpp:52-55), so every rank needs the byte offset at which
    # its records start: record bytes per rank -> all-gather -> exclusive scan. 13 + path_len header bytes per record, +32 for
    # the MD5 behind the file's last record. Sequence ids are offset the same way (records, not chunks: split chunks count twice).
    exchange = None
    if args.workload == "c3":
        path_len = len("big/huge.log")
        n_records = int(n + int((res["len1"] > 0).sum()))
        rec_bytes = Cbytes + n_records * (13 + path_len) + (32 if rank == world - 1 else 0)
        mine = torch.tensor([float(rec_bytes), float(n_records)], dtype=torch.float64, device="cuda")
        allv = [torch.zeros_like(mine) for _ in range(world)]
        if world > 1:
            dist.all_gather(allv, mine)
        else:
            allv = [mine]
        sizes_ = [int(v[0]) for v in allv]
        recs_ = [int(v[1]) for v in allv]
        exchange = {"rank_record_bytes": sizes_, "rank_write_offset": [int(sum(sizes_[:i])) for i in range(world)],
                    "rank_first_sequence_id": [int(sum(recs_[:i])) for i in range(world)], "archive_bytes": int(sum(sizes_))}

    if rank == 0:
        K = args.steps
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
        ms = {k: v[0] for k, v in prof.items()}
        nl = {k: v[1] for k, v in prof.items()}
        # dominant kernel by device time on rank 0
        dom = max(("lz_match", "deflate_encode", "inflate") + (("md5",) if do_md5 else ()), key=lambda k: ms[k])
        alg_bytes = {"lz_match": U, "deflate_encode": U + Cbytes, "inflate": U + Cbytes, "md5": 2 * U}[dom] * K  # per K steps on rank 0
        achieved = alg_bytes / (ms[dom] * 1e-3) / 1e9 if ms[dom] > 0 else 0.0
        traffic = None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))   # measured on one workload only
            if tj.get("workload") == args.workload and (args.files in (0, tj.get("files"))):
                traffic = tj["kernels"].get(dom) * K / max(nl[dom], 1)   # per launch, like algorithmic_bytes_per_launch
        except Exception:
            pass
        line = {
            "metric": METRIC, "value": U_all * K / (dev_ms * 1e-3) / 1e9, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": args.warmup,
            "ms_per_step": dev_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": desc, "e2e_pipeline": f"{W} workers x {nparts} parts, H2D/kernels/D2H overlapped", "step": ("deflate+MD5(src)+pack+inflate+MD5(out)" if do_md5 else "deflate+pack+inflate (no MD5: one file = one serial chain)") + ", all through the C ABI", "chunks": int(n_all),
                       "files": int(nf_all), "uncompressed_bytes": int(U_all), "l2": "inputs (>= 2 GB/GPU) exceed the 126 MB L2",
                       "level": args.level, "parallelism": f"files dealt size-descending round-robin over {world} GPU(s)"},
            "ratio": U_all / C_all,
            "deflate_gbs": U * K / ((ms["lz_match"] + ms["deflate_encode"]) * 1e-3) / 1e9,
            "inflate_gbs": U * K / (ms["inflate"] * 1e-3) / 1e9,
            "md5_gbs": (2 * U * K / (ms["md5"] * 1e-3) / 1e9) if do_md5 and ms["md5"] > 0 else None,
            "kernel_ms_per_step": {k: v / K for k, v in ms.items()},
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg_bytes / max(nl[dom], 1), "launches": nl[dom]},
            "e2e": {"value": U_all * K / (e2e_ms * 1e-3) / 1e9, "unit": UNIT, "h2d_bytes_per_step": int(U + state["C"]),
                    "d2h_bytes_per_step": int(U + state["C"]), "ms_per_step": e2e_ms / K},
            "gpu_launches": int(launches_all),
            "zwz_offset_exchange": exchange,
            "clocks": clocks,
        }
        if not args.no_cpu_baseline:
            target = int(args.cpu_sample_mb * 1e6) if args.cpu_sample_mb else int(min(8e6 * cores, 400e6))
            sb, so, what = sample_of(buf, foffs, target)
            scoff, sclen, _, _ = corpus.chunk_table(so)
            dt, comp, ok = cpu_step(sb, so, scoff, sclen, cores, do_md5)
            # our size on the very same sample, for the ratio criterion
            sres = ctx.deflate_batch(sb, scoff, sclen, args.level)[2]
            ours = int(sres["len0"].sum() + sres["len1"].sum())
            line["cpu_baseline"] = {"value": len(sb) / dt / 1e9, "unit": UNIT, "cores": cores, "kind": "port",
                                    "what": "the reference's call sequences (zlib L6 deflate, zlib inflate, OpenSSL MD5: oracle_ref_*) over the sample's chunks in memory, all host threads",
                                    "sample": what, "ratio": len(sb) / max(comp, 1), "roundtrip_ok": ok}
            line["size_vs_zlib6"] = ours / max(comp, 1)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
