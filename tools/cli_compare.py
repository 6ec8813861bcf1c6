#!/usr/bin/env python
"""End-to-end CLI comparison on the GPU box: the reference binary (oracle/_ref/main_ref, CPU: zlib + OpenSSL, P emulated MPI
ranks) against this repo's `main` (one B200 per rank), on a C1-shaped synthetic tree (BASELINE.json configs[0], scaled).

Prints one JSON object: wall-clock seconds of `Time Taken` for both, end-to-end MB/s, archive sizes, and the cross checks
  (1) the reference decompresses OUR archives to the original tree with "MD5 match" everywhere,
  (2) we decompress the REFERENCE's archives to the same bytes as the reference does.
usage: tools/cli_compare.py [--mb 500] [--ranks 2] [--out gpurun_out/cli_compare.json]
"""
import argparse
import filecmp
import json
import os
import shutil
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools import corpus  # noqa: E402

MAIN_REF = os.path.join(ROOT, "oracle", "_ref", "main_ref")
MAIN = os.path.join(ROOT, "parallel-data-compression-and-decompression_b200", "host", "main")


def run_ranks(binary, args, ranks, env_of_rank):
    t0 = time.time()
    procs = []
    for r in range(ranks):
        e = dict(os.environ)
        e.update(env_of_rank(r))
        procs.append(subprocess.Popen([binary] + args, env=e, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
        if r == 0 and ranks > 1:
            time.sleep(0.5)  # rank 0 publishes the record file first (the stub has no barrier)
    outs = [p.communicate()[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs[-1][-2000:]
    return time.time() - t0, outs


def same_tree(a, b):
    stack = [filecmp.dircmp(a, b)]
    n = 0
    while stack:
        d = stack.pop()
        if d.left_only or d.right_only:
            return False, n
        for f in d.common_files:
            if not filecmp.cmp(os.path.join(d.left, f), os.path.join(d.right, f), shallow=False):
                return False, n
            n += 1
        stack += list(d.subdirs.values())
    return True, n


def du(path):
    return sum(os.path.getsize(os.path.join(d, f)) for d, _, fs in os.walk(path) for f in fs)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=int, default=500)
    ap.add_argument("--ranks", type=int, default=2)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "cli_compare.json"))
    ap.add_argument("--tmp", default=None)
    ap.add_argument("--shape", default="c1", choices=["c1", "c2"], help="c1: 1000 mixed files up to 16 MiB; c2: image-like files of ~6.7 KB, 1000 per directory")
    a = ap.parse_args()
    tmp = tempfile.mkdtemp(prefix="zwz_cli_", dir=a.tmp)
    res = {"config": f"C1-shaped tree, ~{a.mb} MB, mixed T/S/I/R, sizes log-uniform 4 KiB..16 MiB", "ranks": a.ranks, "cores": os.cpu_count()}
    try:
        src = os.path.join(tmp, "w", "src")
        os.makedirs(src)
        specs, tot = [], 0
        pool = corpus.c1_specs(1000, corpus.BASE_SEED) if a.shape == "c1" else corpus.c2_specs(370_000, corpus.BASE_SEED)
        for s in pool:
            if tot >= a.mb * 1e6:
                break
            specs.append(s)
            tot += s.size
        corpus.write_tree(src, specs)
        if a.shape == "c2":
            res["config"] = f"C2-shaped tree, ~{a.mb} MB of image-like files (median 5.5 KiB), 1000 per directory"
        res["files"], res["bytes"] = len(specs), tot

        # ---- reference on the CPU: P emulated ranks, then 1-process decompress
        bc = os.path.join(tmp, "bc")
        os.makedirs(bc)
        ref_arch = os.path.join(tmp, "ref_arch")
        t, _ = run_ranks(MAIN_REF, ["compress", src, ref_arch], a.ranks,
                         lambda r: {"ZWZ_STUB_SIZE": str(a.ranks), "ZWZ_STUB_RANK": str(r), "ZWZ_STUB_DIR": bc})
        res["ref_compress_s"] = t
        ref_out = os.path.join(tmp, "ref_out")
        t, outs = run_ranks(MAIN_REF, ["decompress", ref_arch, ref_out], 1, lambda r: {})
        res["ref_decompress_s"] = t
        res["ref_archive_bytes"] = du(ref_arch)
        res["ref_verdicts"] = {"match": outs[0].count("MD5 match for file"), "mismatch": outs[0].count("MD5 mismatch for file")}

        # ---- ours: one process per GPU
        our_arch = os.path.join(tmp, "our_arch")
        import torch
        ngpu = max(1, torch.cuda.device_count())
        our_ranks = min(a.ranks, ngpu)
        t, outs = run_ranks(MAIN, ["compress", src, our_arch], our_ranks,
                            lambda r: {"ZWZ_WORLD": str(our_ranks), "ZWZ_RANK": str(r), "ZWZ_TIMING": os.environ.get("ZWZ_TIMING", "1")})
        res["our_compress_s"], res["our_ranks"] = t, our_ranks
        res["our_compress_phases"] = [l for o in outs for l in o.splitlines() if "[zwz timing]" in l]
        our_out = os.path.join(tmp, "our_out")
        t, outs = run_ranks(MAIN, ["decompress", our_arch, our_out], 1, lambda r: {"ZWZ_TIMING": os.environ.get("ZWZ_TIMING", "1")})
        res["our_decompress_s"] = t
        res["our_decompress_phases"] = [l for o in outs for l in o.splitlines() if "[zwz timing]" in l]
        res["our_archive_bytes"] = du(our_arch)
        res["our_verdicts"] = {"match": outs[0].count("MD5 match for file"), "mismatch": outs[0].count("MD5 mismatch for file")}
        ok, n = same_tree(src, our_out)
        res["our_roundtrip_identical"] = ok

        # ---- cross checks
        x1 = os.path.join(tmp, "x1")
        _, outs = run_ranks(MAIN_REF, ["decompress", our_arch, x1], 1, lambda r: {})
        ok1, _ = same_tree(src, x1)
        res["ref_reads_ours"] = {"identical_to_source": ok1, "md5_match": outs[0].count("MD5 match for file"),
                                 "md5_mismatch": outs[0].count("MD5 mismatch for file")}
        x2 = os.path.join(tmp, "x2")
        _, outs = run_ranks(MAIN, ["decompress", ref_arch, x2], 1, lambda r: {})
        ok2, _ = same_tree(ref_out, x2)
        if not ok2:  # diagnostics: which files differ, where
            diffs = []
            for d, _, fs in os.walk(ref_out):
                for f in fs:
                    pa = os.path.join(d, f)
                    pb = os.path.join(x2, os.path.relpath(pa, ref_out))
                    if not os.path.exists(pb):
                        diffs.append({"file": os.path.relpath(pa, ref_out), "missing": True})
                        continue
                    a_, b_ = open(pa, "rb").read(), open(pb, "rb").read()
                    if a_ != b_:
                        k = next((i for i in range(min(len(a_), len(b_))) if a_[i] != b_[i]), min(len(a_), len(b_)))
                        diffs.append({"file": os.path.relpath(pa, ref_out), "ref_size": len(a_), "our_size": len(b_), "first_diff": k,
                                      "src_size": os.path.getsize(os.path.join(src, os.path.relpath(pa, ref_out)))})
            res["we_read_refs_diffs"] = diffs[:40]
            keep = os.path.join(ROOT, "gpurun_out", "diff_case")
            shutil.rmtree(keep, ignore_errors=True)
            os.makedirs(keep, exist_ok=True)
            if diffs and "ref_size" in diffs[0] and diffs[0]["src_size"] < 6_000_000:   # keep the smallest differing source file for a local repro
                small = min((d_ for d_ in diffs if "src_size" in d_), key=lambda d_: d_["src_size"])
                shutil.copy(os.path.join(src, small["file"]), os.path.join(keep, "source.bin"))
                res["kept_case"] = small
        res["we_read_refs"] = {"identical_to_ref_output": ok2, "md5_match": outs[0].count("MD5 match for file"),
                               "md5_mismatch": outs[0].count("MD5 mismatch for file")}
        res["size_vs_ref"] = res["our_archive_bytes"] / res["ref_archive_bytes"]
        for k in ("ref_compress", "ref_decompress", "our_compress", "our_decompress"):
            res[k + "_MBps"] = tot / 1e6 / res[k + "_s"]
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    json.dump(res, open(a.out, "w"), indent=1)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
