#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ by RUNNING THE UNMODIFIED REFERENCE (oracle/_ref/main_ref,
built by `make -C oracle` from /root/reference with the single-process MPI stub).

Run in the build container only (needs /root/reference to have been compiled); the outputs are committed:

  edge_r1/compressed_0.zwz            main_ref compress, 1 rank, SURVEY.md §8(c) edge-case corpus
  edge_r2/compressed_{0,1}.zwz        same corpus, 2 emulated ranks (size-desc round-robin split)
  foreign/foreign.zwz                 hand-built archive: 32 768-byte stored records, Z_FIXED / Z_HUFFMAN_ONLY / Z_RLE /
                                      level-0..9 streams, out-of-order records, one record inflating to 1 000 000 bytes
  manifest.json                       md5+size of every input file, and md5+size of every file main_ref decompress
                                      wrote for each archive set (that IS the byte-exact target, including the short
                                      output the reference produces for its own truncated records), plus its verdicts

Usage: python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import shutil
import subprocess
import sys
import tempfile
import zlib

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from tools import corpus  # noqa: E402
import zwz_format  # noqa: E402

MAIN_REF = os.path.join(ROOT, "oracle", "_ref", "main_ref")


def md5_tree(root):
    out = {}
    for d, _, files in os.walk(root):
        for f in files:
            p = os.path.join(d, f)
            b = open(p, "rb").read()
            out[os.path.relpath(p, root)] = {"size": len(b), "md5": hashlib.md5(b).hexdigest()}
    return dict(sorted(out.items()))


def run_ref(args, env_extra=None):
    env = dict(os.environ)
    env.update(env_extra or {})
    r = subprocess.run([MAIN_REF] + args, env=env, capture_output=True, text=True, check=True)
    return r.stdout + r.stderr


def verdicts(log):
    return {"match": log.count("MD5 match for file"), "mismatch": log.count("MD5 mismatch for file")}


def build_foreign(path):
    """Streams the reference never writes but its reader accepts (SURVEY.md §5.1 reader constraints)."""
    import numpy as np
    recs = []
    rng = np.random.Generator(np.random.Philox(key=[596, 77]))

    def md5hex(b):
        return hashlib.md5(b).hexdigest().encode()

    def comp(b, level=6, strategy=zlib.Z_DEFAULT_STRATEGY, wbits=15):
        c = zlib.compressobj(level, zlib.DEFLATED, wbits, 8, strategy)
        return c.compress(b) + c.flush()

    # (a) 32 768-byte stored records (level 0), 3 of them + empty tail
    a = rng.integers(0, 256, size=3 * 32768, dtype=np.uint8).tobytes()
    for i in range(3):
        recs.append(zwz_format.Record("stored/three.bin", i, False, comp(a[i * 32768:(i + 1) * 32768], 0), None))
    recs.append(zwz_format.Record("stored/three.bin", 3, True, comp(b"", 6), md5hex(a)))
    # (b) one record inflating to 1 000 000 bytes
    big = (b"0123456789abcdef" * 4096)[:50000] * 20
    recs.append(zwz_format.Record("big/million.txt", 0, True, comp(big, 9), md5hex(big)))
    assert len(recs[-1].payload) <= 65535
    # (c) strategies / levels on text, one file with records written OUT OF ORDER
    txt = corpus.gen_text(5 * 20000, 596, 4242).tobytes()
    parts = [txt[i * 20000:(i + 1) * 20000] for i in range(5)]
    enc = [comp(parts[0], 6, zlib.Z_FIXED), comp(parts[1], 6, zlib.Z_HUFFMAN_ONLY), comp(parts[2], 6, zlib.Z_RLE),
           comp(parts[3], 1), comp(parts[4], 9, wbits=9)]
    order = [2, 0, 1, 4, 3]  # last chunk (4) arrives before chunk 3
    for k in order:
        recs.append(zwz_format.Record("mix/ooo.txt", k, k == 4, enc[k], md5hex(txt) if k == 4 else None))
    # (d) multi-block stream with sync flushes (stored empty blocks inside) and a fixed-Huffman tiny file
    c = zlib.compressobj(6)
    mb = b"".join(c.compress(parts[i][:3000]) + c.flush(zlib.Z_SYNC_FLUSH) for i in range(4)) + c.flush()
    raw_mb = b"".join(parts[i][:3000] for i in range(4))
    recs.append(zwz_format.Record("mix/multiblock.txt", 0, True, mb, md5hex(raw_mb)))
    recs.append(zwz_format.Record("mix/tiny", 0, True, comp(b"a"), md5hex(b"a")))
    # (e) bitmap-like data at level 4 (different parser) in two in-order records
    bm = corpus.gen_bitmap_like(90000, 596, 4343).tobytes()
    recs.append(zwz_format.Record("mix/bmp.raw", 0, False, comp(bm[:65535], 4), None))
    recs.append(zwz_format.Record("mix/bmp.raw", 1, True, comp(bm[65535:], 4), md5hex(bm)))
    open(path, "wb").write(zwz_format.serialize(recs))


def main():
    if not os.path.exists(MAIN_REF):
        sys.exit("build oracle/_ref/main_ref first: make -C oracle")
    manifest = {}
    tmp = tempfile.mkdtemp(prefix="zwz_golden_")
    try:
        src = os.path.join(tmp, "work", "edge")
        os.makedirs(src)
        corpus.edge_case_tree(src)
        manifest["inputs"] = md5_tree(src)

        # ---- 1 rank
        dst1 = os.path.join(tmp, "arch1")
        log = run_ref(["compress", src, dst1])
        out1 = os.path.join(tmp, "out1")
        dlog = run_ref(["decompress", dst1, out1])
        manifest["edge_r1"] = {"archives": sorted(os.listdir(dst1)), "outputs": md5_tree(out1), "verdicts": verdicts(dlog),
                               "record_file": open(os.path.join(tmp, "work", "sorted_files_by_size.txt")).read().split("\n")[:-1]}
        shutil.rmtree(os.path.join(HERE, "edge_r1"), ignore_errors=True)
        shutil.copytree(dst1, os.path.join(HERE, "edge_r1"))

        # ---- 2 emulated ranks (rank 0 first: it writes the record file and the bcast files)
        dst2 = os.path.join(tmp, "arch2")
        bc = os.path.join(tmp, "bc")
        os.makedirs(bc)
        for r in (0, 1):
            run_ref(["compress", src, dst2], {"ZWZ_STUB_SIZE": "2", "ZWZ_STUB_RANK": str(r), "ZWZ_STUB_DIR": bc})
        out2 = os.path.join(tmp, "out2")
        dlog = run_ref(["decompress", dst2, out2])
        manifest["edge_r2"] = {"archives": sorted(os.listdir(dst2)), "outputs": md5_tree(out2), "verdicts": verdicts(dlog)}
        shutil.rmtree(os.path.join(HERE, "edge_r2"), ignore_errors=True)
        shutil.copytree(dst2, os.path.join(HERE, "edge_r2"))

        # ---- foreign archive
        fdir = os.path.join(HERE, "foreign")
        shutil.rmtree(fdir, ignore_errors=True)
        os.makedirs(fdir)
        build_foreign(os.path.join(fdir, "foreign.zwz"))
        out3 = os.path.join(tmp, "out3")
        dlog = run_ref(["decompress", fdir, out3])
        manifest["foreign"] = {"archives": ["foreign.zwz"], "outputs": md5_tree(out3), "verdicts": verdicts(dlog)}

        json.dump(manifest, open(os.path.join(HERE, "manifest.json"), "w"), indent=1, sort_keys=True)
        print(json.dumps({k: v.get("verdicts") for k, v in manifest.items() if isinstance(v, dict) and "verdicts" in v}))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()
