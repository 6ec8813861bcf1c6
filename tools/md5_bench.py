#!/usr/bin/env python
"""MD5 kernel micro-benchmark on the GPU box: few long files (the C1 shape) vs many short ones. Prints kernel ms (events)."""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import zwz_b200 as zwz  # noqa: E402


def run(ctx, d, name, off, ln, host=None):
    ctx.profile_enable(True)
    ctx.profile_read(reset=True)
    for _ in range(3):
        dg = ctx.md5_batch_device(d, off, ln)
    ms, launches = ctx.profile_read(reset=True)["md5"]
    t = ms / launches
    tot = int(np.sum(ln))
    print(f"{name}: n={len(off)} total={tot/1e6:.1f} MB max={int(np.max(ln))/1e6:.2f} MB  {t:.2f} ms  {tot/t/1e6:.2f} GB/s  "
          f"longest-file rate {int(np.max(ln))/t/1e3:.1f} MB/s", flush=True)
    if host is not None:
        i = int(np.argmax(ln))
        assert dg[i].tobytes() == hashlib.md5(host[int(off[i]):int(off[i]) + int(ln[i])].tobytes()).digest()


def main():
    ctx = zwz.Context(0)
    rng = np.random.default_rng(1)
    total = 1 << 31
    host = rng.integers(0, 256, total // 8, dtype=np.uint8)
    host = np.tile(host, 8)
    d = ctx.malloc_device(total + 64)
    ctx.h2d(d, host)
    # one long file; 32 long files; 1000 log-uniform files; the same sorted by size; skewed starts
    run(ctx, d, "1 x 16 MiB", np.array([0], dtype=np.uint64), np.array([1 << 24], dtype=np.uint64), host)
    run(ctx, d, "1 x 16 MiB skew 1", np.array([1], dtype=np.uint64), np.array([1 << 24], dtype=np.uint64), host)
    run(ctx, d, "32 x 16 MiB", (np.arange(32, dtype=np.uint64) << np.uint64(24)), np.full(32, 1 << 24, dtype=np.uint64), host)
    sizes = np.exp(rng.uniform(np.log(4096), np.log(1 << 24), 1000)).astype(np.uint64)
    sizes = (sizes * np.uint64(total - 64) // np.uint64(sizes.sum() + 1)) if sizes.sum() > total - 64 else sizes
    off = np.zeros(1000, dtype=np.uint64)
    off[1:] = np.cumsum(sizes)[:-1]
    run(ctx, d, "1000 log-uniform (C1 shape)", off, sizes, host)
    o = np.argsort(-sizes.astype(np.int64), kind="stable")
    run(ctx, d, "1000 log-uniform, size-sorted", off[o], sizes[o], host)
    n = 200000
    ln = np.full(n, 7000, dtype=np.uint64)
    run(ctx, d, "200000 x 7000 B", np.arange(n, dtype=np.uint64) * np.uint64(7001), ln, host)
    ctx.free_device(d)


if __name__ == "__main__":
    main()
