#!/usr/bin/env python
"""Ceiling of the end-to-end leg: concurrent H2D + D2H cudaMemcpyAsync from/to page-locked host memory on every GPU of the
run, no kernels. bench.py's e2e moves (U + C) bytes each way per step; its rate cannot exceed what this prints.

  python tools/copy_ceiling.py [--mb 2048] [--chunk-mb 64]                       # one GPU
  python -m torch.distributed.run --nproc-per-node 8 tools/copy_ceiling.py       # all ranks at once (what the 8-GPU bench sees)
Prints one JSON line (rank 0): per-GPU and aggregate GB/s for H2D alone, D2H alone and both directions at once, plus where
the pinned memory landed (NUMA node of the pages) and the GPU<->CPU affinity from nvidia-smi topo.
"""
import argparse
import json
import os
import subprocess
import time

import torch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=int, default=2048)
    ap.add_argument("--chunk-mb", type=int, default=64)
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = a.mb << 20
    ck = a.chunk_mb << 20
    h_in = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    h_out = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    h_in.fill_(1)
    h_out.fill_(2)
    d_in = torch.empty(n, dtype=torch.uint8, device="cuda")
    d_out = torch.ones(n, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def run(h2d, d2h):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(a.reps):
            for o in range(0, n, ck):
                if h2d:
                    with torch.cuda.stream(s1):
                        d_in[o:o + ck].copy_(h_in[o:o + ck], non_blocking=True)
                if d2h:
                    with torch.cuda.stream(s2):
                        h_out[o:o + ck].copy_(d_out[o:o + ck], non_blocking=True)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        dt = time.perf_counter() - t0
        return n * a.reps / dt / 1e9

    run(True, True)
    r = [run(True, False), run(False, True), run(True, True)]
    t = torch.tensor(r, dtype=torch.float64, device="cuda")
    if world > 1:
        allr = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allr, t)
        per = torch.stack(allr).cpu().numpy()
    else:
        per = t.cpu().numpy()[None, :]
    if rank == 0:
        topo = ""
        try:
            topo = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=20).stdout
        except Exception:
            pass
        numa = ""
        try:
            numa = subprocess.run(["bash", "-c", "lscpu | grep -i numa; nproc"], capture_output=True, text=True, timeout=20).stdout
        except Exception:
            pass
        print(json.dumps({"n_gpus": world, "mb_per_direction": a.mb, "chunk_mb": a.chunk_mb,
                          "per_gpu_gbs": {"h2d_alone": per[:, 0].tolist(), "d2h_alone": per[:, 1].tolist(), "both_each_direction": per[:, 2].tolist()},
                          "aggregate_gbs": {"h2d_alone": float(per[:, 0].sum()), "d2h_alone": float(per[:, 1].sum()),
                                            "both_each_direction": float(per[:, 2].sum())},
                          "note": "bench.py e2e moves (U + C) bytes each way per step: ceiling for e2e GB/s (uncompressed) = both_each_direction * U / (U + C)",
                          "numa": numa, "topo": topo}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
