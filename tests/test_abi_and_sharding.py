"""CPU-only checks: the C-ABI library exports every symbol include/zwz_cuda.h declares (no compute without a GPU), the
product package refuses to run without a GPU (no silent fallback), and the multi-GPU path (size-descending round-robin
deal + the one collective: an all-gather of per-rank counters) under gloo with world_size 2."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "parallel-data-compression-and-decompression_b200")


def _declared(header):
    src = open(header).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = set(re.findall(r"\b(zwz_[a-z0-9_]+)\s*\(", src))
    names.discard("zwz_deflate_bound")  # static inline
    return sorted(names)


def test_cuda_library_exports_every_declared_symbol():
    so = os.path.join(PKG, "csrc", "libzwz_cuda.so")
    assert os.path.exists(so), "run __graft_entry__.build() first"
    lib = ctypes.CDLL(so)
    names = _declared(os.path.join(ROOT, "include", "zwz_cuda.h"))
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/zwz_cuda.h but not exported"
    assert lib.zwz_abi_version() == 1


def test_host_library_exports():
    so = os.path.join(PKG, "host", "libzwz_host.so")
    assert os.path.exists(so)
    lib = ctypes.CDLL(so)
    for n in ("zwz_host_compress", "zwz_host_decompress", "zwz_host_md5_of_file", "zwz_host_sort_files_by_size", "zwz_host_last_stats"):
        assert hasattr(lib, n)


def test_no_cpu_fallback_in_product_path():
    """On a box without a GPU the product must fail loudly, not fall back to anything."""
    import zwz_b200
    lib = zwz_b200.load_library()
    if lib.zwz_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(zwz_b200.ZwzError):
        zwz_b200.Context(0)
    # and nothing under the package imports the oracle or the emulator
    for d, _, files in os.walk(PKG):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                s = open(os.path.join(d, f), errors="ignore").read()
                assert "liboracle" not in s and "oracle_lib" not in s and "libzwz_emu" not in s, os.path.join(d, f)


def test_chunk_table_matches_reference_rule():
    import zwz_b200
    sizes = [0, 1, 65534, 65535, 65536, 131070, 131071, 200000]
    offs = np.concatenate([[0], np.cumsum(sizes)])
    coff, clen, cfile, cseq = zwz_b200.chunk_table(offs)
    for i, s in enumerate(sizes):
        lens = clen[cfile == i]
        assert len(lens) == s // 65535 + 1                      # compression.cpp:52-64 [probed in SURVEY.md §5.1]
        assert list(lens[:-1]) == [65535] * (len(lens) - 1) and lens[-1] == s % 65535
        assert list(cseq[cfile == i]) == list(range(len(lens)))
    assert list(zwz_b200.deflate_bound([0, 1, 65535])) == [48, 64, 65584]


WORKER = r'''
import os, sys, json
sys.path.insert(0, {root!r})
import numpy as np, torch, torch.distributed as dist
from tools import corpus
dist.init_process_group("gloo", rank=int(os.environ["RANK"]), world_size=int(os.environ["WORLD_SIZE"]))
rank, world = dist.get_rank(), dist.get_world_size()
sizes = corpus.c2_sizes(1000 * world, 596)
order = np.argsort(-sizes, kind="stable")
mine = order[rank::world]                       # compression.cpp:31-41 with GPU index in place of MPI rank
cnt = torch.tensor([float(sizes[mine].sum()), float(len(mine))], dtype=torch.float64)
allc = [torch.zeros_like(cnt) for _ in range(world)]
dist.all_gather(allc, cnt)                      # the only collective of the path: per-rank counters
tot = torch.stack(allc).sum(0)
t = torch.tensor([1.0 + rank], dtype=torch.float64)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print(json.dumps({{"bytes": tot[0].item(), "files": tot[1].item(), "tmax": t.item(), "mine0": int(mine[0]), "n_mine": len(mine)}}))
dist.destroy_process_group()
'''


def test_two_rank_sharding_and_counter_exchange_gloo(tmp_path):
    from tools import corpus
    script = tmp_path / "w.py"
    script.write_text(WORKER.format(root=ROOT))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29611", str(script)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    import json
    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    d = json.loads(line)
    sizes = corpus.c2_sizes(2000, 596)
    assert d["bytes"] == float(sizes.sum()) and d["files"] == 2000.0 and d["tmax"] == 2.0 and d["n_mine"] == 1000
    assert d["mine0"] == int(np.argmax(sizes))   # rank 0 owns the largest file
    # the deal is balanced because it is size-descending round-robin: shards differ by at most the largest file
    order = np.argsort(-sizes, kind="stable")
    assert abs(int(sizes[order[0::2]].sum()) - int(sizes[order[1::2]].sum())) <= int(sizes.max())
