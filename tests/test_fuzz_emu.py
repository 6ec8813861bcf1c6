"""One seeded round of each CPU fuzzer (tools/fuzz_deflate_emu.py, tools/fuzz_inflate_emu.py) as part of the suite: chunks
assembled from noise, text, copies of earlier bytes, tables, runs and ramps. Deflate: every stream inflates (zlib) to its input.
Inflate: zlib streams of random level/strategy/window, truncated and bit-flipped ones, and our own streams give the oracle's
bytes, count and status. Longer runs: `python tools/fuzz_deflate_emu.py 20 <seed>`."""
from tools import fuzz_deflate_emu, fuzz_inflate_emu


def test_fuzz_deflate_one_round(emu_ctx):
    assert fuzz_deflate_emu.run(1, 101)


def test_fuzz_inflate_one_round(emu_ctx):
    assert fuzz_inflate_emu.run(1, 102)
