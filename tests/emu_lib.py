"""Test-only: build + load tests/simt/libzwz_emu.so (the csrc kernels compiled for the CPU SIMT emulator)."""
import ctypes
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SIMT = os.path.join(ROOT, "tests", "simt")


def emu_context():
    import sys
    sys.path.insert(0, ROOT)
    import zwz_b200
    subprocess.run(["make", "-C", SIMT, "libzwz_emu.so"], check=True, capture_output=True)
    lib = zwz_b200.load_library(os.path.join(SIMT, "libzwz_emu.so"))
    return zwz_b200.Context(0, library=lib)
