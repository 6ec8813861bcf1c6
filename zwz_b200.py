"""Import shim: the package directory name mandated for this repo contains hyphens, which Python's `import` statement
cannot spell. `import zwz_b200` gives the same module object as
importlib.import_module("parallel-data-compression-and-decompression_b200")."""
import importlib
import sys

_m = importlib.import_module("parallel-data-compression-and-decompression_b200")
sys.modules[__name__] = _m
