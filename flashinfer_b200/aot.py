"""Ahead-of-time build of every native module (reference flashinfer/aot.py:447-941).  With a single target
architecture there is no dtype x head-dim x arch cross product to enumerate: kernels dispatch on run-time dtype codes
and template instantiations live inside each translation unit, so AOT == build every ModuleSpec once."""
from __future__ import annotations

import argparse

from . import jit


def gen_all_modules():
    return list(jit.REGISTRY.values())


def main(argv=None) -> int:
    ap = argparse.ArgumentParser("flashinfer_b200.aot")
    ap.add_argument("--force", action="store_true")
    ap.add_argument("-v", "--verbose", action="store_true")
    ap.add_argument("-j", "--jobs", type=int, default=None)
    a = ap.parse_args(argv)
    built = jit.build_all(verbose=a.verbose, jobs=a.jobs, force=a.force)
    print(f"[aot] {len(built)} modules in {jit.LIB_DIR}")
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
