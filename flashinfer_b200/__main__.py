"""Command line interface:  python -m flashinfer_b200 <command>

Parity: reference flashinfer/__main__.py:64-393 (show-config, module-status, list-modules, clear-cache, build / AOT,
replay).  click is used when available, argparse otherwise."""
from __future__ import annotations

import argparse
import json
import os
import shutil
import sys


def _cmd_show_config(args) -> int:
    import torch

    from . import jit

    info = {
        "package": "flashinfer_b200",
        "torch": torch.__version__,
        "cuda_runtime": torch.version.cuda,
        "cuda_available": torch.cuda.is_available(),
        "device": torch.cuda.get_device_name(0) if torch.cuda.is_available() else None,
        "arch": "sm_100a",
        "nvcc": shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc",
        "lib_dir": str(getattr(jit, "LIB_DIR", "")),
        "env": {k: v for k, v in os.environ.items() if k.startswith(("FLASHINFER_", "FIB200_"))},
    }
    print(json.dumps(info, indent=1))
    return 0


def _cmd_module_status(args) -> int:
    from . import jit

    st = jit.module_status()
    w = max(len(k) for k in st)
    for k, v in st.items():
        print(f"{k:<{w}}  {v}")
    return 0


def _cmd_list_modules(args) -> int:
    from . import jit

    for name, spec in jit.REGISTRY.items():
        print(name, " ".join(spec.sources))
    return 0


def _cmd_build(args) -> int:
    from . import jit

    names = args.modules or None
    built = jit.build_all(verbose=args.verbose) if names is None else [jit.build_module(jit.REGISTRY[n], verbose=args.verbose) for n in names]
    print(f"built {len(built)} modules")
    return 0


def _cmd_clear_cache(args) -> int:
    from . import jit

    lib = str(getattr(jit, "LIB_DIR", ""))
    n = 0
    if lib and os.path.isdir(lib):
        for f in os.listdir(lib):
            if f.endswith((".so", ".hash")):
                os.remove(os.path.join(lib, f))
                n += 1
    print(f"removed {n} files from {lib}")
    return 0


def _cmd_replay(args) -> int:
    import flashinfer_b200  # noqa: F401 - registers the APIs

    from .api_logging import replay_from_dump, replay_sequence

    res = replay_sequence(args.path) if args.sequence else [replay_from_dump(args.path)]
    for r in res:
        print(r["api"], "match=" + str(r.get("match")))
    return 0 if all(r.get("match", True) for r in res) else 1


def _cmd_trace_templates(args) -> int:
    from .trace import registered_templates

    for k in registered_templates():
        print(k)
    return 0


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(prog="flashinfer_b200")
    sub = ap.add_subparsers(dest="cmd", required=True)
    sub.add_parser("show-config").set_defaults(fn=_cmd_show_config)
    sub.add_parser("module-status").set_defaults(fn=_cmd_module_status)
    sub.add_parser("list-modules").set_defaults(fn=_cmd_list_modules)
    b = sub.add_parser("build", help="AOT-compile native modules for sm_100a")
    b.add_argument("modules", nargs="*")
    b.add_argument("-v", "--verbose", action="store_true")
    b.set_defaults(fn=_cmd_build)
    sub.add_parser("clear-cache").set_defaults(fn=_cmd_clear_cache)
    r = sub.add_parser("replay", help="re-run dumped API calls (FLASHINFER_DUMP_DIR)")
    r.add_argument("path")
    r.add_argument("--sequence", action="store_true")
    r.set_defaults(fn=_cmd_replay)
    sub.add_parser("trace-templates").set_defaults(fn=_cmd_trace_templates)
    args = ap.parse_args(argv)
    return args.fn(args)


if __name__ == "__main__":
    sys.exit(main())
