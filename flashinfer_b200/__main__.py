"""Command line interface:  python -m flashinfer_b200 <command>

Parity: reference flashinfer/__main__.py:64-393 (show-config, module-status, list-modules, clear-cache, build / AOT,
replay, export-compile-commands, list-/download-/clear-cubin)."""
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
    from .trace import Const, registered_templates

    for k, t in sorted(registered_templates().items()):
        if getattr(args, "verbose", False):
            axes = " ".join(("%s=const" if isinstance(a, Const) else "%s=var") % a.name for a in t.axes)
            print(f"{k:<90} {t.fi_api or '(unbound)':<70} ref={'yes' if t.reference else 'no '} init={'yes' if t.init else 'no '} {axes}")
        else:
            print(k)
    return 0


def _cmd_export_compile_commands(args) -> int:
    """compile_commands.json for clangd / IDEs: one entry per translation unit of every (or the named) native module
    (reference __main__.py:334 export-compile-commands)."""
    from . import jit

    names = args.modules or list(jit.REGISTRY)
    entries = []
    for name in names:
        spec = jit.REGISTRY[name]
        cmd = spec.nvcc_command()
        srcs = {str(p) for p in spec.source_paths()}
        shared = [c for c in cmd if c not in srcs]
        for src in sorted(srcs):
            entries.append({"directory": str(jit.CSRC), "file": src, "arguments": shared[:1] + ["-c", src] + shared[1:],
                            "output": str(spec.so_path)})
    out = args.output or "compile_commands.json"
    with open(out, "w") as f:
        json.dump(entries, f, indent=1)
    print(f"wrote {len(entries)} entries to {out}")
    return 0


def _cmd_list_cubins(args) -> int:
    """The reference downloads prebuilt cubins (trtllm-gen, cuDNN-frontend ...) from an artifact server; every kernel here is
    compiled from the in-tree sources, so the 'cubin store' is the in-tree library directory."""
    from . import artifacts, jit

    st = jit.module_status()
    for name, spec in jit.REGISTRY.items():
        size = spec.so_path.stat().st_size if spec.so_path.exists() else 0
        print(f"{name:<28} {st.get(name, '?'):<8} {size:>10} B  {spec.so_path}")
    print(f"downloadable artifacts: {len(getattr(artifacts, 'get_available_cubin_files', lambda *a, **k: [])())} (none: built from source)")
    return 0


def _cmd_download_cubin(args) -> int:
    print("nothing to download: flashinfer_b200 has no prebuilt-cubin dependencies; run `python -m flashinfer_b200 build`")
    return 0


def _cmd_clear_cubin(args) -> int:
    return _cmd_clear_cache(args)


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
    t = sub.add_parser("trace-templates", help="list the fi_trace templates (-v: bound API, reference / init presence, axes)")
    t.add_argument("-v", "--verbose", action="store_true")
    t.set_defaults(fn=_cmd_trace_templates)
    e = sub.add_parser("export-compile-commands", help="write compile_commands.json for the native modules")
    e.add_argument("modules", nargs="*")
    e.add_argument("-o", "--output", default=None)
    e.set_defaults(fn=_cmd_export_compile_commands)
    sub.add_parser("list-cubins").set_defaults(fn=_cmd_list_cubins)
    sub.add_parser("download-cubin").set_defaults(fn=_cmd_download_cubin)
    sub.add_parser("clear-cubin", help="same as clear-cache: the compiled libraries are the only binary artifacts").set_defaults(fn=_cmd_clear_cubin)
    args = ap.parse_args(argv)
    return args.fn(args)


if __name__ == "__main__":
    sys.exit(main())
