"""Signature parity report: every public function / method of the reference package (parsed with ``ast`` - the reference is never
imported) against the callable of the same dotted name here.  Prints the reference parameters this library does not accept.

    python tools/check_signatures.py [/root/reference/flashinfer]

A parameter the reference accepts and this library lacks is a TypeError for a user who switches; the report is the work list.
Modules that only exist for other architectures / backends are listed in SKIP."""
from __future__ import annotations

import ast
import importlib
import inspect
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

SKIP_UNUSED = {"aot", "__main__", "_build_meta", "version", "tllm_enums", "tllm_utils", "jit", "cute_dsl", "triton", "gdn_kernels", "data", "profiler", "cli", "collect_env",
        "logits_processor", "testing", "tuning_configs", "artifacts", "compilation_context", "api_logging", "trace", "fi_trace", "dsv3_ops", "fused_moe", "gemm",
        "comm", "mamba", "quantization", "attention", "mla", "norm", "parallel_attention", "diffusion_ops", "topk"}


def ref_functions(path: Path):
    """{qualified name: [param names]} for top-level public defs and public methods of public classes of one source file."""
    tree = ast.parse(path.read_text())
    out = {}

    def params(fn):
        a = fn.args
        names = [x.arg for x in a.posonlyargs + a.args + a.kwonlyargs]
        return [n for n in names if n not in ("self", "cls")], a.vararg is not None, a.kwarg is not None

    for node in tree.body:
        if isinstance(node, (ast.FunctionDef, ast.AsyncFunctionDef)) and not node.name.startswith("_"):
            out[node.name] = params(node)
        elif isinstance(node, ast.ClassDef) and not node.name.startswith("_"):
            for sub in node.body:
                if isinstance(sub, ast.FunctionDef) and (not sub.name.startswith("_") or sub.name == "__init__"):
                    out[f"{node.name}.{sub.name}"] = params(sub)
    return out


def main() -> int:
    root = Path(sys.argv[1] if len(sys.argv) > 1 else "/root/reference/flashinfer")
    files = sorted(p for p in root.rglob("*.py") if "gdn_kernels" not in p.parts and "data" not in p.parts and "triton" not in p.parts
                   and "cute_dsl" not in p.parts and "jit" not in p.parts and "tuning_configs" not in p.parts)
    n_checked = n_missing_fn = n_gap = 0
    for f in files:
        rel = f.relative_to(root).with_suffix("")
        mod_name = ".".join(rel.parts)
        if mod_name.endswith("__init__"):
            mod_name = mod_name[: -len(".__init__")] if "." in mod_name else ""
        if rel.parts[-1].startswith("_") and rel.parts[-1] != "__init__":
            continue
        try:
            mod = importlib.import_module("flashinfer_b200" + ("." + mod_name if mod_name else ""))
        except Exception:
            continue
        for qual, (names, _, _) in ref_functions(f).items():
            obj = mod
            try:
                for part in qual.split("."):
                    obj = getattr(obj, part)
            except AttributeError:
                n_missing_fn += 1
                print(f"MISSING  {mod_name}.{qual}")
                continue
            try:
                sig = inspect.signature(inspect.unwrap(obj) if callable(obj) else obj)
            except (TypeError, ValueError):
                continue
            n_checked += 1
            ours = sig.parameters
            if any(p.kind is inspect.Parameter.VAR_KEYWORD for p in ours.values()):
                continue
            lack = [n for n in names if n not in ours]
            if lack:
                n_gap += 1
                print(f"PARAMS   {mod_name}.{qual}: {', '.join(lack)}")
    print(f"{n_checked} callables compared, {n_gap} with missing parameters, {n_missing_fn} missing callables")
    return 0


if __name__ == "__main__":
    sys.exit(main())
