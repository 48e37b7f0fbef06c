"""One benchmark driver for every op that carries a trace template.

    python benchmarks/flashinfer_benchmark.py --list
    python benchmarks/flashinfer_benchmark.py --routine rmsnorm --set batch_size=4096 hidden_size=8192 --refcheck
    python benchmarks/flashinfer_benchmark.py --routine mm_bf16 --set M=64 N=28672 K=4096 --output_path out.csv
    python benchmarks/flashinfer_benchmark.py --testlist benchmarks/samples/decode_layer.txt --output_path out.csv

Parity: reference benchmarks/flashinfer_benchmark.py + benchmarks/routines/*.py (one argparse routine per op family, CSV
writer, ``--refcheck``, ``--testlist``).  Here a routine is generated from the op's TraceTemplate: ``init`` builds the inputs
for the requested axis sizes (defaults are the serving shapes named in the template), the bound API is the timed callable and
``--refcheck`` compares it with the template's reference.  Timing: CUDA events with a cold L2 (``--use_cuda_graph`` times a
graph of 10 calls over rotated inputs instead); on a machine without a GPU it falls back to wall clock so the harness
itself is testable (``tests/test_benchmark_harness_cpu.py``).  Reported: median / std ms, effective TB/s (bytes of every input and
output tensor once) and, for GEMM / attention templates, TFLOP/s; roofline fractions against MEASURED_PEAKS.json if present.
"""
from __future__ import annotations

import argparse
import csv
import json
import os
import shlex
import statistics
import sys
import time
from typing import Any, Dict, List, Optional

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

import flashinfer_b200 as fi  # noqa: E402
from flashinfer_b200.trace import FLAT_BINDINGS as BINDINGS, Const  # noqa: E402  (one row per concrete template)
from flashinfer_b200.trace.bindings import _resolve  # noqa: E402

COLUMNS = ["routine", "fi_api", "definition", "axes", "device", "median_ms", "std_ms", "iters", "bytes", "tb_per_sec", "flops",
           "tflops_per_sec", "frac_hbm_peak", "frac_bf16_peak", "refcheck", "timer"]


def routines() -> Dict[str, Any]:
    """routine name -> (module, attribute path, template): the template's name stem (``rmsnorm``, ``gqa_paged_decode`` ...)."""
    out = {}
    for mod, path, tpl in BINDINGS:
        if tpl.init is None:
            continue
        stem = tpl.name_fmt.split("{")[0].rstrip("_")
        for cut in ("_h", "_v", "_n", "_k", "_hq", "_q", "_e", "_in", "_ckv"):   # drop the dangling axis abbreviation
            if stem.endswith(cut) and tpl.name_fmt[len(stem):].startswith("{"):
                stem = stem[: -len(cut)]
                break
        out.setdefault(stem, (mod, path, tpl))
    return out


def _tensor_bytes(obj) -> int:
    if isinstance(obj, torch.Tensor):
        return obj.numel() * obj.element_size()
    if isinstance(obj, (tuple, list)):
        return sum(_tensor_bytes(o) for o in obj)
    return 0


def _flops(tpl, sizes: Dict[str, int], kwargs: Dict[str, Any]) -> Optional[float]:
    if tpl.op_type == "gemm" and {"M", "N", "K"} <= sizes.keys():
        return 2.0 * sizes["M"] * sizes["N"] * sizes["K"] * sizes.get("batch", 1)
    if tpl.op_type == "gemm" and "d_in" in sizes:
        return 2.0 * sizes.get("total_rows", 0) * sizes["d_in"] * sizes["d_out"]
    if tpl.op_type in ("gqa_paged", "gqa_ragged", "mla_paged"):
        w = kwargs.get("self")
        q_host, kv = getattr(w, "_qo_indptr_host", getattr(w, "_qo_host", None)), getattr(w, "_kv_lens_host", getattr(w, "_kvl_host", None))
        if kv is None:
            return None
        ql = (q_host[1:] - q_host[:-1]).double() if q_host is not None else torch.ones_like(kv).double()
        heads = sizes.get("num_qo_heads", sizes.get("num_heads", 1))
        d_qk = sizes.get("head_dim", 0) or (sizes.get("head_dim_ckv", 0) + sizes.get("head_dim_kpe", 0))
        d_vo = sizes.get("head_dim", 0) or sizes.get("head_dim_ckv", 0)
        return float((ql * kv.double()).sum()) * heads * 2.0 * (d_qk + d_vo)          # causal masking not discounted
    if tpl.op_type == "gqa_single":
        lq = sizes.get("qo_len", 1)
        return 4.0 * lq * sizes["kv_len"] * sizes["num_qo_heads"] * sizes["head_dim"]
    if tpl.op_type == "moe" and "intermediate_size" in sizes:
        return 6.0 * sizes["seq_len"] * sizes["top_k"] * sizes["hidden_size"] * sizes["intermediate_size"]
    return None


def _peaks() -> Dict[str, float]:
    from flashinfer_b200.testing import measured_peaks

    d = measured_peaks()
    gbs = d.get("hbm_gbs", d.get("hbm_gbps", 0.0)) or 0.0
    return {"hbm_tbs": float(gbs) / 1e3, "bf16_tflops": float(d.get("bf16_tflops", 0.0) or 0.0)}


def _clone(v):
    if isinstance(v, torch.Tensor):
        return v.clone()
    if isinstance(v, tuple):
        return tuple(_clone(x) for x in v)
    return v


def _time(call, device: str, iters: Optional[int], use_graph: bool, rotate: Optional[List[Dict[str, Any]]] = None):
    if device == "cpu" or not torch.cuda.is_available():
        call()
        n = iters or 5
        ts = []
        for _ in range(n):
            t0 = time.perf_counter()
            call()
            ts.append((time.perf_counter() - t0) * 1e3)
        return ts, "wall_clock"
    from flashinfer_b200.testing import bench_gpu_time_with_cuda_event, bench_gpu_time_with_cudagraph

    if use_graph:
        return bench_gpu_time_with_cudagraph(call, repeat_iters=iters), "cuda_graph"
    return bench_gpu_time_with_cuda_event(call, repeat_iters=iters, l2_flush=True), "cuda_events_cold_l2"


def run_case(name: str, sizes: Dict[str, int], device: str = "cuda", iters: Optional[int] = None, refcheck: bool = False,
             use_cuda_graph: bool = False, seed: int = 0) -> Dict[str, Any]:
    table = routines()
    if name not in table:
        raise SystemExit(f"unknown routine '{name}'; --list shows the {len(table)} available")
    mod, path, tpl = table[name]
    owner, attr = _resolve(mod, path)
    api = getattr(owner, attr)
    kwargs = tpl.make_inputs(device=device, seed=seed, **sizes)
    resolved = tpl.resolve_axes(kwargs)
    check = "skipped"
    if refcheck:
        probe = {k: _clone(v) for k, v in kwargs.items()}
        ref_in = {k: _clone(v) for k, v in kwargs.items()}
        expect = tpl.run_reference(ref_in)
        expect = list(expect) if isinstance(expect, (tuple, list)) else [expect]
        got = tpl.collect_outputs(api(**probe), probe)
        try:
            if tpl.compare is not None:
                tpl.compare(got, expect, ref_in)
            elif tpl.tolerance == "support":
                assert all(e[torch.arange(g.numel(), device=g.device), g.long()].all() for g, e in zip(got, expect))
            elif tpl.tolerance == "cos":
                for g, e in zip(got, expect):
                    assert torch.nn.functional.cosine_similarity(g.float().flatten(), e.float().flatten(), dim=0) > 0.99
            else:
                tol = {"exact": 0.0, "fp32": 1e-5, "bf16_norm": 2e-2, "bf16": 3e-2, "fp8_quant": 0.13}[tpl.tolerance]
                for g, e in zip(got, expect):
                    torch.testing.assert_close(g.float(), e.float(), atol=tol, rtol=tol)
            check = "pass"
        except AssertionError as exc:
            check = "FAIL: " + str(exc).splitlines()[0][:80]
    # in-place ops keep running on the same buffers: values drift but the work per call does not
    times, timer = _time(lambda: api(**kwargs), device, iters, use_cuda_graph)
    med = statistics.median(times)
    nbytes = sum(_tensor_bytes(v) for k, v in kwargs.items() if k != "self")
    out = api(**kwargs)
    nbytes += sum(_tensor_bytes(o) for o in (out if isinstance(out, (tuple, list)) else [out])
                  if not any(o is v for v in kwargs.values()))
    flops = _flops(tpl, resolved, kwargs)
    peaks = _peaks()
    tbs = nbytes / (med * 1e-3) / 1e12 if med > 0 else 0.0
    tfl = flops / (med * 1e-3) / 1e12 if flops and med > 0 else None
    return {"routine": name, "fi_api": f"{mod}.{path}", "definition": tpl.definition_name(resolved),
            "axes": json.dumps({a.name: resolved.get(a.name) for a in tpl.axes if resolved.get(a.name) is not None}), "device": device,
            "median_ms": round(med, 5), "std_ms": round(statistics.pstdev(times), 5), "iters": len(times), "bytes": nbytes,
            "tb_per_sec": round(tbs, 4), "flops": flops, "tflops_per_sec": round(tfl, 3) if tfl is not None else None,
            "frac_hbm_peak": round(tbs / peaks["hbm_tbs"], 3) if peaks["hbm_tbs"] else None,
            "frac_bf16_peak": round(tfl / peaks["bf16_tflops"], 3) if tfl is not None and peaks["bf16_tflops"] else None,
            "refcheck": check, "timer": timer}


def _parse_sets(items: List[str]) -> Dict[str, int]:
    out = {}
    for it in items or []:
        k, _, v = it.partition("=")
        out[k] = int(v)
    return out


def build_parser() -> argparse.ArgumentParser:
    ap = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    ap.add_argument("--list", action="store_true", help="print the routines with their axes and default sizes")
    ap.add_argument("--routine", "-R")
    ap.add_argument("--set", nargs="*", default=[], metavar="AXIS=N", help="axis sizes passed to the template's input builder")
    ap.add_argument("--device", default="cuda" if torch.cuda.is_available() else "cpu")
    ap.add_argument("--iters", type=int, default=None)
    ap.add_argument("--refcheck", action="store_true")
    ap.add_argument("--use_cuda_graph", action="store_true")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--output_path", default=None, help="append results to this CSV")
    ap.add_argument("--testlist", default=None, help="file with one set of arguments per line ('#' comments)")
    return ap


def _list() -> None:
    import inspect

    for name, (mod, path, tpl) in sorted(routines().items()):
        sig = inspect.signature(tpl.init).parameters
        axes = " ".join(f"{k}={v.default}" for k, v in sig.items() if k not in ("device", "seed"))
        print(f"{name:<40} {mod}.{path:<48} {axes}")


def main(argv: Optional[List[str]] = None) -> List[Dict[str, Any]]:
    ap = build_parser()
    args = ap.parse_args(argv)
    if args.list:
        _list()
        return []
    cases = []
    if args.testlist:
        with open(args.testlist) as f:
            for line in f:
                line = line.split("#", 1)[0].strip()
                if line:
                    cases.append(ap.parse_args(shlex.split(line)))
        for c in cases:                                           # file-level defaults come from the command line
            c.output_path = c.output_path or args.output_path
            c.device = args.device if "--device" not in (argv or sys.argv) else c.device
    elif args.routine:
        cases.append(args)
    else:
        ap.error("one of --list, --routine, --testlist is required")
    rows = []
    for c in cases:
        row = run_case(c.routine, _parse_sets(c.set), c.device, c.iters, c.refcheck, c.use_cuda_graph, c.seed)
        rows.append(row)
        print(json.dumps(row), flush=True)
        if c.output_path:
            new = not os.path.exists(c.output_path)
            with open(c.output_path, "a", newline="") as f:
                w = csv.DictWriter(f, fieldnames=COLUMNS)
                if new:
                    w.writeheader()
                w.writerow(row)
    return rows


if __name__ == "__main__":
    main()
