"""The unified benchmark driver (reference benchmarks/test_flashinfer_benchmark.py): every sample line runs, reference checks
pass, the CSV has one well-formed row per case.  CPU run with wall-clock timing; the same code path times with CUDA events
on a GPU."""
import csv
import importlib.util
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load():
    spec = importlib.util.spec_from_file_location("fib200_benchmark", os.path.join(ROOT, "benchmarks", "flashinfer_benchmark.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_testlist_runs_and_writes_csv(tmp_path, capsys):
    bench = _load()
    out = tmp_path / "r.csv"
    rows = bench.main(["--testlist", os.path.join(ROOT, "benchmarks", "samples", "smoke_cpu.txt"), "--device", "cpu", "--output_path", str(out)])
    assert len(rows) == 6 and all(r["refcheck"] == "pass" for r in rows), [r["refcheck"] for r in rows]
    assert all(r["median_ms"] > 0 and r["bytes"] > 0 and r["timer"] == "wall_clock" for r in rows)
    gemm = next(r for r in rows if r["routine"] == "mm_bf16")
    assert gemm["flops"] == 2.0 * 8 * 64 * 128 and gemm["definition"] == "mm_bf16_n64_k128"
    dec = next(r for r in rows if r["routine"] == "gqa_paged_decode")
    assert dec["flops"] == 3 * 19 * 4 * 2.0 * 128                 # sum(q_len * kv_len) * heads * 2 * (d_qk + d_vo)
    with open(out) as f:
        table = list(csv.DictReader(f))
    assert [t["routine"] for t in table] == [r["routine"] for r in rows] and set(table[0]) == set(bench.COLUMNS)


def test_every_routine_has_defaults_and_samples_parse():
    bench = _load()
    table = bench.routines()
    assert len(table) >= 55 and {"rmsnorm", "mm_bf16", "gqa_paged_decode", "mla_paged", "fused_moe_bf16", "softmax"} <= set(table)
    ap = bench.build_parser()
    import shlex

    for name in ("decode_layer.txt", "smoke_cpu.txt"):
        for line in open(os.path.join(ROOT, "benchmarks", "samples", name)):
            line = line.split("#", 1)[0].strip()
            if line:
                a = ap.parse_args(shlex.split(line))
                assert a.routine in table, a.routine
                import inspect

                accepted = inspect.signature(table[a.routine][2].init).parameters
                assert set(bench._parse_sets(a.set)) <= set(accepted), (a.routine, a.set)
