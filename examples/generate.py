"""Generate with one of the op-by-op engines:  python examples/generate.py --family gemma2_9b --tiny --device cpu --new-tokens 8

Random-init weights (there are no checkpoints in this environment): the point is the serving path - page allocation, one batched
prefill, batched greedy decode - on this library's kernels (``--device cuda``) or their eager paths (``--device cpu``)."""
import argparse
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from flashinfer_b200.models import TransformerConfig, TransformerDecodeEngine  # noqa: E402
from flashinfer_b200.models.serving import PagedKVAllocator, generate  # noqa: E402


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--family", default="qwen3_8b", choices=["plain", "mixtral_8x7b", "qwen3_8b", "qwen3_30b_a3b", "gemma2_9b"])
    ap.add_argument("--tiny", action="store_true", help="test-size dimensions (the full presets need the real memory footprint)")
    ap.add_argument("--device", default="cuda" if torch.cuda.is_available() else "cpu")
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--prompt-len", type=int, default=24)
    ap.add_argument("--new-tokens", type=int, default=16)
    ap.add_argument("--page-size", type=int, default=16)
    a = ap.parse_args()
    cfg = TransformerConfig() if a.family == "plain" else getattr(TransformerConfig, a.family)()
    if a.tiny:
        cfg = cfg.tiny()
    pages = a.batch * (-(-(a.prompt_len + a.new_tokens) // a.page_size)) + 4
    eng = TransformerDecodeEngine(cfg, a.batch, pages, a.page_size, a.device, torch.bfloat16)
    g = torch.Generator().manual_seed(0)
    prompts = [torch.randint(0, cfg.vocab_size, (a.prompt_len - i % 3,), generator=g).tolist() for i in range(a.batch)]
    t0 = time.perf_counter()
    outs = generate(eng, prompts, a.new_tokens, PagedKVAllocator(pages, a.page_size))
    if a.device == "cuda":
        torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    for i, o in enumerate(outs):
        print(f"request {i}: prompt {len(prompts[i])} tokens -> {o}")
    print(f"{cfg.name} on {a.device}: {sum(len(o) for o in outs)} tokens in {dt:.2f} s (wall clock incl. planning; not a benchmark)")


if __name__ == "__main__":
    main()
