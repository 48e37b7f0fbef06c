"""Decode engines assembled from this library's ops.

* :mod:`.llama` - the flagship: Llama-family step on the fused decode GEMM family (5 launches per layer, in-GEMM tensor-parallel
  all-reduce), what ``bench.py`` measures;
* :mod:`.deepseek` - DeepSeek-V2 / V3: absorbed Multi-head Latent Attention over a paged latent cache + grouped-top-k MoE with a shared expert;
* :mod:`.transformer` - a configurable GQA decoder: Mixtral / Qwen-MoE (renormalised top-k experts), Qwen3 (q / k norm), Gemma-2 / 3
  (soft-caps, sliding-window layers, post norms, GeGLU), plain Llama / Mistral;
* :mod:`.mamba2` - Mamba-2: rolling depthwise-conv state + ``selective_state_update`` recurrence + gated norm;
* :mod:`.gdn` - gated delta net (the recurrent layer of Qwen3-Next style hybrids) on ``gdn.gated_delta_rule_mtp`` with slot-addressed states.

The last four are op-by-op and device agnostic (native kernels on CUDA, eager paths on CPU) and are tested against plain PyTorch models."""
from .deepseek import DeepSeekConfig, DeepSeekDecodeEngine  # noqa: F401
from .gdn import GDNConfig, GDNDecodeEngine  # noqa: F401
from .llama import LlamaConfig, LlamaDecodeEngine  # noqa: F401
from .mamba2 import Mamba2Config, Mamba2DecodeEngine  # noqa: F401
from .transformer import TransformerConfig, TransformerDecodeEngine  # noqa: F401
