"""Index of the op namespaces (one import for code that wants ``ops.norm.rmsnorm`` style access).  The ops themselves live in the
modules named like the reference's (``flashinfer_b200.norm``, ``.rope``, ``.sampling`` ...); nothing is defined here."""
from .. import activation, cascade, decode, fused_moe, gdn, gemm, mamba, mla, norm, page, pod, prefill, quantization, rope, sampling, sparse, topk  # noqa: F401

__all__ = ["activation", "cascade", "decode", "fused_moe", "gdn", "gemm", "mamba", "mla", "norm", "page", "pod", "prefill", "quantization",
           "rope", "sampling", "sparse", "topk"]
