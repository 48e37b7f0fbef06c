"""Import path of the reference's CuTe-DSL kernels (flashinfer/cute_dsl/__init__.py).  The reference writes these ops in the
CUTLASS Python DSL; here the same entry points are served by the hand-written sm_100a kernels of this library (csrc/), so the
package is a set of names, not a second implementation:

* ``rmsnorm_fp4quant`` / ``add_rmsnorm_fp4quant``: one fused kernel (csrc/elementwise/quantization.cu);
* ``grouped_gemm_nt_masked``, ``create_scale_factor_tensor``: the block-scaled grouped tcgen05 GEMM family (gemm/);
* ``*_cute`` norm functions: the norm kernels (csrc/elementwise/norm.cu);
* ``attention``: the variant objects and the CuTe-DSL wrapper classes on top of the tcgen05 prefill / MLA kernels."""
from .utils import (  # noqa: F401
    ceil_div,
    convert_sf_from_mma_layout,
    convert_sf_to_mma_layout,
    get_mma_sf_shape,
    get_num_sm,
    is_cute_dsl_available,
)
from ..norm import add_rmsnorm_fp4quant, rmsnorm_fp4quant  # noqa: F401
from ..norm import (  # noqa: F401
    fused_add_rmsnorm_cute,
    fused_add_rmsnorm_quant_cute,
    layernorm_cute,
    qk_rmsnorm_cute,
    rmsnorm_cute,
    rmsnorm_quant_cute,
)
from ..gemm.grouped import grouped_gemm_nt_masked  # noqa: F401
from . import attention  # noqa: F401


def get_sm_version(device=None) -> int:
    """Compute capability as an integer (100 on B200); the reference's fused norm + FP4 kernels branch on it."""
    import torch

    if not torch.cuda.is_available():
        return 100
    major, minor = torch.cuda.get_device_capability(device)
    return major * 10 + minor


def create_scale_factor_tensor(l: int, mn: int, k: int, sf_vec_size: int = 16, dtype=None, device="cuda"):
    """Random e4m3 scale factors for ``l`` operands of shape ``[mn, k]``: returns ``(reference fp32 [mn, k, l] with every scale
    repeated over its sf_vec_size columns, swizzled byte buffer viewed in the 6-d MMA layout)`` - the pair the reference's GEMM
    tests feed to its kernel and to its oracle."""
    import torch

    from ..quantization.fp4 import block_scale_interleave

    kc = ceil_div(k, sf_vec_size)
    sf = (torch.rand(l, mn, kc, device=device) * 1.5 + 0.25).to(torch.float8_e4m3fn)
    ref = sf.float().repeat_interleave(sf_vec_size, -1)[..., :k].permute(1, 2, 0).contiguous()
    swz = block_scale_interleave(sf.view(torch.uint8)).reshape(-1)
    return ref, convert_sf_to_mma_layout(swz, mn, k, l, sf_vec_size)


__all__ = ["is_cute_dsl_available", "get_num_sm", "convert_sf_to_mma_layout", "convert_sf_from_mma_layout", "get_mma_sf_shape",
           "grouped_gemm_nt_masked", "create_scale_factor_tensor", "rmsnorm_fp4quant", "add_rmsnorm_fp4quant", "get_sm_version",
           "rmsnorm_cute", "qk_rmsnorm_cute", "rmsnorm_quant_cute", "fused_add_rmsnorm_cute", "fused_add_rmsnorm_quant_cute", "layernorm_cute", "attention"]
