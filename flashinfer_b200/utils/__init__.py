"""Shared enums, dtype codes, layout helpers and argument checks.

Parity: reference flashinfer/utils.py:32-51 (PosEncodingMode / MaskMode / TensorLayout),
``_unpack_paged_kv_cache`` (flashinfer/utils.py), ``get_seq_lens`` (flashinfer/page.py:203-228).
"""
from __future__ import annotations

import math
from enum import Enum
from typing import Optional, Tuple, Union

import torch


class PosEncodingMode(Enum):
    NONE = 0
    ROPE_LLAMA = 1
    ALIBI = 2


class MaskMode(Enum):
    NON_CAUSAL = 0
    CAUSAL = 1
    CUSTOM = 2
    MULTIITEMSCORING = 3


class TensorLayout(Enum):
    NHD = 0
    HND = 1


# dtype codes shared with csrc/include/fib200/common.cuh
DTYPE_CODE = {
    torch.float16: 0,
    torch.bfloat16: 1,
    torch.float32: 2,
    torch.float8_e4m3fn: 3,
    torch.float8_e5m2: 4,
    torch.uint8: 5,
    torch.int32: 6,
    torch.int64: 7,
}


def dtype_code(dt: torch.dtype) -> int:
    try:
        return DTYPE_CODE[dt]
    except KeyError as e:
        raise TypeError(f"unsupported dtype {dt}") from e


def check_kv_layout(kv_layout: str) -> None:
    if kv_layout not in ("NHD", "HND"):
        raise KeyError(f"Invalid kv_layout {kv_layout}")


def check_pos_encoding_mode(mode: str) -> None:
    if not hasattr(PosEncodingMode, mode):
        raise KeyError(f"Invalid pos_encoding_mode {mode}")


def is_cuda(*tensors) -> bool:
    return any(isinstance(t, torch.Tensor) and t.is_cuda for t in tensors)


def next_positive_power_of_2(x: int) -> int:
    if x < 1:
        return 1
    return 1 << (x - 1).bit_length()


def ceil_div(x: int, y: int) -> int:
    return (x + y - 1) // y


def round_up(x: int, y: int) -> int:
    return ceil_div(x, y) * y


def unpack_paged_kv_cache(
    paged_kv_cache: Union[torch.Tensor, Tuple[torch.Tensor, torch.Tensor]], kv_layout: str
) -> Tuple[torch.Tensor, torch.Tensor]:
    """Accept ``(k_cache, v_cache)`` or a combined 5-D ``[pages, 2, ...]`` tensor and return
    4-D views ``[pages, page, H, D]`` (NHD) / ``[pages, H, page, D]`` (HND) without copying."""
    if isinstance(paged_kv_cache, tuple):
        k, v = paged_kv_cache
    elif torch.is_tensor(paged_kv_cache):
        if paged_kv_cache.ndim != 5:
            raise ValueError(f"combined paged_kv_cache must be 5-D, got {paged_kv_cache.ndim}-D")
        k, v = paged_kv_cache[:, 0], paged_kv_cache[:, 1]
    else:
        raise TypeError("paged_kv_cache must be a tensor or a (k_cache, v_cache) tuple")
    if k.ndim == 3:  # page_size 1 shorthand [pages, H, D]
        k = k.unsqueeze(1 if kv_layout == "NHD" else 2)
        v = v.unsqueeze(1 if kv_layout == "NHD" else 2)
    return k, v


def host_i32(t: torch.Tensor) -> torch.Tensor:
    """Private int32 CPU copy of plan() metadata.  ``Tensor.to`` returns the SAME object when the input already is CPU int32,
    so a caller that edits its preallocated indptr in place would silently change planned state (and id()-keyed caches would
    never notice): always detach from the caller's storage."""
    h = t.to("cpu", torch.int32).contiguous()
    if h.data_ptr() == t.data_ptr():
        h = h.clone()
    return h


def paged_kv_strides(k_cache: torch.Tensor, kv_layout: str):
    """(stride_page, stride_n, stride_h, page_size, num_kv_heads, head_dim) in elements."""
    if kv_layout == "NHD":
        pages, page_size, h, d = k_cache.shape
        sp, sn, sh, sd = k_cache.stride()
    else:
        pages, h, page_size, d = k_cache.shape
        sp, sh, sn, sd = k_cache.stride()
    if sd != 1:
        raise ValueError("last dim of the KV cache must be contiguous")
    return sp, sn, sh, page_size, h, d


def get_seq_lens(kv_indptr: torch.Tensor, kv_last_page_len: torch.Tensor, page_size: int) -> torch.Tensor:
    """kv_len = (n_pages - 1) * page_size + last_page_len (0 pages -> 0)."""
    n_pages = kv_indptr[1:] - kv_indptr[:-1]
    return torch.clamp(n_pages - 1, min=0) * page_size + torch.where(
        n_pages > 0, kv_last_page_len, torch.zeros_like(kv_last_page_len)
    )


def default_sm_scale(head_dim: int) -> float:
    return 1.0 / math.sqrt(head_dim)


LOG2E = 1.4426950408889634


def device_sm_count(device=None) -> int:
    if torch.cuda.is_available():
        return torch.cuda.get_device_properties(device or torch.cuda.current_device()).multi_processor_count
    return 148


def is_sm100a_supported(device=None) -> bool:
    if not torch.cuda.is_available():
        return False
    major, _ = torch.cuda.get_device_capability(device)
    return major == 10


def stream_ptr(t: Optional[torch.Tensor] = None) -> int:
    dev = t.device if t is not None else None
    return torch.cuda.current_stream(dev).cuda_stream


def device_support_pdl(device=None) -> bool:
    if not torch.cuda.is_available():
        return False
    major, _ = torch.cuda.get_device_capability(device)
    return major >= 9


def tensor_stats(t: torch.Tensor) -> dict:
    """min / max / mean / NaN count / Inf count in one pass (native kernel on CUDA; used by API logging level 5)."""
    n = t.numel()
    if not t.is_cuda or t.dtype not in (torch.float16, torch.bfloat16, torch.float32):
        f = t.detach().float().flatten()
        fin = f[torch.isfinite(f)]
        return {"min": float(fin.min()) if fin.numel() else float("nan"), "max": float(fin.max()) if fin.numel() else float("nan"),
                "mean": float(fin.mean()) if fin.numel() else float("nan"), "nan": int(torch.isnan(f).sum()),
                "inf": int(torch.isinf(f).sum())}
    from .. import jit

    out = torch.tensor([float("inf"), float("-inf"), 0.0, 0.0, 0.0, float("inf")], dtype=torch.float32, device=t.device)
    x = t.contiguous()
    jit.load("runtime").call("tensor_stats", x, n, dtype_code(x.dtype), out, stream_ptr(x))
    o = out.tolist()
    finite = max(n - int(o[3]) - int(o[4]), 1)
    return {"min": o[5], "max": o[1], "mean": o[2] / finite, "nan": int(o[3]), "inf": int(o[4])}


# ====================================================================================================================
# Helper names of the reference's flashinfer/utils.py that serving frameworks import (capability probes, dtype / shape
# helpers, backend selectors).  This library has ONE backend (hand-written sm_100a kernels), so the selectors answer that.
# ====================================================================================================================
import functools as _functools
import math as _math
from typing import Callable, Iterable, Sequence, Tuple, Union


class GPUArchitectureError(RuntimeError):
    """Raised when an op is asked to run on a GPU architecture it does not support."""


class LibraryError(RuntimeError):
    """Raised when a required library / native module is missing."""


class BackendSupportedError(RuntimeError):
    """Raised when a requested backend cannot serve the call."""


class LogLevel(Enum):
    NOTSET = 0
    DEBUG = 10
    INFO = 20
    WARNING = 30
    ERROR = 40
    CRITICAL = 50


def set_log_level(lvl_str: str) -> None:
    import logging

    logging.getLogger("flashinfer_b200").setLevel(getattr(logging, str(lvl_str).upper(), logging.INFO))


def get_logging_module():
    import logging

    return logging.getLogger("flashinfer_b200")


def calculate_tile_tokens_dim(num_tokens: int, num_experts: int, top_k: int, max_tile_tokens_dim: int = 128) -> int:
    """Token-tile size heuristic of the reference's trtllm-gen MoE (expected tokens per expert x 1.3, next power of two, [8, max])."""
    per_expert = int(((num_tokens * top_k) // max(num_experts, 1)) * 1.3)
    return min(max(next_positive_power_of_2(max(per_expert, 1)), 8), max_tile_tokens_dim)


def is_float8(x: torch.Tensor) -> bool:
    return x.dtype in (torch.float8_e4m3fn, torch.float8_e5m2)


def get_indptr(x: torch.Tensor) -> torch.Tensor:
    x = x.to(torch.int64)
    ret = torch.zeros(x.shape[0] + 1, dtype=x.dtype, device=x.device)
    ret[1:] = x.cumsum(0)
    return ret


def get_alibi_slopes(n_heads: int, device: Optional[torch.device] = None) -> torch.Tensor:
    n = 2 ** _math.floor(_math.log2(n_heads))
    m = torch.pow(2.0 ** (-8.0 / n), torch.arange(1, 1 + n, device=device))
    if n < n_heads:
        m = torch.cat([m, torch.pow(2.0 ** (-4.0 / n), torch.arange(1, 1 + 2 * (n_heads - n), 2, device=device))])
    return m.float()


def canonicalize_torch_dtype(dtype: Union[torch.dtype, str]) -> torch.dtype:
    if isinstance(dtype, str):
        return getattr(torch, dtype)
    if isinstance(dtype, torch.dtype):
        return dtype
    raise TypeError(f"dtype must be a string or torch.dtype, got {type(dtype)}")


@_functools.lru_cache(maxsize=None)
def get_compute_capability(device: torch.device) -> Tuple[int, int]:
    device = torch.device(device)
    if device.type != "cuda":
        raise ValueError("device must be a cuda device")
    return torch.cuda.get_device_capability(device.index)


def get_device_sm_count(device=None) -> int:
    return device_sm_count(device)


def get_gpu_memory_bandwidth(device) -> float:
    """Peak DRAM bandwidth in GB/s from the device properties (memory clock x bus width x 2)."""
    device = torch.device(device)
    if device.type != "cuda":
        raise ValueError("device must be a cuda device")
    p = torch.cuda.get_device_properties(device)
    clk = getattr(p, "memory_clock_rate", 0)
    width = getattr(p, "memory_bus_width", 0)
    return clk * 1e3 * (width / 8) * 2 / 1e9 if clk and width else 8000.0  # B200 HBM3e nominal


def get_shared_bytes_per_block_optin(device=None) -> int:
    return 227 * 1024


def _cc(device) -> Tuple[int, int]:
    try:
        return get_compute_capability(torch.device(device) if device is not None else torch.device("cuda", 0))
    except Exception:  # noqa: BLE001  (no GPU in this process)
        return (0, 0)


def is_sm90a_supported(device=None) -> bool:
    return _cc(device)[0] == 9


def is_sm100f_supported(device=None) -> bool:
    return _cc(device)[0] == 10


def is_sm110a_supported(device=None) -> bool:
    return _cc(device)[0] == 11


def is_sm120a_supported(device=None) -> bool:
    return _cc(device) == (12, 0)


def is_sm120f_supported(device=None) -> bool:
    return _cc(device)[0] == 12


def is_sm121a_supported(device=None) -> bool:
    return _cc(device) == (12, 1)


def is_sm12x_supported(device=None) -> bool:
    return _cc(device)[0] == 12


def is_cvt_rs_supported(device=None) -> bool:
    return _cc(device)[0] in (10, 11)


def version_at_least(version: str, base_version: str) -> bool:
    from packaging import version as _v

    return _v.parse(version) >= _v.parse(base_version)


def has_cuda_cudart() -> bool:
    import importlib.util

    try:
        return importlib.util.find_spec("cuda.cudart") is not None
    except ModuleNotFoundError:
        return False


def get_cuda_python_version() -> str:
    try:
        import cuda

        return getattr(cuda, "__version__", "0")
    except ImportError:
        return "0"


def determine_gemm_backend(device) -> str:
    return "sm100"


def determine_attention_backend(device, pos_encoding_mode: int, use_fp16_qk_reductions: bool, use_custom_mask: bool, dtype_q,
                                dtype_kv) -> str:
    """The reference picks fa2 / fa3; here every attention call runs the sm_100a kernels."""
    return "sm100"


def determine_mla_backend(device) -> str:
    return "sm100"


def is_fa3_backend_supported(*args, **kwargs) -> bool:
    return False


def is_cutlass_backend_supported(*args, **kwargs) -> bool:
    return False


def check_shape_dtype_device(x: torch.Tensor, expected_shape: Optional[Sequence[int]], expected_dtype: Optional[torch.dtype],
                             expected_device: Optional[torch.device], name: str) -> None:
    if expected_shape and x.shape != torch.Size(expected_shape):
        raise ValueError(f"Invalid shape of {name}: expected {expected_shape}, got {x.shape}")
    if expected_dtype and x.dtype != expected_dtype:
        raise ValueError(f"Invalid dtype of {name}: expected {expected_dtype}, got {x.dtype}")
    if expected_device and x.device != expected_device:
        raise ValueError(f"Invalid device of {name}: expected {expected_device}, got {x.device}")


class FP4Tensor:
    """Packed e2m1 data (two values per uint8) + UE4M3 block scales (reference utils.py:727)."""

    def __init__(self, data: torch.Tensor, scale: torch.Tensor, scale_start_index: int = 0,
                 original_shape: Optional[Tuple[int, ...]] = None):
        if data.dtype != torch.uint8:
            raise ValueError(f"data must be uint8 tensor, got {data.dtype}")
        if original_shape is None:
            original_shape = tuple(data.shape[:-1]) + (data.shape[-1] * 2,)
        self.data, self.scale, self.scale_start_index, self.original_shape = data, scale, scale_start_index, tuple(original_shape)
        self.dtype = "nvfp4"


def get_shuffle_block_size(epilogue_tile_m: int) -> int:
    return 32 if epilogue_tile_m % 128 == 0 else 16


def get_shuffle_matrix_a_row_indices(input_tensor: torch.Tensor, epilogue_tile_m: int) -> torch.Tensor:
    from ..quantization.fp4 import _shuffle_row_indices

    assert input_tensor.dim() == 2, f"input_tensor should be a 2D tensor, not {input_tensor.dim()}"
    return _shuffle_row_indices(input_tensor.shape[0], epilogue_tile_m, input_tensor.device)


def get_shuffle_matrix_sf_a_row_indices(input_tensor: torch.Tensor, epilogue_tile_m: int, num_elts_per_sf: int = 16) -> torch.Tensor:
    return get_shuffle_matrix_a_row_indices(input_tensor, epilogue_tile_m)


def get_native_fp4_dtype():
    return getattr(torch, "float4_e2m1fn_x2", torch.uint8)


def supported_compute_capability(supported_ccs: Iterable[int]) -> Callable:
    """Decorator: annotate a function with the compute capabilities it supports (``fn.is_compute_capability_supported(cc)``)."""
    ccs = set(supported_ccs)

    def deco(fn):
        fn._supported_ccs = ccs
        fn.is_compute_capability_supported = lambda cc: cc in ccs
        return fn

    return deco


def backend_requirement(backend_checks=None, common_check=None, heuristic_func=None):
    """Decorator of the reference that validates ``backend=`` choices; the single backend here always qualifies."""
    def deco(fn):
        fn.is_backend_supported = lambda backend=None, cc=None: True
        fn.is_compute_capability_supported = lambda cc: True
        fn.has_backend = lambda backend: True
        fn.has_backend_choices = lambda: False
        return fn

    return deco


def get_default_generators(device):
    torch.cuda.init()
    return torch.cuda.default_generators[torch.device(device).index or 0]


def prepare_jit_additional_args(*args, **kwargs):
    """Reference utils.py: marshals extra tensor / scalar arguments of user-defined attention variants for its JIT templates.
    Variants here are compiled C++ functors selected by name (``attention/generic``), so there is nothing to marshal: the
    arguments are returned unchanged as ``(tensors, scalars)``."""
    tensors = [a for a in args if hasattr(a, "data_ptr")]
    scalars = [a for a in args if not hasattr(a, "data_ptr")]
    return tensors, scalars


def register_custom_op(name, fn=None, /, *, mutates_args, device_types=None, schema=None):
    """Reference name (flashinfer/utils.py:330): real ``torch.library.custom_op`` registration (see flashinfer_b200/torch_ops.py)."""
    return torch.library.custom_op(name, fn, mutates_args=mutates_args, device_types=device_types, schema=schema)


def register_fake_op(name, fn=None):
    """Reference name (flashinfer/utils.py:365): ``torch.library.register_fake``."""
    return torch.library.register_fake(name, fn)


def reject_unsupported(api: str, **arguments) -> None:
    """Raise for arguments an entry point accepts for signature parity but does not implement.  An argument counts as "set" when it is
    not None / False (pass ``name=(value, default)`` to compare against another default): accepting it silently would return a
    result computed without it."""
    bad = []
    for name, v in arguments.items():
        if isinstance(v, tuple) and len(v) == 2 and not any(isinstance(e, torch.Tensor) for e in v):     # (value, default); a pair of
            if v[0] is not None and v[0] != v[1]:                                                          # tensors is a value
                bad.append(name)
        elif v is not None and v is not False:
            bad.append(name)
    if bad:
        raise NotImplementedError(f"{api}: argument(s) {bad} are not implemented by this library")


def remember_plan(wrapper, local_vars: dict) -> None:
    """First line of a wrapper's ``plan()``: keep its arguments so the deprecated ``forward()`` entry points can plan again."""
    cls = type(wrapper)
    names = _PLAN_PARAMS.get(cls)
    if names is None:
        import inspect

        names = _PLAN_PARAMS[cls] = tuple(n for n in inspect.signature(inspect.unwrap(cls.plan)).parameters if n != "self")
    wrapper._plan_locals = {k: local_vars[k] for k in names}


_PLAN_PARAMS: dict = {}


def legacy_forward_replan(wrapper, **overrides) -> None:
    """The reference's deprecated ``forward(...)`` methods take the attention parameters (causal, pos_encoding_mode, window_left,
    logits_soft_cap, sm_scale, rope_scale, rope_theta) at call time and they REPLACE what ``begin_forward`` recorded - defaults
    included (reference prefill.py :2117, decode.py :1159, sparse.py :469).  The plans here bake those parameters in, so a value
    that differs from the planned one is a new plan with the remembered arguments."""
    args = getattr(wrapper, "_plan_locals", None)
    if args is None:
        raise RuntimeError("begin_forward() / plan() must be called before forward()")
    diff = {k: v for k, v in overrides.items() if k in args and args[k] != v}
    if diff:
        wrapper.plan(**{**args, **diff})
