"""XQA entry points (reference flashinfer/xqa.py:155,447).  On B200 they are served by the tcgen05 decode / MLA kernels."""
from __future__ import annotations

from typing import Optional

import torch

from .decode import trtllm_batch_decode_with_kv_cache
from .mla import trtllm_batch_decode_with_kv_cache_mla


def xqa(q: torch.Tensor, k_cache: torch.Tensor, v_cache: torch.Tensor, page_table: torch.Tensor, seq_lens: torch.Tensor,
        output: torch.Tensor, workspace_buffer: torch.Tensor, semaphores=None, num_kv_heads: Optional[int] = None,
        page_size: Optional[int] = None, sinks=None, q_scale: float = 1.0, kv_scale=None, sliding_win_size: int = 0,
        kv_layout: str = "NHD", sm_count=None, enable_pdl=None, rcp_out_scale: float = 1.0, q_seq_len: int = 1, mask=None,
        k_sf_cache=None, v_sf_cache=None):
    """q ``[B, beam(=1), Hq, D]`` -> output of the same shape; paged K/V + page table (reference semantics)."""
    from .utils import reject_unsupported

    # speculative-decoding tree masks and NVFP4 KV scale factors of the reference's XQA kernel are not implemented here
    reject_unsupported("xqa", mask=mask, k_sf_cache=k_sf_cache, v_sf_cache=v_sf_cache)
    b = q.shape[0]
    d = q.shape[-1]
    qq = q.reshape(b * q_seq_len, -1, d)
    sm_scale = q_scale * (float(kv_scale) if kv_scale is not None else 1.0) / (d ** 0.5)
    res = trtllm_batch_decode_with_kv_cache(qq, (k_cache, v_cache), workspace_buffer, page_table, seq_lens.reshape(-1),
                                            int(seq_lens.max()), bmm1_scale=sm_scale,
                                            window_left=sliding_win_size - 1 if sliding_win_size > 0 else -1,
                                            kv_layout=kv_layout, sinks=sinks, q_len_per_req=q_seq_len,
                                            bmm2_scale=(float(kv_scale) if kv_scale is not None else 1.0) * float(rcp_out_scale))
    output.copy_(res.reshape(output.shape))
    return output


def xqa_mla(q, k_cache, v_cache, page_table, seq_lens, output, workspace_buffer, semaphores=None, page_size=None,
            q_scale: float = 1.0, kv_scale=None, sm_count=None, enable_pdl=None):
    """MLA variant: q ``[B, 1, H, 576]``, latent cache ``[pages, page, 576]``."""
    b = q.shape[0]
    sm = q_scale * (float(kv_scale) if kv_scale is not None else 1.0) / ((128 + 64) ** 0.5)
    res = trtllm_batch_decode_with_kv_cache_mla(q.reshape(b, 1, q.shape[-2], 576), k_cache, workspace_buffer, 128, 512, 64,
                                                page_table, seq_lens.reshape(-1), int(seq_lens.max()), bmm1_scale=sm)
    output.copy_(res.reshape(output.shape))
    return output


def get_xqa_module(*args, **kwargs):
    """The native module behind this file's ops (reference xqa.py get_xqa_module: the JIT module accessor)."""
    from . import jit

    return jit.load("decode_sm100")


def get_xqa_module_mla(*args, **kwargs):
    """The native module behind this file's ops (reference xqa.py get_xqa_module_mla: the JIT module accessor)."""
    from . import jit

    return jit.load("mla_sm100")
