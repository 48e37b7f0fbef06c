"""Rotary position embeddings.  Parity: reference flashinfer/rope.py:433-1691.

All variants funnel into one CUDA kernel (csrc/elementwise/rope.cu); CPU tensors use the fp32 oracle.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import jit, reference
from .utils import dtype_code, stream_ptr


def _launch(q, k, q_out, k_out, *, pos_ids=None, indptr=None, offsets=None, cos_sin_cache=None, rotary_dim=None,
            interleave=False, rope_scale=1.0, rope_theta=1e4, llama31=None, q_out_scale=1.0, k_out_scale=1.0):
    nnz, hq, d = q.shape
    hk = k.shape[1] if k is not None else 0
    rd = rotary_dim if rotary_dim is not None else (cos_sin_cache.shape[-1] if cos_sin_cache is not None else d)
    if not q.is_cuda:
        if pos_ids is None:
            lens = (indptr[1:] - indptr[:-1]).long()
            b = torch.repeat_interleave(torch.arange(lens.numel()), lens)
            pos_ids = torch.arange(nnz) - indptr[:-1].long()[b] + offsets.long()[b]
        if cos_sin_cache is not None:
            qo = reference.apply_rope_cos_sin_cache_ref(q, pos_ids, cos_sin_cache, interleave)
            ko = reference.apply_rope_cos_sin_cache_ref(k, pos_ids, cos_sin_cache, interleave) if k is not None else None
        else:
            qo = reference.apply_rope_ref(q, pos_ids, rd, interleave, rope_scale, rope_theta, llama31)
            ko = reference.apply_rope_ref(k, pos_ids, rd, interleave, rope_scale, rope_theta, llama31) if k is not None else None
        q_out.copy_((qo.float() * q_out_scale).to(q_out.dtype))
        if k is not None:
            k_out.copy_((ko.float() * k_out_scale).to(k_out.dtype))
        return
    for t in (q, k, q_out, k_out):
        if t is not None and t.stride(-1) != 1:
            raise ValueError("rope: last dim must be contiguous")
    pos_is_i64 = 0
    if pos_ids is not None:
        if pos_ids.dtype == torch.int64:
            pos_is_i64 = 1
        elif pos_ids.dtype != torch.int32:
            pos_ids = pos_ids.int()
    if cos_sin_cache is not None and cos_sin_cache.dtype != torch.float32:
        raise ValueError("cos_sin_cache must be float32")
    l31 = llama31 or (1.0, 4.0, 8192.0)
    batch = indptr.numel() - 1 if indptr is not None else 0
    jit.load("rope").call(
        "rope_run", q, k, q_out, k_out, pos_ids, pos_is_i64, indptr, offsets, cos_sin_cache, nnz, batch, hq, hk, d, rd,
        1 if interleave else 0, q.stride(0), q.stride(1), k.stride(0) if k is not None else 0,
        k.stride(1) if k is not None else 0, q_out.stride(0), q_out.stride(1),
        k_out.stride(0) if k_out is not None else 0, k_out.stride(1) if k_out is not None else 0,
        float(rope_scale), float(rope_theta), 1 if llama31 is not None else 0, float(l31[0]), float(l31[1]),
        float(l31[2]), float(q_out_scale), float(k_out_scale), dtype_code(q.dtype), dtype_code(q_out.dtype), 1,
        stream_ptr(q),
    )


# ------------------------------------------------------------------ indptr/offsets API
def apply_rope_inplace(q, k, indptr, offsets, rotary_dim=None, interleave=False, rope_scale=1, rope_theta=1e4) -> None:
    _launch(q, k, q, k, indptr=indptr.int(), offsets=offsets.int(), rotary_dim=rotary_dim, interleave=interleave,
            rope_scale=rope_scale, rope_theta=rope_theta)


def apply_rope(q, k, indptr, offsets, rotary_dim=None, interleave=False, rope_scale=1, rope_theta=1e4):
    qo, ko = torch.empty_like(q), torch.empty_like(k)
    _launch(q, k, qo, ko, indptr=indptr.int(), offsets=offsets.int(), rotary_dim=rotary_dim, interleave=interleave,
            rope_scale=rope_scale, rope_theta=rope_theta)
    return qo, ko


def apply_llama31_rope_inplace(q, k, indptr, offsets, rotary_dim=None, interleave=False, rope_scale=8, rope_theta=5e5,
                               low_freq_factor=1, high_freq_factor=4, old_context_len=8192) -> None:
    _launch(q, k, q, k, indptr=indptr.int(), offsets=offsets.int(), rotary_dim=rotary_dim, interleave=interleave,
            rope_scale=rope_scale, rope_theta=rope_theta, llama31=(low_freq_factor, high_freq_factor, old_context_len))


def apply_llama31_rope(q, k, indptr, offsets, rotary_dim=None, interleave=False, rope_scale=8, rope_theta=5e5,
                       low_freq_factor=1, high_freq_factor=4, old_context_len=8192):
    qo, ko = torch.empty_like(q), torch.empty_like(k)
    _launch(q, k, qo, ko, indptr=indptr.int(), offsets=offsets.int(), rotary_dim=rotary_dim, interleave=interleave,
            rope_scale=rope_scale, rope_theta=rope_theta, llama31=(low_freq_factor, high_freq_factor, old_context_len))
    return qo, ko


# ------------------------------------------------------------------ pos_ids API
def apply_rope_pos_ids_inplace(q, k, pos_ids, rotary_dim=None, interleave=False, rope_scale=1, rope_theta=1e4) -> None:
    _launch(q, k, q, k, pos_ids=pos_ids, rotary_dim=rotary_dim, interleave=interleave, rope_scale=rope_scale,
            rope_theta=rope_theta)


def apply_rope_pos_ids(q, k, pos_ids, rotary_dim=None, interleave=False, rope_scale=1, rope_theta=1e4):
    qo, ko = torch.empty_like(q), torch.empty_like(k)
    _launch(q, k, qo, ko, pos_ids=pos_ids, rotary_dim=rotary_dim, interleave=interleave, rope_scale=rope_scale,
            rope_theta=rope_theta)
    return qo, ko


def apply_llama31_rope_pos_ids_inplace(q, k, pos_ids, rotary_dim=None, interleave=False, rope_scale=8, rope_theta=5e5,
                                       low_freq_factor=1, high_freq_factor=4, old_context_len=8192) -> None:
    _launch(q, k, q, k, pos_ids=pos_ids, rotary_dim=rotary_dim, interleave=interleave, rope_scale=rope_scale,
            rope_theta=rope_theta, llama31=(low_freq_factor, high_freq_factor, old_context_len))


def apply_llama31_rope_pos_ids(q, k, pos_ids, rotary_dim=None, interleave=False, rope_scale=8, rope_theta=5e5,
                               low_freq_factor=1, high_freq_factor=4, old_context_len=8192):
    qo, ko = torch.empty_like(q), torch.empty_like(k)
    _launch(q, k, qo, ko, pos_ids=pos_ids, rotary_dim=rotary_dim, interleave=interleave, rope_scale=rope_scale,
            rope_theta=rope_theta, llama31=(low_freq_factor, high_freq_factor, old_context_len))
    return qo, ko


# ------------------------------------------------------------------ cos/sin cache API (vLLM / SGL compatible)
def apply_rope_with_cos_sin_cache(positions, query, key, head_size, cos_sin_cache, is_neox=True) -> Tuple[torch.Tensor, torch.Tensor]:
    """query ``[nnz, Hq*head_size]``, key ``[nnz, Hk*head_size]``, cos_sin_cache ``[max_pos, rotary_dim]`` fp32."""
    qo, ko = torch.empty_like(query), torch.empty_like(key)
    _launch(query.view(query.shape[0], -1, head_size), key.view(key.shape[0], -1, head_size),
            qo.view(qo.shape[0], -1, head_size), ko.view(ko.shape[0], -1, head_size), pos_ids=positions,
            cos_sin_cache=cos_sin_cache, interleave=not is_neox)
    return qo, ko


def apply_rope_with_cos_sin_cache_inplace(positions, query, key, head_size, cos_sin_cache, is_neox=True) -> None:
    q3, k3 = query.view(query.shape[0], -1, head_size), key.view(key.shape[0], -1, head_size)
    _launch(q3, k3, q3, k3, pos_ids=positions, cos_sin_cache=cos_sin_cache, interleave=not is_neox)


# ------------------------------------------------------------------ RoPE + fp8 quantisation
def rope_quantize_fp8(q_rope, k_rope, q_nope, k_nope, cos_sin_cache, pos_ids, is_neox=True,
                      quantize_dtype: Optional[torch.dtype] = None, quant_scale_q: float = 1.0,
                      quant_scale_kv: float = 1.0, q_rope_out=None, k_rope_out=None, q_nope_out=None, k_nope_out=None,
                      enable_pdl: bool = False):
    """Apply RoPE to the rope slices, scale + quantise both rope and nope slices to fp8
    (reference flashinfer/rope.py:1313).  q_* are ``[nnz, H, d]``; k_* are ``[nnz, Hk, d]`` or ``[nnz, d]``."""
    qd = quantize_dtype or torch.float8_e4m3fn
    k2d = k_rope.ndim == 2
    kr = k_rope.unsqueeze(1) if k2d else k_rope
    kn = k_nope.unsqueeze(1) if (k_nope is not None and k_nope.ndim == 2) else k_nope
    q_rope_out = q_rope_out if q_rope_out is not None else torch.empty_like(q_rope, dtype=qd)
    k_rope_out = k_rope_out if k_rope_out is not None else torch.empty_like(k_rope, dtype=qd)
    kro = k_rope_out.unsqueeze(1) if k2d else k_rope_out
    _launch(q_rope, kr, q_rope_out, kro, pos_ids=pos_ids, cos_sin_cache=cos_sin_cache, interleave=not is_neox,
            q_out_scale=quant_scale_q, k_out_scale=quant_scale_kv)
    lim = 448.0 if qd == torch.float8_e4m3fn else 57344.0
    if q_nope is not None and q_nope.is_cuda and q_nope.shape[-1] % 8 == 0 and (k_nope is None or kn.shape[-1] == q_nope.shape[-1]):
        # the no-rope slices ride the same native kernel with rotary_dim = 0: a saturating scale + fp8 cast, no rotation
        q_nope_out = q_nope_out if q_nope_out is not None else torch.empty_like(q_nope, dtype=qd)
        kno = None
        if k_nope is not None:
            k_nope_out = k_nope_out if k_nope_out is not None else torch.empty_like(k_nope, dtype=qd)
            kno = k_nope_out.unsqueeze(1) if k_nope.ndim == 2 else k_nope_out
        _launch(q_nope, kn if k_nope is not None else None, q_nope_out, kno, pos_ids=pos_ids, rotary_dim=0, interleave=False,
                q_out_scale=quant_scale_q, k_out_scale=quant_scale_kv)
        return q_rope_out, k_rope_out, q_nope_out, k_nope_out
    if q_nope is not None:
        qn = (q_nope.float() * quant_scale_q).clamp(-lim, lim).to(qd)
        q_nope_out = qn if q_nope_out is None else q_nope_out.copy_(qn)
    if k_nope is not None:
        knq = (kn.float() * quant_scale_kv).clamp(-lim, lim).to(qd)
        knq = knq.squeeze(1) if k_nope.ndim == 2 else knq
        k_nope_out = knq if k_nope_out is None else k_nope_out.copy_(knq)
    return q_rope_out, k_rope_out, q_nope_out, k_nope_out


def mla_rope_quantize_fp8(q_rope, k_rope, q_nope, k_nope, cos_sin_cache, pos_ids, is_neox=True, quantize_dtype=None,
                          quant_scale_q=1.0, quant_scale_kv=1.0, q_rope_out=None, k_rope_out=None, q_nope_out=None,
                          k_nope_out=None, enable_pdl=False):
    return rope_quantize_fp8(q_rope, k_rope, q_nope, k_nope, cos_sin_cache, pos_ids, is_neox, quantize_dtype,
                             quant_scale_q, quant_scale_kv, q_rope_out, k_rope_out, q_nope_out, k_nope_out, enable_pdl)


# ------------------------------------------------------------------ fused RoPE + paged-KV append (one kernel)
def apply_rope_append_paged_kv_cache(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, pos_ids: torch.Tensor,
                                     batch_indices: torch.Tensor, paged_kv_cache, kv_indices: torch.Tensor,
                                     kv_indptr: torch.Tensor, kv_layout: str = "NHD", rotary_dim=None, interleave: bool = False,
                                     rope_scale: float = 1.0, rope_theta: float = 1e4, llama31=None, cos_sin_cache=None,
                                     q_out: Optional[torch.Tensor] = None, q_scale: float = 1.0, kv_scale: float = 1.0) -> torch.Tensor:
    """``q`` gets RoPE in place (or into ``q_out``); ``k`` gets RoPE and is written, with ``v``, straight into the cache pages
    of its request (``pos_ids`` is the token position = page slot).  One launch replaces apply_rope + append_paged_kv_cache;
    with fp8 caches the same kernel quantises (``*_scale`` are multipliers).  Reference: rope_quantize_fp8_append_paged_kv_cache."""
    from .utils import paged_kv_strides, unpack_paged_kv_cache

    k_cache, v_cache = unpack_paged_kv_cache(paged_kv_cache, kv_layout)
    q_out = q if q_out is None else q_out
    if not q.is_cuda:
        from . import page as _page

        kr = torch.empty_like(k)
        _launch(q, k, q_out, kr, pos_ids=pos_ids, cos_sin_cache=cos_sin_cache, rotary_dim=rotary_dim, interleave=interleave,
                rope_scale=rope_scale, rope_theta=rope_theta, llama31=llama31, q_out_scale=q_scale, k_out_scale=kv_scale)
        _page.append_paged_kv_cache(kr.to(k_cache.dtype), (v.float() * kv_scale).to(v_cache.dtype), batch_indices, pos_ids,
                                    (k_cache, v_cache), kv_indices, kv_indptr, None, kv_layout)
        return q_out
    sp, sn, sh, page_size, _, _ = paged_kv_strides(k_cache, kv_layout)
    if paged_kv_strides(v_cache, kv_layout)[:3] != (sp, sn, sh) or v_cache.dtype != k_cache.dtype:
        raise ValueError("k_cache and v_cache must share strides and dtype")
    if k_cache.dtype != q_out.dtype:
        raise ValueError("q_out dtype must match the cache dtype (one output type per launch)")
    jit.load("rope").call("rope_set_append", v, k_cache, v_cache, batch_indices.int(), kv_indices.int(), kv_indptr.int(), v.stride(0),
                          v.stride(1), sp, sn, sh, page_size, float(kv_scale))
    _launch(q, k, q_out, k, pos_ids=pos_ids.int() if pos_ids.dtype != torch.int32 else pos_ids, cos_sin_cache=cos_sin_cache,
            rotary_dim=rotary_dim, interleave=interleave, rope_scale=rope_scale, rope_theta=rope_theta, llama31=llama31,
            q_out_scale=q_scale, k_out_scale=kv_scale)
    return q_out


def rope_quantize_fp8_append_paged_kv_cache(q_rope, k_rope, q_nope, k_nope, v, cos_sin_cache, pos_ids, paged_kv_cache, kv_indices,
                                            kv_indptr, batch_indices, positions, is_neox: bool = True, quantize_dtype=None,
                                            quant_scale_q: float = 1.0, quant_scale_kv: float = 1.0, page_size: int = 16,
                                            kv_layout: str = "NHD", q_rope_out=None, q_nope_out=None, enable_pdl: bool = False):
    """GQA / MHA flavour of the reference op (flashinfer/rope.py:1453): RoPE on the rotary slices, fp8 quantisation of Q, K, V
    and the paged append of K / V in ONE kernel.  Returns ``(q_rope_out, q_nope_out)`` in fp8.  (MLA caches: use
    ``mla_rope_quantize_fp8`` + ``append_paged_mla_kv_cache``.)"""
    qdt = quantize_dtype or torch.float8_e4m3fn
    if v is None:
        # MLA layout (reference rope.py:1600-1691): k_rope [T, 64] is the shared rope key (kpe), k_nope [T, 512] the compressed
        # latent (ckv); paged_kv_cache = (ckv_cache [pages, page, 512], kpe_cache [pages, page, 64]).  Two native kernels: RoPE +
        # fp8 quantisation of q / k (rope_quantize_fp8), then the latent-cache append.
        from . import page as _page

        ckv_cache, kpe_cache = paged_kv_cache
        qr, kr, qn, kn = rope_quantize_fp8(q_rope, k_rope, q_nope, k_nope, cos_sin_cache, pos_ids, is_neox, qdt, quant_scale_q,
                                           quant_scale_kv, q_rope_out, None, q_nope_out, None, enable_pdl)
        kr2 = kr.reshape(kr.shape[0], -1)
        kn2 = kn.reshape(kn.shape[0], -1)
        if ckv_cache.dtype != kn2.dtype or kpe_cache.dtype != kr2.dtype:
            raise ValueError("cache dtype must equal quantize_dtype")
        _page.append_paged_mla_kv_cache(kn2, kr2, batch_indices, positions, ckv_cache, kpe_cache, kv_indices, kv_indptr, None)
        return qr, qn
    rd = q_rope.shape[-1]
    # assemble full heads [nope | rope]?  the reference keeps rope dims first in memory for GQA: [rope | nope]
    q_full = torch.cat([q_rope, q_nope], -1) if q_nope is not None and q_nope.shape[-1] else q_rope
    k_full = torch.cat([k_rope, k_nope], -1) if k_nope is not None and k_nope.shape[-1] else k_rope
    q_out = torch.empty(q_full.shape, dtype=qdt, device=q_full.device)
    if q_full.is_cuda:
        k_cache = paged_kv_cache[0]
        if k_cache.dtype != qdt:
            raise ValueError("cache dtype must equal quantize_dtype")
        apply_rope_append_paged_kv_cache(q_full, k_full, v, pos_ids, batch_indices, paged_kv_cache, kv_indices, kv_indptr, kv_layout,
                                         rotary_dim=rd, interleave=not is_neox, cos_sin_cache=cos_sin_cache, q_out=q_out,
                                         q_scale=quant_scale_q, kv_scale=quant_scale_kv)
    else:
        apply_rope_append_paged_kv_cache(q_full, k_full, v, pos_ids, batch_indices, paged_kv_cache, kv_indices, kv_indptr, kv_layout,
                                         rotary_dim=rd, interleave=not is_neox, cos_sin_cache=cos_sin_cache, q_out=q_out,
                                         q_scale=quant_scale_q, kv_scale=quant_scale_kv)
    qr, qn = q_out[..., :rd], q_out[..., rd:]
    if q_rope_out is not None:
        q_rope_out.copy_(qr)
        qr = q_rope_out
    if q_nope_out is not None:
        q_nope_out.copy_(qn)
        qn = q_nope_out
    return qr, qn


from . import jit as _jit_acc  # noqa: E402

get_rope_module = _jit_acc.module_accessor("rope")
