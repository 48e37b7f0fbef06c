"""Block-sparse attention.  Parity: reference flashinfer/sparse.py:69-1182.

A BSR mask with block size (R, C) is exactly a paged-KV problem: every row block is a "request" with R queries whose KV
"pages" are its non-zero column blocks (page_size = C), so the tcgen05 paged prefill / decode kernels run it unchanged.
"""
from __future__ import annotations

from typing import Optional, Union

import torch

from . import reference
from .decode import BatchDecodeWithPagedKVCacheWrapper
from .prefill import BatchPrefillWithPagedKVCacheWrapper
from .utils import legacy_forward_replan, remember_plan


def convert_bsr_mask_layout(mask: torch.Tensor, indptr: torch.Tensor) -> torch.Tensor:
    """BSR block masks ``[nnz, R, C]`` -> the flattened per-block-row layout of the reference (sparse.py:44): for every block row
    the blocks are laid side by side, i.e. ``[R, n_blocks * C]`` flattened."""
    nnz, R, C = mask.shape
    out = torch.empty(nnz * R * C, dtype=mask.dtype, device=mask.device)
    ip = indptr.tolist()
    for i in range(len(ip) - 1):
        out[ip[i] * R * C: ip[i + 1] * R * C] = mask[ip[i]: ip[i + 1]].transpose(0, 1).reshape(-1)
    return out


class BlockSparseAttentionWrapper:
    """Attention with a fixed-size block-sparse (BSR) mask ``[M/R, N/C]``."""

    def __init__(self, float_workspace_buffer: torch.Tensor, backend: str = "auto") -> None:
        self._ws = float_workspace_buffer
        self.device = float_workspace_buffer.device
        self._prefill = BatchPrefillWithPagedKVCacheWrapper(float_workspace_buffer, "NHD")
        self._decode = BatchDecodeWithPagedKVCacheWrapper(float_workspace_buffer, "NHD")

    def reset_workspace_buffer(self, float_workspace_buffer, int_workspace_buffer, **kw) -> None:
        self._prefill.reset_workspace_buffer(float_workspace_buffer, int_workspace_buffer)

    def plan(self, indptr: torch.Tensor, indices: torch.Tensor, M: int, N: int, R: int, C: int, num_qo_heads: int,
             num_kv_heads: int, head_dim: int, mask: Optional[torch.Tensor] = None,
             packed_mask: Optional[torch.Tensor] = None, causal: bool = False, pos_encoding_mode: str = "NONE",
             use_fp16_qk_reduction: bool = False, logits_soft_cap: Optional[float] = None,
             sm_scale: Optional[float] = None, rope_scale=None, rope_theta=None, q_data_type="float16",
             kv_data_type=None, o_data_type="float16", non_blocking: bool = True) -> None:
        remember_plan(self, locals())
        if M % R or N % C:
            raise ValueError("M must be a multiple of R and N of C")
        # element-level mask inside the non-zero blocks: ``mask [nnz, R, C]`` (or already in the flattened per-block-row layout of
        # convert_bsr_mask_layout) / its little-endian bit-packed form -> the custom-mask stream of the tcgen05 prefill kernel
        flat_mask = None
        if mask is not None:
            flat_mask = (convert_bsr_mask_layout(mask.bool(), indptr) if mask.dim() == 3 else mask.flatten().bool())
        elif packed_mask is not None:
            from .prefill import _unpack_segmented

            ip = indptr.to("cpu", torch.int64)
            flat_mask = _unpack_segmented(packed_mask, ((ip[1:] - ip[:-1]) * (R * C)).tolist())  # one byte-aligned segment per block row
        self._M, self._N, self._R, self._C = M, N, R, C
        self._hq, self._hkv, self._d = num_qo_heads, num_kv_heads, head_dim
        self._bsr_indptr, self._bsr_indices = indptr.to("cpu", torch.int32), indices.to("cpu", torch.int32)  # (fi_trace: run() inputs)
        self._sm_scale = float(sm_scale) if sm_scale is not None else 1.0 / (head_dim ** 0.5)
        mb = M // R
        # page granularity of the paged kernels: a page must tile the 128-token KV tile (power of two <= 128) or be a multiple of
        # 128; any other block width C is cut into C / g pages of g = largest power of two dividing C
        if (C <= 128 and C & (C - 1) == 0) or C % 128 == 0:
            g = C
        else:
            g = C & -C
            g = min(g, 128)
        ppb = C // g
        self._g = g
        indptr_i = indptr.to("cpu", torch.int64)
        indices_i = indices.to("cpu", torch.int64)
        if ppb > 1:
            indices_i = (indices_i[:, None] * ppb + torch.arange(ppb)[None, :]).reshape(-1)
            indptr_i = indptr_i * ppb
        self._indptr_host = indptr_i.to(torch.int32)
        self._indices = indices_i.to(torch.int32)
        last = torch.full((mb,), g, dtype=torch.int32)
        # one query row per block row and no intra-block mask: the decode kernel (swap-AB, bandwidth-bound); otherwise prefill
        self._use_decode = R == 1 and (num_qo_heads // num_kv_heads) <= 32 and not causal and flat_mask is None
        if self._use_decode:
            self._decode.plan(self._indptr_host, self._indices, last, num_qo_heads, num_kv_heads, head_dim, g, q_data_type=q_data_type,
                              logits_soft_cap=logits_soft_cap, sm_scale=sm_scale)
        else:
            qo = torch.arange(0, (mb + 1) * R, R, dtype=torch.int32)
            self._prefill.plan(qo, self._indptr_host, self._indices, last, num_qo_heads, num_kv_heads, head_dim, g, causal=causal,
                               logits_soft_cap=logits_soft_cap, sm_scale=sm_scale, q_data_type=q_data_type, custom_mask=flat_mask,
                               pos_encoding_mode=pos_encoding_mode)

    begin_forward = plan

    def run(self, q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, scale_q=None, scale_k=None, scale_v=None,
            out: Optional[torch.Tensor] = None, lse: Optional[torch.Tensor] = None, return_lse: bool = False,
            enable_pdl=None):
        g = self._g
        kc = k.reshape(-1, g, self._hkv, self._d)
        vc = v.reshape(-1, g, self._hkv, self._d)
        scales = {n: v_ for n, v_ in (("q_scale", scale_q), ("k_scale", scale_k), ("v_scale", scale_v)) if v_ is not None}
        if self._use_decode:
            res = self._decode.run(q, (kc, vc), out=out, lse=lse, return_lse=return_lse, **scales)
        else:
            res = self._prefill.run(q, (kc, vc), out=out, lse=lse, return_lse=return_lse, **scales)
        return res

    def forward(self, q, k, v, scale_q=None, scale_k=None, scale_v=None, pos_encoding_mode="NONE", use_fp16_qk_reduction=False,
                logits_soft_cap=None, sm_scale=None, rope_scale=None, rope_theta=None):
        """Deprecated (use :meth:`run`): the attention parameters given here replace the planned ones, defaults included."""
        legacy_forward_replan(self, pos_encoding_mode=pos_encoding_mode, logits_soft_cap=logits_soft_cap, sm_scale=sm_scale, rope_scale=rope_scale,
                              rope_theta=rope_theta)
        return self.run(q, k, v, scale_q, scale_k, scale_v)

    def end_forward(self) -> None:
        pass


class VariableBlockSparseAttentionWrapper:
    """Block-sparse attention with per-block variable sizes (reference :658).  Column blocks are expanded to a
    token-granular page list (page_size 1), which the paged kernels gather through TMA."""

    def __init__(self, float_workspace_buffer: torch.Tensor, backend: str = "auto") -> None:
        self.device = float_workspace_buffer.device
        self._prefill = BatchPrefillWithPagedKVCacheWrapper(float_workspace_buffer, "NHD")

    def reset_workspace_buffer(self, float_workspace_buffer, int_workspace_buffer=None, **kw) -> None:
        self._float_workspace_buffer = float_workspace_buffer

    def plan(self, block_mask_map: torch.Tensor, block_row_sz: torch.Tensor, block_col_sz: torch.Tensor,
             num_qo_heads: int, num_kv_heads: int, head_dim: int, causal: bool = False, pos_encoding_mode: str = "NONE",
             use_fp16_qk_reduction: bool = False, logits_soft_cap: Optional[float] = None,
             sm_scale: Optional[float] = None, rope_scale=None, rope_theta=None, non_blocking: bool = True,
             q_data_type="float16", kv_data_type=None) -> None:
        """``block_mask_map [num_kv_heads, MB, NB]`` bool, ``block_row_sz [num_kv_heads, MB]``,
        ``block_col_sz [num_kv_heads, NB]``.  Heads are folded into the batch dimension."""
        remember_plan(self, locals())
        hkv, mb, nb = block_mask_map.shape
        self._hq, self._hkv, self._d = num_qo_heads, num_kv_heads, head_dim
        bm = block_mask_map.cpu().bool()
        rs, cs = block_row_sz.cpu().long(), block_col_sz.cpu().long()
        self._seq_q = int(rs[0].sum())
        self._seq_kv = int(cs[0].sum())
        self._block_mask_map, self._block_row_sz, self._block_col_sz = bm, rs, cs   # host copies (read by the fi_trace template)
        self._sm_scale = sm_scale if sm_scale is not None else head_dim ** -0.5
        qo, kvp, idx = [0], [0], []
        for h in range(hkv):
            col_start = torch.cat([torch.zeros(1, dtype=torch.long), cs[h].cumsum(0)])
            for i in range(mb):
                toks = [torch.arange(int(col_start[j]), int(col_start[j + 1])) + h * self._seq_kv
                        for j in range(nb) if bm[h, i, j]]
                t = torch.cat(toks) if toks else torch.empty(0, dtype=torch.long)
                idx.append(t)
                kvp.append(kvp[-1] + t.numel())
                qo.append(qo[-1] + int(rs[h, i]))
        kv_indices = torch.cat(idx).int() if idx else torch.empty(0, dtype=torch.int32)
        n_req = len(qo) - 1
        group = num_qo_heads // num_kv_heads
        self._prefill.plan(torch.tensor(qo, dtype=torch.int32), torch.tensor(kvp, dtype=torch.int32), kv_indices,
                           torch.ones(n_req, dtype=torch.int32), group, 1, head_dim, 1, causal=causal,  # causal over the gathered keys, like the reference
                           logits_soft_cap=logits_soft_cap, sm_scale=sm_scale, q_data_type=q_data_type)

    def run(self, q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, out=None, lse=None, return_lse: bool = False,
            enable_pdl=None):
        """q ``[num_qo_heads, seq_q, D]``, k/v ``[num_kv_heads, seq_kv, D]`` (HND like the reference)."""
        hq, sq, d = q.shape
        hkv = k.shape[0]
        g = hq // hkv
        # rows ordered (kv head, token) with the `group` q heads of that kv head as the head dimension
        qf = q.view(hkv, g, sq, d).permute(0, 2, 1, 3).reshape(hkv * sq, g, d).contiguous()
        kf = k.reshape(hkv * k.shape[1], 1, 1, d)
        vf = v.reshape(hkv * v.shape[1], 1, 1, d)
        res = self._prefill.run(qf, (kf, vf), return_lse=return_lse)
        o = res[0] if return_lse else res
        o = o.view(hkv, sq, g, d).permute(0, 2, 1, 3).reshape(hq, sq, d)
        if out is not None:
            out.copy_(o)
            o = out
        if return_lse:
            l = res[1].view(hkv, sq, g).permute(0, 2, 1).reshape(hq, sq)
            if lse is not None:
                lse.copy_(l)
                l = lse
            return o, l
        return o

    def forward(self, q, k, v, pos_encoding_mode="NONE", use_fp16_qk_reduction=False, logits_soft_cap=None, sm_scale=None, rope_scale=None,
                rope_theta=None):
        """Deprecated (use :meth:`run`): the attention parameters given here replace the planned ones, defaults included."""
        legacy_forward_replan(self, pos_encoding_mode=pos_encoding_mode, logits_soft_cap=logits_soft_cap, sm_scale=sm_scale, rope_scale=rope_scale,
                              rope_theta=rope_theta)
        return self.run(q, k, v)
