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
        if M % R or N % C:
            raise ValueError("M must be a multiple of R and N of C")
        if mask is not None or packed_mask is not None:
            raise NotImplementedError("element-level masks inside blocks")
        self._M, self._N, self._R, self._C = M, N, R, C
        self._hq, self._hkv, self._d = num_qo_heads, num_kv_heads, head_dim
        mb = M // R
        self._indptr_host = indptr.to("cpu", torch.int32)
        self._indices = indices.to(torch.int32)
        last = torch.full((mb,), C, dtype=torch.int32)
        self._use_decode = (R * (num_qo_heads // num_kv_heads) <= 32) and not causal
        if self._use_decode:
            qo = torch.arange(0, (mb + 1) * R, R, dtype=torch.int32) if R > 1 else None
            self._decode.plan(indptr, indices, last, num_qo_heads, num_kv_heads, head_dim, C, q_data_type=q_data_type,
                              logits_soft_cap=logits_soft_cap, sm_scale=sm_scale, qo_indptr=qo)
            self._decode_noncausal = R > 1
        else:
            if C & (C - 1) and C % 128:
                raise NotImplementedError("block-sparse prefill needs C to be a power of two or a multiple of 128")
            qo = torch.arange(0, (mb + 1) * R, R, dtype=torch.int32)
            self._prefill.plan(qo, indptr, indices, last, num_qo_heads, num_kv_heads, head_dim, C, causal=causal,
                               logits_soft_cap=logits_soft_cap, sm_scale=sm_scale, q_data_type=q_data_type)

    begin_forward = plan

    def run(self, q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, scale_q=None, scale_k=None, scale_v=None,
            out: Optional[torch.Tensor] = None, lse: Optional[torch.Tensor] = None, return_lse: bool = False,
            enable_pdl=None):
        C = self._C
        kc = k.reshape(-1, C, self._hkv, self._d)
        vc = v.reshape(-1, C, self._hkv, self._d)
        if self._use_decode:
            if self._decode_noncausal:
                # decode kernel applies causal masking among the R new tokens; block-sparse rows see whole blocks,
                # so run row by row semantics through the prefill wrapper instead when R > 1
                raise NotImplementedError("R > 1 with small groups: plan with causal=False uses the prefill kernel")
            res = self._decode.run(q, (kc, vc), out=out, lse=lse, return_lse=return_lse)
        else:
            res = self._prefill.run(q, (kc, vc), out=out, lse=lse, return_lse=return_lse)
        return res

    forward = run

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
        if causal:
            raise NotImplementedError("causal variable block-sparse")
        hkv, mb, nb = block_mask_map.shape
        self._hq, self._hkv, self._d = num_qo_heads, num_kv_heads, head_dim
        bm = block_mask_map.cpu().bool()
        rs, cs = block_row_sz.cpu().long(), block_col_sz.cpu().long()
        self._seq_q = int(rs[0].sum())
        self._seq_kv = int(cs[0].sum())
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
                           torch.ones(n_req, dtype=torch.int32), group, 1, head_dim, 1, causal=False,
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
            return o, l
        return o

    forward = run
