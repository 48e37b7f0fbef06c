"""MoE building blocks as standalone ops (reference flashinfer/fused_moe/cute_dsl/moe_utils.py): tile-aligned expert sort, row
permute / un-permute, masked activations, output zeroing - for callers that assemble their own expert pipeline around the grouped
GEMMs (``flashinfer_b200.gemm``) instead of calling a fused entry point.

Relation to the fused pipeline (``fused_moe.core.moe_forward*``): that path runs ONE sort kernel (csrc/moe/moe.cu ``moe_sort``) and
fuses the gather into the quantisers / the activation into the re-quantiser, and never materialises these index tensors in the
layout below.  The functions here produce exactly the reference's tensors with static-shape tensor ops only (no host sync, CUDA-graph
capturable, identical on CPU - which is how they are tested); the un-permute step is the pipeline's native ``moe_finalize`` kernel.

Layout produced by :func:`moe_sort` (``tile`` = ``tile_tokens_dim``): the rows routed to local expert ``e`` occupy the permuted range
``[start_e, start_e + count_e)`` with ``start_e`` a multiple of ``tile`` (every expert is padded to whole tiles, experts with no rows
take no tile); tile ``t`` belongs to ``tile_idx_to_expert_idx[t]`` and its valid rows are ``[t * tile, tile_idx_to_mn_limit[t])``."""
from __future__ import annotations

from enum import IntEnum
from typing import Dict, Optional, Tuple

import torch

from .. import jit
from ..utils import dtype_code, stream_ptr


class MoeActivationType(IntEnum):
    Gelu = 0
    Relu = 1
    Silu = 2
    Swiglu = 3
    Geglu = 4
    Identity = 5


def get_max_num_tiles(num_tokens: int, top_k: int, num_local_experts: int, tile_size: int) -> int:
    """Most tiles any routing of ``num_tokens * top_k`` rows over ``num_local_experts`` experts can need: all experts but one hold a
    single row (one mostly-empty tile each), the remaining rows fill the last expert."""
    rows = num_tokens * top_k
    if rows <= num_local_experts:
        return rows
    return (rows - (num_local_experts - 1) + tile_size - 1) // tile_size + (num_local_experts - 1)


def get_max_num_permuted_tokens(num_tokens: int, top_k: int, num_local_experts: int, tile_size: int) -> int:
    return get_max_num_tiles(num_tokens, top_k, num_local_experts, tile_size) * tile_size


def allocate_moe_sort_buffers(num_tokens: int, num_experts: int, top_k: int, num_local_experts: Optional[int] = None,
                              tile_tokens_dim: int = 128, device: str = "cuda") -> Dict[str, torch.Tensor]:
    """Output buffers of :func:`moe_sort` (pass as ``**kwargs``): allocate once before CUDA-graph capture."""
    local = num_local_experts if num_local_experts is not None else num_experts
    tiles = get_max_num_tiles(num_tokens, top_k, local, tile_tokens_dim)
    i32 = dict(dtype=torch.int32, device=device)
    return {"out_tile_idx_to_expert_idx": torch.empty(tiles, **i32), "out_tile_idx_to_mn_limit": torch.empty(tiles, **i32),
            "out_expanded_idx_to_permuted_idx": torch.empty(num_tokens, top_k, **i32),
            "out_permuted_idx_to_expanded_idx": torch.empty(tiles * tile_tokens_dim, **i32),
            "out_total_num_padded_tokens": torch.empty(1, **i32), "out_num_non_exiting_tiles": torch.empty(1, **i32)}


def moe_sort(token_selected_experts: torch.Tensor, token_final_scales: torch.Tensor, num_experts: int, top_k: int,
             local_expert_offset: int = 0, num_local_experts: Optional[int] = None, tile_tokens_dim: int = 128, enable_pdl: bool = False,
             out_tile_idx_to_expert_idx: Optional[torch.Tensor] = None, out_tile_idx_to_mn_limit: Optional[torch.Tensor] = None,
             out_expanded_idx_to_permuted_idx: Optional[torch.Tensor] = None, out_permuted_idx_to_expanded_idx: Optional[torch.Tensor] = None,
             out_total_num_padded_tokens: Optional[torch.Tensor] = None, out_num_non_exiting_tiles: Optional[torch.Tensor] = None
             ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    """Group the ``[num_tokens, top_k]`` expert choices by local expert, every expert padded to whole tiles.  Returns
    ``(tile_idx_to_expert_idx [max_tiles], tile_idx_to_mn_limit [max_tiles], expanded_idx_to_permuted_idx [T, K] (-1 = expert of
    another rank), permuted_idx_to_expanded_idx [max_tiles * tile] (-1 = padding), total_num_padded_tokens [1], num_non_exiting_tiles [1])``;
    rows keep their token order inside an expert.  ``token_final_scales`` is not consumed (it travels to :func:`moe_unpermute`)."""
    t_, k_ = token_selected_experts.shape
    if k_ != top_k:
        raise ValueError(f"token_selected_experts has {k_} columns, top_k = {top_k}")
    local = num_local_experts if num_local_experts is not None else num_experts
    tile = int(tile_tokens_dim)
    dev = token_selected_experts.device
    n = t_ * k_
    max_tiles = get_max_num_tiles(t_, k_, local, tile)
    max_rows = max_tiles * tile
    e = token_selected_experts.reshape(-1).to(torch.int64) - local_expert_offset
    valid = (e >= 0) & (e < local)
    key = torch.where(valid, e, torch.full_like(e, local))                       # foreign experts sort behind every local one
    order = torch.argsort(key, stable=True)
    counts = torch.zeros(local + 1, dtype=torch.int64, device=dev).scatter_add_(0, key, torch.ones_like(key))[:local]
    padded = (counts + tile - 1) // tile * tile
    pad_end = padded.cumsum(0)
    pad_start = pad_end - padded
    start = counts.cumsum(0) - counts
    skey = key[order]
    svalid = skey < local
    se = skey.clamp(max=local - 1)
    perm = pad_start[se] + (torch.arange(n, device=dev) - start[se])             # destination row of the j-th sorted entry
    e2p = torch.full((n,), -1, dtype=torch.int64, device=dev)
    e2p[order] = torch.where(svalid, perm, torch.full_like(perm, -1))
    p2e = torch.full((max_rows + 1,), -1, dtype=torch.int64, device=dev)          # slot max_rows swallows the foreign entries
    p2e[torch.where(svalid, perm, torch.full_like(perm, max_rows))] = torch.where(svalid, order, torch.full_like(order, -1))
    p2e = p2e[:max_rows]
    tile_row0 = torch.arange(max_tiles, device=dev) * tile
    tile_e = torch.searchsorted(pad_end, tile_row0, right=True)                   # first expert whose padded range ends beyond the tile
    live = tile_row0 < pad_end[-1] if local else torch.zeros(max_tiles, dtype=torch.bool, device=dev)
    te = tile_e.clamp(max=max(local - 1, 0))
    limit = torch.minimum(tile_row0 + tile, pad_start[te] + counts[te])
    outs = (torch.where(live, te, torch.full_like(te, -1)), torch.where(live, limit, torch.zeros_like(limit)), e2p.view(t_, k_), p2e,
            pad_end[-1:].clone() if local else torch.zeros(1, dtype=torch.int64, device=dev),
            (pad_end[-1:] // tile) if local else torch.zeros(1, dtype=torch.int64, device=dev))
    bufs = (out_tile_idx_to_expert_idx, out_tile_idx_to_mn_limit, out_expanded_idx_to_permuted_idx, out_permuted_idx_to_expanded_idx,
            out_total_num_padded_tokens, out_num_non_exiting_tiles)
    res = []
    for val, buf in zip(outs, bufs):
        if buf is None:
            res.append(val.to(torch.int32))
        else:
            buf.copy_(val.view_as(buf))
            res.append(buf)
    return tuple(res)


def _row_valid(tile_idx_to_mn_limit: torch.Tensor, rows: int, tile_size: int) -> torch.Tensor:
    """Mask of the permuted rows that carry a token (row index below its tile's limit)."""
    r = torch.arange(rows, device=tile_idx_to_mn_limit.device)
    t = (r // tile_size).clamp(max=tile_idx_to_mn_limit.numel() - 1)
    return (r < tile_idx_to_mn_limit[t].to(torch.int64)) & (r // tile_size < tile_idx_to_mn_limit.numel())


def moe_permute(input: torch.Tensor, permuted_output: torch.Tensor, tile_idx_to_mn_limit: torch.Tensor, permuted_idx_to_expanded_idx: torch.Tensor,
                num_non_exiting_tiles: torch.Tensor, max_num_permuted_tokens: int, top_k: int, tile_size: int, enable_pdl: bool = False,
                input_sf: Optional[torch.Tensor] = None, permuted_sf: Optional[torch.Tensor] = None) -> None:
    """``permuted_output[p] = input[permuted_idx_to_expanded_idx[p] // top_k]`` for the rows that carry a token, zeros for padding rows.
    ``input`` may be any row format (bf16 / fp16 / e4m3 / packed fp4 bytes).  With ``input_sf`` (linear ``[num_tokens, hidden / 16]`` NVFP4
    scale bytes) ``permuted_sf`` receives the gathered scale rows in the 128x4 swizzled layout of the block-scaled grouped GEMM."""
    rows = min(int(max_num_permuted_tokens), permuted_output.shape[0])
    idx = permuted_idx_to_expanded_idx[:rows].to(torch.int64)
    has = idx >= 0
    tok = (idx.clamp(min=0) // top_k)
    gathered = input.view(torch.uint8 if input.dtype == torch.uint8 else input.dtype)[tok]
    permuted_output[:rows].copy_(torch.where(has[:, None], gathered, torch.zeros_like(gathered)))
    if input_sf is not None:
        if permuted_sf is None:
            raise ValueError("moe_permute: permuted_sf is required with input_sf")
        from ..quantization.fp4 import block_scale_interleave

        sf = input_sf.view(torch.uint8).reshape(input.shape[0], -1)[tok]
        sf = torch.where(has[:, None], sf, torch.zeros_like(sf))
        swz = block_scale_interleave(sf).reshape(-1)      # round_up(rows, 128) x round_up(hidden / 16, 4) bytes
        dst = permuted_sf.view(torch.uint8).reshape(-1)
        if dst.numel() < swz.numel():
            raise ValueError(f"moe_permute: permuted_sf holds {dst.numel()} bytes, the swizzled scale factors of {rows} rows need {swz.numel()}")
        dst[: swz.numel()].copy_(swz)


def moe_unpermute(permuted_input: torch.Tensor, output: torch.Tensor, expanded_idx_to_permuted_idx: torch.Tensor, topk_scales: torch.Tensor,
                  num_tokens: int, top_k: int, enable_pdl: bool = False) -> None:
    """``output[i] = sum_k topk_scales[i, k] * permuted_input[expanded_idx_to_permuted_idx[i, k]]`` (entries of -1 contribute nothing).
    On CUDA this is the ``moe_finalize`` kernel of the fused pipeline (one pass, fp32 accumulation, 16-byte vectors)."""
    e2p = expanded_idx_to_permuted_idx.reshape(num_tokens, top_k)
    hidden = permuted_input.shape[-1]
    if (permuted_input.is_cuda and permuted_input.dtype in (torch.float16, torch.bfloat16) and output.dtype == permuted_input.dtype
            and hidden % 8 == 0 and permuted_input.is_contiguous() and output.is_contiguous()):
        jit.load("moe").call("moe_finalize", permuted_input, output, e2p.to(torch.int32).contiguous(), topk_scales.float().contiguous(),
                             num_tokens, top_k, hidden, 0, dtype_code(permuted_input.dtype), 1 if enable_pdl else 0, stream_ptr(permuted_input))
        return
    idx = e2p.to(torch.int64)
    rows = permuted_input[idx.clamp(min=0)].float()                                # [T, K, H]
    w = torch.where(idx >= 0, topk_scales.reshape(num_tokens, top_k).float(), torch.zeros((), device=idx.device))
    output.copy_((rows * w[..., None]).sum(1).to(output.dtype))


def moe_output_memset(output: torch.Tensor, tile_idx_to_mn_limit: torch.Tensor, expanded_idx_to_permuted_idx: torch.Tensor,
                      permuted_idx_to_expanded_idx: torch.Tensor, num_non_exiting_tiles: torch.Tensor, max_num_permuted_tokens: int, top_k: int,
                      tile_size: int, enable_pdl: bool = False) -> None:
    """Zero the rows of ``output`` that a scatter-add style finalize will accumulate into: every token with at least one local expert
    (tokens whose experts all live on other ranks are left untouched, as in the reference)."""
    touched = (expanded_idx_to_permuted_idx.reshape(output.shape[0], top_k) >= 0).any(-1)
    output.mul_((~touched).to(output.dtype)[:, None])


def moe_output_memset_inplace(output: torch.Tensor) -> None:
    """Zero the whole output on the current stream (one memset node)."""
    output.zero_()


def moe_activation(input: torch.Tensor, output: torch.Tensor, tile_idx_to_mn_limit: torch.Tensor, num_non_exiting_tiles: torch.Tensor,
                   activation_type: MoeActivationType, max_num_permuted_tokens: int, tile_size: int, enable_pdl: bool = False) -> None:
    """Activation over the permuted rows.  Gated types (Swiglu, Geglu) read ``[rows, 2 * I]`` = ``[linear | gate]`` and write
    ``act(gate) * linear``; the others map ``[rows, I]`` element-wise.  Padding rows are written as zeros."""
    rows = min(int(max_num_permuted_tokens), input.shape[0], output.shape[0])
    inter = output.shape[-1]
    act = MoeActivationType(int(activation_type))
    x = input[:rows]
    if act in (MoeActivationType.Swiglu, MoeActivationType.Geglu):
        if x.shape[-1] != 2 * inter:
            raise ValueError("gated activations expect input [rows, 2 * intermediate] = [linear | gate]")
        if x.is_cuda and x.dtype in (torch.float16, torch.bfloat16) and x.is_contiguous() and output.is_contiguous() and output.dtype == x.dtype:
            from ..activation import _act_and_mul

            y = torch.empty(rows, inter, dtype=x.dtype, device=x.device)
            _act_and_mul("silu" if act == MoeActivationType.Swiglu else "gelu", x, y, enable_pdl, gate_second=True)   # native fused kernel
        else:
            lin, gate = x[:, :inter].float(), x[:, inter:].float()
            y = ((torch.nn.functional.silu(gate) if act == MoeActivationType.Swiglu else torch.nn.functional.gelu(gate)) * lin).to(output.dtype)
    else:
        xf = x.float()
        y = {MoeActivationType.Gelu: torch.nn.functional.gelu, MoeActivationType.Relu: torch.relu, MoeActivationType.Silu: torch.nn.functional.silu,
             MoeActivationType.Identity: lambda t: t}[act](xf).to(output.dtype)
    keep = _row_valid(tile_idx_to_mn_limit, rows, tile_size)
    output[:rows].copy_(torch.where(keep[:, None], y, torch.zeros_like(y)))


def _activation(kind: MoeActivationType):
    def fn(input: torch.Tensor, output: torch.Tensor, tile_idx_to_mn_limit: torch.Tensor, num_non_exiting_tiles: torch.Tensor,
           max_num_permuted_tokens: int, tile_size: int, enable_pdl: bool = False) -> None:
        moe_activation(input, output, tile_idx_to_mn_limit, num_non_exiting_tiles, kind, max_num_permuted_tokens, tile_size, enable_pdl)

    fn.__name__ = f"moe_{kind.name.lower()}"
    fn.__doc__ = f":func:`moe_activation` with ``MoeActivationType.{kind.name}``."
    return fn


moe_swiglu = _activation(MoeActivationType.Swiglu)
moe_geglu = _activation(MoeActivationType.Geglu)
moe_gelu = _activation(MoeActivationType.Gelu)
moe_silu = _activation(MoeActivationType.Silu)
moe_relu = _activation(MoeActivationType.Relu)

__all__ = ["MoeActivationType", "get_max_num_tiles", "get_max_num_permuted_tokens", "allocate_moe_sort_buffers", "moe_sort", "moe_permute",
           "moe_unpermute", "moe_output_memset", "moe_output_memset_inplace", "moe_activation", "moe_swiglu", "moe_geglu", "moe_gelu",
           "moe_silu", "moe_relu"]
