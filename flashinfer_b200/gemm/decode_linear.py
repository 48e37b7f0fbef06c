"""Decode-layer linears: the small-M GEMM of a decode step with the neighbouring ops folded into its epilogue
(csrc/gemm/decode_linear_sm100.cu).  A Llama layer becomes five launches: QKV (+RMSNorm +RoPE +paged-KV append), attention,
O (+all-reduce +residual +norm statistics), gate/up (+RMSNorm +SwiGLU), down (+all-reduce +residual +norm statistics).

Parity: the reference composes these from ``tgv_gemm_sm100`` / ``mm_bf16`` (flashinfer/gemm/gemm_base.py:485,1446),
``fused_add_rmsnorm`` (flashinfer/norm.py), ``apply_llama31_rope_pos_ids_inplace`` + ``append_paged_kv_cache``
(flashinfer/rope.py, flashinfer/page.py), ``silu_and_mul`` (flashinfer/activation.py) and
``trtllm_allreduce_fusion(kARResidualRMSNorm)`` (flashinfer/comm/trtllm_ar.py:951-1060).

RMSNorm folding: ``rmsnorm(x) * g @ W^T == rstd(x) * (x @ (W * g)^T)`` - ``fold_rmsnorm_weight`` bakes ``g`` into ``W`` once, the
kernel applies ``rstd`` per token to the fp32 accumulators; the sum of squares comes from the residual epilogue of the previous
GEMM (or :func:`decode_prep` for the embedding rows).  Every function has a fp32 PyTorch path (CPU tensors) that is the oracle of
the CUDA kernels.
"""
from __future__ import annotations

import contextlib
import math
import os
from typing import Optional, Tuple

import torch

from .. import jit
from ..utils import dtype_code, stream_ptr

EPI_PLAIN, EPI_GATED_SILU, EPI_RESIDUAL, EPI_ROPE_APPEND = 0, 1, 2, 3
_SUMSQ_ROWS = 64  # row pitch of the sum-of-squares accumulators (= max decode batch of the kernel)


# ------------------------------------------------------------------ weight preparation (load time)
def fold_rmsnorm_weight(w: torch.Tensor, norm_weight: torch.Tensor, weight_bias: float = 0.0) -> torch.Tensor:
    """``W [N, K]`` -> ``W * (g + weight_bias)[None, :]`` (the RMSNorm gain moves into the weight columns)."""
    return (w.float() * (norm_weight.float() + weight_bias)[None, :]).to(w.dtype).contiguous()


def permute_rope_rows(wqkv: torch.Tensor, num_q_heads: int, num_kv_heads: int, head_dim: int) -> torch.Tensor:
    """NeoX-style RoPE rotates ``(j, j + head_dim/2)``: re-order the Q and K rows of every head to ``(0, hd/2, 1, hd/2+1, ...)`` so
    that a rotation pair sits in adjacent accumulator columns of the QKV GEMM (the epilogue writes them back un-permuted)."""
    k = wqkv.shape[1]
    nqk = (num_q_heads + num_kv_heads) * head_dim
    qk = wqkv[:nqk].view(num_q_heads + num_kv_heads, 2, head_dim // 2, k).transpose(1, 2).reshape(nqk, k)
    return torch.cat([qk, wqkv[nqk:]], 0).contiguous()


def to_block_major_k(w: torch.Tensor, block_k: int = 64) -> torch.Tensor:
    """``W [N, K]`` -> BlockMajorK ``[K / 64, N, 64]`` (reference WeightLayout.BlockMajorK, flashinfer/tllm_enums.py:141-150): all N rows
    of one 64-wide K block are contiguous, so the TMA box of a weight tile is ONE contiguous ``BN x 128 B`` chunk of HBM instead of
    ``BN`` 128-byte pieces a whole row pitch apart."""
    n, k = w.shape
    if k % block_k:
        raise ValueError("to_block_major_k: K must be a multiple of 64")
    return w.view(n, k // block_k, block_k).transpose(0, 1).contiguous()


def from_block_major_k(wb: torch.Tensor) -> torch.Tensor:
    kb, n, bk = wb.shape
    return wb.transpose(0, 1).reshape(n, kb * bk)


# ------------------------------------------------------------------ tensor-parallel context
class FusedLinearTP:
    """Receive buffers of the in-kernel all-reduce (``EPI_RESIDUAL`` with tp > 1): ``[3 rotating][world slots][64 rows][hidden]``
    in a symmetric heap (+ its NVLS multicast alias), pre-filled with the sentinel (-0.0), and the device-side epoch word that
    selects the rotating buffer (CUDA-graph replay safe: nothing about the rotation is baked into the launch arguments)."""

    ONE_SHOT, TWO_SHOT = 1, 2

    def __init__(self, group, max_tokens: int, hidden: int, dtype: torch.dtype = torch.bfloat16, heap=None, algo: Optional[int] = None):
        """``algo``: ``ONE_SHOT`` - every rank multicasts its partial strip and gathers all ``world`` partials (one NVLink hop,
        ``world x`` ingress: best for 2 ranks); ``TWO_SHOT`` - reduce-scatter by push to the row owner (row ``m`` belongs to rank
        ``m % world``), the owner adds the residual and multicasts the new residual rows (two hops, ``(1 + 1/world) x`` ingress:
        best for 8 ranks); ``None`` / 0 - by world size (env ``FIB200_DL_AR_ALGO`` overrides).  The algorithm is a property
        of the context because the two use the rotating receive buffers differently: do not change it between calls."""
        import torch.distributed as dist

        from ..comm.symm import SymmetricHeap

        self.group = group if group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
        if self.world > 8:
            raise ValueError("FusedLinearTP: the one-shot push all-reduce supports at most 8 ranks")
        esz = torch.empty(0, dtype=dtype).element_size()
        self.rows, self.hidden, self.dtype = _SUMSQ_ROWS, hidden, dtype
        self.slot_elems = self.rows * hidden
        nbytes = 3 * self.world * self.slot_elems * esz
        self.heap = heap if heap is not None else SymmetricHeap(self.group, nbytes + 16384)
        buf, off = self.heap.alloc(nbytes)
        buf.view(torch.int16).fill_(-32768)  # 0x8000 = -0.0 in bf16 / fp16: "nothing has arrived yet"
        self.recv = buf.view(dtype)
        self.mc_recv = self.heap.mc(off)
        self.peer_recv = self.heap.peer_ptr_table(off)
        self.epoch = torch.zeros(4, dtype=torch.int32, device=self.heap.device)
        algo = int(algo or os.environ.get("FIB200_DL_AR_ALGO", "0"))
        if algo == 0:
            # measured on B200 (tools/tp_breakdown.py, O projection, fused GEMM + all-reduce): world 2: 12.6 (one-shot) vs 13.9 us,
            # world 4: 14.1 vs 14.3 us; at world 8 the one-shot ingress is 8 x the strip and the two-shot wins
            algo = self.TWO_SHOT if (self.world >= 8 and 64 % self.world == 0) else self.ONE_SHOT
        if algo == self.TWO_SHOT and 64 % self.world:
            raise ValueError("FusedLinearTP: the two-shot all-reduce needs a world size that divides 64")
        self.algo = algo
        self.heap.barrier()


class _ptr:
    """Raw device address through the uniform C ABI (void*)."""

    def __init__(self, v: int):
        self.v = int(v)


# ------------------------------------------------------------------ autotuner client: the tile plan (BN, split-K cluster size)
def admissible_plans(n: int, k: int, sms: int = 148):
    """Single-wave tile plans of the kernel for ``W [N, K]``: weight-tile width BN (multiple of 16 x S) and split-K cluster size S
    (clusters of 4 / 8 CTAs are placeable on 132 of 148 SMs); the same enumeration as tools/dl_sweep.py.  A plan is encoded as
    ``BN * 16 + S`` (tactic ``-1`` = the planner in ``dlinear_run``)."""
    kb = k // 64
    out = []
    for s in (1, 2, 4, 8):
        if kb < s:
            continue
        for bn in range(16 * s, 257, 16 * s):
            tiles = (n + bn - 1) // bn
            lim = sms if s < 4 else sms * 132 // 148
            if tiles * s > lim or tiles * s < min(40, lim // 2):
                continue
            out.append(bn * 16 + s)
    return out


def _tuned_plan(x: torch.Tensor, w: torch.Tensor, n: int, k: int, epi: int, launch):
    """``(bn, split_k)`` for this call: (0, 0) = the built-in planner, unless the autotuner is profiling (``with autotune():``) or
    holds a tuned choice for this (bucketed M, N, K, epilogue) - the plan table in the kernel was swept on the Llama-3-8B shapes
    (profiles/decode_linear_plans.md); other models tune theirs once and ship the JSON.  ``launch(bn, split_k)`` runs the op on
    the live tensors (the caller gives idempotent scratch outputs to the residual epilogue while profiling)."""
    from ..autotuner import AutoTuner, DynamicTensorSpec, TunableRunner, TuningConfig

    tuner = AutoTuner.get()
    if not (tuner.is_tuning_mode or tuner.profiling_cache):
        return 0, 0

    class _PlanRunner(TunableRunner):
        def get_valid_tactics(self, inputs, profile):
            return [-1] + admissible_plans(n, k)

        def forward(self, inputs, tactic=-1, do_preparation=False, **kwargs):
            t = int(tactic)
            return launch(0, 0) if t < 0 else launch(t // 16, t % 16)

    cfg = TuningConfig(dynamic_tensor_specs=(DynamicTensorSpec((0,), (0,)),), use_cold_l2_cache=True, synthesize_buckets=False)
    _, tactic = tuner.choose_one("decode_linear", [_PlanRunner()], cfg, [x, w], extras=(int(epi), int(n), int(k), str(x.dtype), w.dim()))
    t = int(tactic)
    return (0, 0) if t < 0 else (t // 16, t % 16)


# ------------------------------------------------------------------ the op
def _rstd(row_sumsq: Optional[torch.Tensor], m: int, norm_dim: int, eps: float) -> Optional[torch.Tensor]:
    if row_sumsq is None:
        return None
    return torch.rsqrt(row_sumsq[:m].float() / norm_dim + eps)


def _reference(x, w, epi, out, bias, row_sumsq, norm_dim, eps, residual, sumsq_out, tp, cos_sin, cache_row, k_cache, v_cache,
               hq, hkv, head_dim, interleave, c_sh=0):
    m = x.shape[0]
    acc = x.float() @ w.float().t()
    rs = _rstd(row_sumsq, m, norm_dim, eps)
    if rs is not None and epi != EPI_RESIDUAL:
        acc = acc * rs[:, None]
    if epi == EPI_PLAIN:
        if bias is not None:
            acc = acc + bias.float()
        out.copy_(acc.to(out.dtype))
        return out
    if epi == EPI_GATED_SILU:
        out.copy_((torch.nn.functional.silu(acc[:, 0::2]) * acc[:, 1::2]).to(out.dtype))
        return out
    if epi == EPI_RESIDUAL:
        if tp is not None and tp.world > 1:
            import torch.distributed as dist

            part = acc.to(x.dtype).float()  # partials travel in the activation dtype
            dist.all_reduce(part, group=tp.group)
            acc = part
        new = (residual[:m].float() + acc).to(residual.dtype)
        residual[:m].copy_(new)
        if sumsq_out is not None:
            sumsq_out[:m] += new.float().pow(2).sum(-1)
        return residual
    # rope + append
    n_q, n_k = hq * head_dim, hkv * head_dim
    half = head_dim // 2
    cs, sn = cos_sin[:m, :half].float(), cos_sin[:m, half:].float()

    def rope(t, heads):
        t = t.view(m, heads, head_dim)
        if interleave:
            x1, x2 = t[..., 0::2], t[..., 1::2]
            o = torch.stack([x1 * cs[:, None] - x2 * sn[:, None], x2 * cs[:, None] + x1 * sn[:, None]], -1).flatten(-2)
        else:  # weights were permuted: column 2j = element j, column 2j+1 = element j + hd/2
            x1, x2 = t[..., 0::2], t[..., 1::2]
            o = torch.cat([x1 * cs[:, None] - x2 * sn[:, None], x2 * cs[:, None] + x1 * sn[:, None]], -1)
        return o

    q = rope(acc[:, :n_q], hq)
    k = rope(acc[:, n_q:n_q + n_k], hkv)
    v = acc[:, n_q + n_k:].view(m, hkv, head_dim)
    out.view(m, hq, head_dim).copy_(q.to(out.dtype))
    kf, vf = k_cache.view(-1), v_cache.view(-1)
    for i in range(m):
        base = int(cache_row[i])
        for h in range(hkv):
            kf[base + h * c_sh: base + h * c_sh + head_dim] = k[i, h].to(k_cache.dtype)
            vf[base + h * c_sh: base + h * c_sh + head_dim] = v[i, h].to(v_cache.dtype)
    return out


def _head_stride(cache: torch.Tensor, head_dim: int) -> int:
    """Element stride between KV heads inside one (page, slot) row of an NHD cache ``[P, page, H, D]`` (HND callers pass
    ``head_stride=`` explicitly: ``page_size * D``)."""
    return int(cache.stride(-2)) if cache.dim() == 4 else head_dim


def decode_linear(x: torch.Tensor, w: torch.Tensor, epi: int = EPI_PLAIN, *, out: Optional[torch.Tensor] = None,
                  bias: Optional[torch.Tensor] = None, row_sumsq: Optional[torch.Tensor] = None, norm_dim: Optional[int] = None,
                  eps: float = 1e-5, residual: Optional[torch.Tensor] = None, sumsq_out: Optional[torch.Tensor] = None,
                  tp: Optional[FusedLinearTP] = None, cos_sin: Optional[torch.Tensor] = None,
                  cache_row: Optional[torch.Tensor] = None, k_cache: Optional[torch.Tensor] = None,
                  v_cache: Optional[torch.Tensor] = None, head_stride: Optional[int] = None, num_q_heads: int = 0,
                  num_kv_heads: int = 0, head_dim: int = 0, interleave: bool = False, bn: int = 0, split_k: int = 0,
                  smem_kb: int = 0, enable_pdl: bool = True) -> torch.Tensor:
    """``epilogue(x[M<=64, K] @ w[N, K]^T)``; see the module docstring for the four epilogues.

    * ``row_sumsq`` (fp32 ``[>=M]``): per-token sum of squares of the un-normalised input over ``norm_dim`` columns; the output is
      scaled by ``rsqrt(row_sumsq / norm_dim + eps)`` (RMSNorm folded; pass weights from :func:`fold_rmsnorm_weight`).
    * ``EPI_GATED_SILU``: ``w`` rows interleaved (gate_0, up_0, ...) (:func:`flashinfer_b200.gemm.dense.interleave_gate_up`),
      ``out [M, N/2]``.
    * ``EPI_RESIDUAL``: ``residual[:M] += x @ w^T`` (summed over the ``tp`` ranks in-kernel), ``sumsq_out[:M] += sum(residual^2)``.
    * ``EPI_ROPE_APPEND``: ``w`` from :func:`permute_rope_rows` (unless ``interleave``), ``out [M, Hq * D]`` receives RoPE(Q),
      RoPE(K) and V are written into ``k_cache`` / ``v_cache`` at ``cache_row`` (from :func:`decode_prep`)."""
    m, k = x.shape
    blockk = w.dim() == 3  # BlockMajorK weights from to_block_major_k: [K / 64, N, 64]
    n = w.shape[1] if blockk else w.shape[0]
    if (w.shape[0] * w.shape[2] if blockk else w.shape[1]) != k:
        raise ValueError(f"decode_linear: K mismatch {tuple(w.shape)} vs {tuple(x.shape)}")
    if m > _SUMSQ_ROWS:
        raise ValueError("decode_linear handles decode batches of at most 64 tokens (use gemm.linear beyond)")
    norm_dim = norm_dim or k
    if epi == EPI_RESIDUAL:
        if residual is None:
            raise ValueError("EPI_RESIDUAL needs residual=")
        out = residual
    elif out is None:
        cols = n // 2 if epi == EPI_GATED_SILU else (num_q_heads * head_dim if epi == EPI_ROPE_APPEND else n)
        out = torch.empty(m, cols, dtype=x.dtype, device=x.device)
    if epi == EPI_ROPE_APPEND and (cos_sin is None or cache_row is None or k_cache is None or v_cache is None):
        raise ValueError("EPI_ROPE_APPEND needs cos_sin / cache_row / k_cache / v_cache")
    c_sh = 0
    if epi == EPI_ROPE_APPEND:
        c_sh = head_stride if head_stride is not None else _head_stride(k_cache, head_dim)
    if not x.is_cuda:
        return _reference(x, from_block_major_k(w) if blockk else w, epi, out, bias, row_sumsq, norm_dim, eps, residual, sumsq_out, tp, cos_sin, cache_row, k_cache,
                          v_cache, num_q_heads, num_kv_heads, head_dim, interleave, c_sh)
    if x.dtype not in (torch.float16, torch.bfloat16) or w.dtype != x.dtype:
        raise TypeError("decode_linear: x / w must both be float16 or bfloat16")
    if x.stride(-1) != 1 or w.stride(-1) != 1 or out.stride(-1) != 1:
        raise ValueError("decode_linear: innermost dimensions must be contiguous")
    world, rank = (tp.world, tp.rank) if (tp is not None and epi == EPI_RESIDUAL) else (1, 0)
    ar = tp if world > 1 else None
    if ar is not None and (n != ar.hidden or x.dtype != ar.dtype):
        raise ValueError("decode_linear: the all-reduce context was built for another hidden size / dtype")
    mod = jit.load("decode_linear_sm100")

    def launch(bn_, s_, residual_=residual, sumsq_=sumsq_out, out_=out):
        mod.call(
            "dlinear_run", x, w, m, n, k, x.stride(0), w.stride(1) if blockk else w.stride(0), int(epi), out_, out_.stride(0), bias,
            row_sumsq, 1.0 / float(norm_dim), float(eps), residual_, residual_.stride(0) if residual_ is not None else 0, sumsq_, world,
            rank, ar.recv if ar else None, ar.hidden if ar else 0, ar.slot_elems if ar else 0, _ptr(ar.mc_recv) if ar else None,
            ar.epoch if ar else None, ar.peer_recv if ar else None, int(ar.algo) if ar else 0, cos_sin, cache_row, k_cache, v_cache,
            int(c_sh), int(num_q_heads), int(num_kv_heads), int(head_dim), 1 if interleave else 0, int(bn_), int(s_), int(smem_kb),
            1 if blockk else 0, dtype_code(x.dtype), 1 if enable_pdl else 0, stream_ptr(x))
        return out_

    if bn == 0 and split_k == 0 and ar is None and not torch.cuda.is_current_stream_capturing():
        scratch = {}

        def probe(bn_, s_):  # profiling runs must not accumulate into the caller's residual stream
            if epi != EPI_RESIDUAL:
                return launch(bn_, s_)
            if not scratch:
                scratch["r"] = residual.clone()
                scratch["s"] = torch.zeros_like(sumsq_out) if sumsq_out is not None else None
            return launch(bn_, s_, scratch["r"], scratch["s"], scratch["r"])

        bn, split_k = _tuned_plan(x, w, n, k, epi, probe)
    launch(bn, split_k)
    return out


def decode_prep(tokens: torch.Tensor, embed: torch.Tensor, residual: torch.Tensor, sumsq: torch.Tensor, positions: torch.Tensor,
                kv_indptr: torch.Tensor, kv_indices: torch.Tensor, page_size: int, page_stride: int, slot_stride: int,
                cos_sin: torch.Tensor, cache_row: torch.Tensor, head_dim: int, batch_indices: Optional[torch.Tensor] = None,
                interleave: bool = False, rope_scale: float = 1.0, rope_theta: float = 1e4,
                llama31: Optional[Tuple[float, float, float]] = None, enable_pdl: bool = False) -> None:
    """One launch at the top of a decode step: ``residual[m] = embed[tokens[m]]``, ``sumsq[0, m] = sum(residual[m]^2)``,
    ``sumsq[1:] = 0`` (the per-layer accumulators), ``cos_sin[m] = cos | sin`` of ``positions[m]`` (Llama-3.1 scaling optional),
    ``cache_row[m]`` = element offset of the token's KV slot (``page * page_stride + slot * slot_stride``)."""
    m = tokens.numel()
    n_sumsq = sumsq.shape[0]
    if not tokens.is_cuda:
        residual[:m].copy_(embed[tokens.long()])
        sumsq.zero_()
        sumsq[0, :m] = residual[:m].float().pow(2).sum(-1)
        half = head_dim // 2
        inv = torch.pow(torch.tensor(float(rope_theta)), -torch.arange(0, half, dtype=torch.float32) * 2 / head_dim)
        if llama31 is not None:
            low, high, old_ctx = llama31
            smooth = (inv * (old_ctx / (2 * math.pi * (high - low))) - 1.0 / (high / low - 1.0)).clamp(0, 1)
            inv = (1 - smooth) * (inv / rope_scale) + smooth * inv
        else:
            inv = inv / rope_scale
        ang = positions[:m].float()[:, None] * inv[None]
        cos_sin[:m, :half] = torch.cos(ang)
        cos_sin[:m, half:] = torch.sin(ang)
        b = batch_indices.long() if batch_indices is not None else torch.arange(m)
        pos = positions[:m].long()
        page = kv_indices.long()[kv_indptr.long()[b] + pos // page_size]
        cache_row[:m] = page * page_stride + (pos % page_size) * slot_stride
        return
    low, high, old_ctx = llama31 if llama31 is not None else (1.0, 1.0, 1.0)
    jit.load("decode_linear_sm100").call(
        "decode_prep_run", tokens, embed, residual, embed.stride(0), residual.stride(0), embed.shape[1], sumsq, n_sumsq,
        positions, batch_indices, kv_indptr, kv_indices, int(page_size), int(page_stride), int(slot_stride), cos_sin, cache_row,
        int(head_dim), 1 if interleave else 0, float(rope_scale), float(rope_theta), 1 if llama31 is not None else 0, float(low),
        float(high), float(old_ctx), m, dtype_code(embed.dtype), 1 if enable_pdl else 0, stream_ptr(embed))


PROFILER_EVENTS = ("setup", "wait_prev_grid", "weight_prefetch_issued", "main_loop", "splitk_exchange", "epilogue", "all_reduce")
PROFILER_GROUPS = ("tma_producer", "mma_issuer", "epilogue")


@contextlib.contextmanager
def profiled(buf: torch.Tensor, max_events: int = 32):
    """Record the NEXT :func:`decode_linear` call of this thread with the intra-kernel profiler (reference flashinfer/profiler,
    include/flashinfer/profiler.cuh): inside the block the op runs from the ``-DFIB200_ENABLE_PROFILER`` build of the kernel
    (module ``decode_linear_sm100_prof``) and writes (event, globaltimer) pairs per (CTA, warp role) into ``buf``
    (``profiler.alloc_profiler_buffer(grid, 3, max_events)``); ``profiler.export_to_perfetto_trace(buf, PROFILER_EVENTS, path,
    max_events, PROFILER_GROUPS)`` turns it into a Perfetto timeline."""
    with jit.redirect({"decode_linear_sm100": "decode_linear_sm100_prof"}):
        jit.load("decode_linear_sm100").call("dlinear_set_profiler", buf, int(max_events))
        yield
