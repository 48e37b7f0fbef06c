"""Tensor-parallel all-reduce fused with residual-add + RMSNorm, issued from inside one CUDA kernel
over NVLink/NVSwitch (csrc/comm/allreduce.cu).  Parity: reference flashinfer/comm/trtllm_ar.py
(trtllm_allreduce_fusion :951, trtllm_custom_all_reduce :809), comm/allreduce.py (allreduce_fusion :460),
comm/trtllm_mnnvl_ar.py and comm/vllm_ar.py.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist

from .. import jit
from ..utils import dtype_code, stream_ptr
from .symm import SymmetricHeap

_MAX_BLOCKS = 296


class TPCommunicator:
    """Per-process-group communicator for the TP hot path.

    * ``gemm_out(tokens)`` hands out the next (ping-pong) symmetric input buffer: the producer GEMM
      writes its partial output straight into it (zero copy).
    * ``allreduce_add_rmsnorm`` reduces that buffer across ranks in-switch (NVLS ``multimem.ld_reduce``)
      and applies ``residual += sum; out = rmsnorm(residual) * weight`` in the same kernel.
    """

    def __init__(self, group: Optional[dist.ProcessGroup], max_tokens: int, hidden: int,
                 dtype: torch.dtype = torch.bfloat16, use_nvls: bool = True):
        self.group = group if group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
        self.max_tokens, self.hidden, self.dtype = max_tokens, hidden, dtype
        esz = torch.empty(0, dtype=dtype).element_size()
        buf_bytes = max_tokens * hidden * esz
        sig_bytes = 2 * _MAX_BLOCKS * 16 * 4
        small_bytes = 1 << 20
        self._argmax_rows = max(256, max_tokens)
        argmax_bytes = 2 * self.world * self._argmax_rows * 16
        self.heap = SymmetricHeap(self.group, sig_bytes + 3 * buf_bytes + small_bytes + argmax_bytes + 16384)
        sig, self._sig_off = self.heap.alloc(sig_bytes)
        self._in = []
        for _ in range(2):
            v, off = self.heap.alloc(buf_bytes)
            self._in.append((v.view(dtype).view(max_tokens, hidden), off))
        v, self._out_off = self.heap.alloc(buf_bytes)
        self._out_sym = v.view(dtype).view(max_tokens, hidden)
        v, self._small_off = self.heap.alloc(small_bytes)
        self._small = v
        self._argmax_box, argmax_off = self.heap.alloc(argmax_bytes)  # zero-filled by the heap: tag 0 = "nothing has arrived"
        self._argmax_tab = self.heap.peer_ptr_table(argmax_off)
        self._argmax_epoch = torch.zeros(4, dtype=torch.int32, device=self.heap.device)
        self._sig_tab = self.heap.peer_ptr_table(self._sig_off)
        self._in_tab = [self.heap.peer_ptr_table(off) for _, off in self._in]
        self._out_tab = self.heap.peer_ptr_table(self._out_off)
        self._small_tab = self.heap.peer_ptr_table(self._small_off)
        self._epochs = torch.zeros(2 * _MAX_BLOCKS, dtype=torch.int32, device=self.heap.device)
        self.use_nvls = bool(use_nvls and self.heap.mc_ptr)
        self._push = None  # receive buffers of the one-shot push all-reduce: allocated on first use (3 x world x tokens x hidden)
        self._turn = 0
        self._mod = jit.load("comm_allreduce")
        self.heap.barrier()

    # ------------------------------------------------------------------ buffers
    def gemm_out(self, tokens: int) -> torch.Tensor:
        """Next symmetric input buffer ``[tokens, hidden]`` (ping-pong)."""
        self._turn ^= 1
        return self._in[self._turn][0][:tokens]

    def _locate(self, x: torch.Tensor):
        for i, (buf, off) in enumerate(self._in):
            if x.data_ptr() == buf.data_ptr():
                return i
        return -1

    # ------------------------------------------------------------------ collectives
    def allreduce_add_rmsnorm(self, x: torch.Tensor, residual: Optional[torch.Tensor], weight: Optional[torch.Tensor],
                              eps: float = 1e-6, out: Optional[torch.Tensor] = None, two_shot: Optional[bool] = None,
                              weight_bias: float = 0.0, quant_out: Optional[torch.Tensor] = None,
                              quant_scale: float = 0.0, enable_pdl: bool = True) -> torch.Tensor:
        """``s = sum_ranks(x); residual += s; out = rmsnorm(residual) * weight`` (``weight=None``: ``out = s``).

        ``x`` should be a buffer obtained from :meth:`gemm_out`; any other tensor is staged into one.
        One-shot (default for <= 1 MB): ``residual`` is replicated.  Two-shot: rank ``r`` owns rows
        ``r::world`` of ``residual`` (token-sharded residual stream); ``out`` is complete on every rank."""
        tokens, hidden = x.shape
        idx = self._locate(x)
        if idx < 0:
            buf = self.gemm_out(tokens)
            buf.copy_(x)
            idx = self._turn
        if two_shot is None:
            two_shot = x.numel() * x.element_size() > (1 << 20)
        if out is None:
            out = torch.empty_like(x)
        mc_in = self.heap.mc(self._in[idx][1]) if self.use_nvls else 0
        target = out
        if two_shot:
            target = self._out_sym[:tokens]
        self._mod.call(
            "allreduce_fusion_run", self._in_tab[idx], self._sig_tab, self._out_tab if two_shot else None,
            _ptr(mc_in), _ptr(self.heap.mc(self._out_off) if (two_shot and self.use_nvls) else 0), target, residual,
            weight, self._epochs, quant_out, tokens, hidden, self.rank, self.world, _MAX_BLOCKS, float(eps),
            float(weight_bias), float(quant_scale), 1 if two_shot else 0, dtype_code(x.dtype), 1 if enable_pdl else 0,
            stream_ptr(x),
        )
        if two_shot:
            out.copy_(target)
        return out

    # ------------------------------------------------------------------ one-shot push (Lamport) with fused prologue / epilogue
    PUSH_MAX_TOKENS = 256
    SF_LAYOUT = {"128x4": 0, "8x4": 1, "linear": 2}

    def _push_state(self):
        if self._push is None:
            from .symm import SymmetricHeap

            esz = torch.empty(0, dtype=self.dtype).element_size()
            rows = min(self.max_tokens, self.PUSH_MAX_TOKENS)
            slot = rows * self.hidden
            nbytes = 3 * self.world * slot * esz
            heap = SymmetricHeap(self.group, nbytes + 8192)
            buf, off = heap.alloc(nbytes)
            buf.view(torch.int16).fill_(-32768)  # -0.0 sentinel: "nothing has arrived"
            self._push = {"heap": heap, "recv": buf.view(self.dtype), "mc": heap.mc(off) if self.use_nvls else 0,
                          "peers": heap.peer_ptr_table(off), "epoch": torch.zeros(4, dtype=torch.int32, device=heap.device),
                          "rows": rows, "slot": slot}
            heap.barrier()
        return self._push

    def push_supported(self, tokens: int, hidden: int, dtype: torch.dtype) -> bool:
        return (hidden == self.hidden and dtype == self.dtype and tokens <= min(self.max_tokens, self.PUSH_MAX_TOKENS)
                and hidden % 16 == 0 and hidden <= 16384 and dtype in (torch.float16, torch.bfloat16))

    def push_allreduce(self, x: Optional[torch.Tensor], *, tokens: Optional[int] = None, ar_out: Optional[torch.Tensor] = None,
                       residual_in: Optional[torch.Tensor] = None, residual_out: Optional[torch.Tensor] = None,
                       rms_gamma: Optional[torch.Tensor] = None, norm_out: Optional[torch.Tensor] = None, eps: float = 1e-6,
                       weight_bias: float = 0.0, quant: str = "none", quant_out: Optional[torch.Tensor] = None,
                       scale_out: Optional[torch.Tensor] = None, scale_factor: Optional[torch.Tensor] = None,
                       sf_layout: str = "128x4", moe_reduction=None, moe_finalize=None, enable_pdl: bool = True) -> None:
        """One kernel: [MoE reduction | MoE finalize |] all-reduce (one-shot push over NVLink) [+ residual][+ RMSNorm][+ e4m3 / NVFP4
        quantisation with linear / 128x4 / 8x4 scale layout] (csrc/comm/allreduce.cu: allreduce_push_kernel).

        ``moe_reduction = (active_experts_token [E, T, H], scale [E, T], token_input [T, H])``;
        ``moe_finalize = (permuted_rows [P, H], expanded_idx_to_permuted_idx [T, K], expert_weights [T, K] or None, shared [T, H] or None)``."""
        st = self._push_state()
        mode, src, msc, mtok, e2p, mn = 0, x, None, None, None, 0
        if moe_reduction is not None:
            act, sc, tok = moe_reduction
            mode, src, msc, mtok, mn = 1, act.contiguous(), sc.float().contiguous(), tok.contiguous(), act.shape[0]
            tokens = tok.shape[0]
        elif moe_finalize is not None:
            rows, idx, w, shared = moe_finalize
            mode, src, e2p, mn = 2, rows.contiguous(), idx.int().contiguous(), idx.shape[1]
            msc = w.float().contiguous() if w is not None else None
            mtok = shared.contiguous() if shared is not None else None
            tokens = idx.shape[0]
        else:
            src = x.contiguous()
            tokens = x.shape[0] if tokens is None else tokens
        ref = src
        q = {"none": 0, "fp8": 1, "nvfp4": 2}[quant]
        sf = None
        if q:
            sf = scale_factor if isinstance(scale_factor, torch.Tensor) else torch.tensor([float(scale_factor if scale_factor is not None else 1.0)],
                                                                                         device=ref.device)
            sf = sf.float().reshape(-1)[:1].contiguous()
        self._mod.call("allreduce_push_run", mode, src, msc, mtok, e2p, mn, st["recv"], _ptr(st["mc"]), st["peers"], st["epoch"], st["slot"],
                       self.rank, self.world, tokens, self.hidden, st["rows"], ar_out, residual_in, residual_out, rms_gamma, norm_out,
                       float(eps), float(weight_bias), q, quant_out, scale_out, sf, self.SF_LAYOUT[sf_layout], dtype_code(self.dtype),
                       1 if enable_pdl else 0, stream_ptr(ref))

    def all_reduce(self, x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Plain sum all-reduce through the same kernel."""
        shape = x.shape
        x2 = x.reshape(-1, shape[-1]) if x.ndim > 1 else x.reshape(1, -1)
        if x2.shape[1] != self.hidden or x2.dtype != self.dtype:
            return self._all_reduce_small(x, out)
        res = self.allreduce_add_rmsnorm(x2, None, None, out=out.view_as(x2) if out is not None else None)
        return res.view(shape)

    def _all_reduce_small(self, x: torch.Tensor, out: Optional[torch.Tensor]) -> torch.Tensor:
        """fp32 all-reduce of a small tensor (<= 1 MB) via the symmetric scratch region."""
        n = x.numel()
        assert n * 4 <= self._small.numel(), "small all-reduce limited to 1 MB"
        pad = (-n) % 4
        stage = self._small.view(torch.float32)[: n + pad].view(1, n + pad)
        stage.zero_()
        stage[0, :n].copy_(x.reshape(-1).float())
        res = torch.empty(1, n + pad, dtype=torch.float32, device=x.device)
        self._mod.call(
            "allreduce_fusion_run", self._small_tab, self._sig_tab, None,
            _ptr(self.heap.mc(self._small_off) if self.use_nvls else 0), _ptr(0), res, None, None, self._epochs, None,
            1, n + pad, self.rank, self.world, _MAX_BLOCKS, 0.0, 0.0, 0.0, 0, dtype_code(torch.float32), 1, stream_ptr(x),
        )
        # the scratch region is reused by the next call: fence with a second (cheap) barrier round
        self._barrier_kernel(x)
        r = res[0, :n].view(x.shape).to(x.dtype)
        if out is not None:
            out.copy_(r)
            return out
        return r

    def _barrier_kernel(self, ref: torch.Tensor) -> None:
        dummy = torch.zeros(1, 4, dtype=torch.float32, device=ref.device)
        self._mod.call(
            "allreduce_fusion_run", self._small_tab, self._sig_tab, None, _ptr(0), _ptr(0), dummy, None, None,
            self._epochs, None, 0 + 1, 4, self.rank, self.world, _MAX_BLOCKS, 0.0, 0.0, 0.0, 0,
            dtype_code(torch.float32), 1, stream_ptr(ref),
        )

    def argmax_logits(self, logits: torch.Tensor, index_offset: int, out: Optional[torch.Tensor] = None,
                      out_val: Optional[torch.Tensor] = None, enable_pdl: bool = True) -> torch.Tensor:
        """Greedy sampling over a vocabulary-sharded LM head, one kernel: ``out[b] = argmax over all ranks' shards`` (global index
        = ``index_offset`` + local column; lowest index on ties, identical on every rank).  ``logits [B, shard]`` (f16 / bf16 / f32,
        row-contiguous), ``out`` int64 ``[B]`` (csrc/comm/allreduce.cu: argmax_push_kernel)."""
        b, shard = logits.shape
        if b > self._argmax_rows:
            raise ValueError(f"argmax_logits: at most {self._argmax_rows} rows")
        if logits.stride(-1) != 1:
            logits = logits.contiguous()
        if out is None:
            out = torch.empty(b, dtype=torch.int64, device=logits.device)
        self._mod.call("argmax_push_run", logits, logits.stride(0), shard, int(index_offset), self._argmax_box, self._argmax_tab,
                       self._argmax_epoch, self.rank, self.world, b, self._argmax_rows, out, out_val, dtype_code(logits.dtype),
                       1 if enable_pdl else 0, stream_ptr(logits))
        return out

    def argmax_gather(self, val: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
        """Global argmax over vocab shards: every rank contributes (max value, global index)."""
        b = val.numel()
        table = torch.zeros(self.world, 2, b, dtype=torch.float32, device=val.device)
        table[self.rank, 0] = val.float()
        table[self.rank, 1] = idx.float()
        table = self._all_reduce_small(table, None)
        best = table[:, 0].argmax(0)
        return table[:, 1].gather(0, best[None])[0].long()


_LOCAL_ARGMAX: dict = {}


def local_argmax(logits: torch.Tensor, out: Optional[torch.Tensor] = None, out_val: Optional[torch.Tensor] = None,
                 enable_pdl: bool = True) -> torch.Tensor:
    """Row-wise argmax of ``logits [B, V]`` (greedy sampling on one GPU) through the same kernel as
    :meth:`TPCommunicator.argmax_logits` with a group of one: one CTA of 1024 threads per row, lowest index on ties."""
    b, v = logits.shape
    if not logits.is_cuda:
        res = logits.float().argmax(-1)
        if out is not None:
            out.copy_(res)
            return out
        return res
    dev = logits.device
    st = _LOCAL_ARGMAX.get(dev.index)
    if st is None or st["rows"] < b:
        rows = max(256, b)
        box = torch.zeros(2 * rows * 16, dtype=torch.uint8, device=dev)
        st = {"rows": rows, "box": box, "tab": torch.tensor([box.data_ptr()], dtype=torch.int64),
              "epoch": torch.zeros(4, dtype=torch.int32, device=dev), "mod": jit.load("comm_allreduce")}
        _LOCAL_ARGMAX[dev.index] = st
    if logits.stride(-1) != 1:
        logits = logits.contiguous()
    if out is None:
        out = torch.empty(b, dtype=torch.int64, device=dev)
    st["mod"].call("argmax_push_run", logits, logits.stride(0), v, 0, st["box"], st["tab"], st["epoch"], 0, 1, b, st["rows"], out, out_val,
                   dtype_code(logits.dtype), 1 if enable_pdl else 0, stream_ptr(logits))
    return out


class _ptr:
    """Marshal a raw device address through the uniform C ABI (void*)."""

    def __init__(self, v: int):
        self.v = int(v)


# ------------------------------------------------------------------ unified fused all-reduce API (reference flashinfer/comm/allreduce.py:
# AllReduceFusionWorkspace / TRTLLMAllReduceFusionWorkspace / MNNVLAllReduceFusionWorkspace, create_allreduce_fusion_workspace :286,
# allreduce_fusion :460).  The pattern implementation is shared with the TRT-LLM named entry points (trtllm_ar.py).
from .workspace_base import AllReduceFusionWorkspace  # noqa: E402,F401
from .trtllm_ar import AllReduceFusionPattern, QuantizationSFLayout, TRTLLMAllReduceFusionWorkspace, allreduce_fusion  # noqa: E402,F401
from .trtllm_mnnvl_ar import MNNVLAllReduceFusionWorkspace  # noqa: E402,F401


def create_allreduce_fusion_workspace(backend: str = "auto", world_size: Optional[int] = None, rank: Optional[int] = None,
                                      max_token_num: Optional[int] = None, hidden_dim: Optional[int] = None,
                                      dtype: Optional[torch.dtype] = None, gpus_per_node: Optional[int] = None,
                                      comm_backend=None, force_oneshot_support: bool = False,
                                      group: Optional[dist.ProcessGroup] = None) -> AllReduceFusionWorkspace:
    g = group if group is not None else dist.group.WORLD
    world_size = world_size or dist.get_world_size(g)
    rank = dist.get_rank(g) if rank is None else rank
    cls = {"trtllm": TRTLLMAllReduceFusionWorkspace, "mnnvl": MNNVLAllReduceFusionWorkspace}.get(backend, AllReduceFusionWorkspace)
    return cls(world_size, rank, max_token_num, hidden_dim, dtype or torch.bfloat16, g)
