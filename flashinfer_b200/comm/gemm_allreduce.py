"""Fused GEMM + all-reduce (row-parallel linear of tensor parallelism) in ONE kernel per rank.

Parity: reference flashinfer/cute_dsl/gemm_allreduce_two_shot.py (G17) and its test
tests/gemm/test_cute_dsl_gemm_allreduce_two_shot.py.  Kernel: csrc/gemm/gemm_allreduce_sm100.cu.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist

from .. import jit
from ..utils import dtype_code, stream_ptr
from .allreduce import _ptr

_MAX_TILES = 8192
_MAX_CTAS = 148


class GemmAllReduce:
    """``out = all_reduce(a @ w.T)`` with ``a [M, K_local]``, ``w [N, K_local]`` (both K-major, bf16 / fp16).

    one-shot (default for small M): every rank pulls the in-switch sum of every tile -> ``out`` is an ordinary tensor.
    two-shot: tile ``t`` is reduced by rank ``t % world`` and multicast-stored to all ranks (less switch traffic for large M).
    """

    def __init__(self, group: Optional[dist.ProcessGroup], max_m: int, n: int, dtype: torch.dtype = torch.bfloat16,
                 use_nvls: bool = True) -> None:
        from .symm import SymmetricHeap

        self.group = group if group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
        self.max_m, self.n, self.dtype = max_m, n, dtype
        esz = torch.empty(0, dtype=dtype).element_size()
        cbytes = (max_m * n * esz + 1023) // 1024 * 1024
        fbytes = _MAX_TILES * 4
        dbytes = _MAX_CTAS * 16 * 4
        self.heap = SymmetricHeap(self.group, 3 * cbytes + fbytes + 2 * dbytes + 16384)
        self._stage = []
        for _ in range(2):
            v, off = self.heap.alloc(cbytes)
            self._stage.append((v.view(dtype)[: max_m * n].view(max_m, n), off, self.heap.peer_ptr_table(off)))
        v, self._out_off = self.heap.alloc(cbytes)
        self._out = v.view(dtype)[: max_m * n].view(max_m, n)
        self._out_tab = self.heap.peer_ptr_table(self._out_off)
        _, self._flag_off = self.heap.alloc(fbytes)
        self._flag_tab = self.heap.peer_ptr_table(self._flag_off)
        _, self._done_off = self.heap.alloc(dbytes)
        self._done_tab = self.heap.peer_ptr_table(self._done_off)
        _, self._rs_sig_off = self.heap.alloc(dbytes)  # epoch slots of the chunk-pipelined reduce-scatter (rs_pull_rows)
        self._rs_sig_tab = self.heap.peer_ptr_table(self._rs_sig_off)
        dev = self.heap.device
        self._rs_epoch = torch.zeros(_MAX_CTAS, dtype=torch.int32, device=dev)
        self._side = None
        self._expect = torch.zeros(_MAX_TILES, dtype=torch.int32, device=dev)
        self._done_epoch = torch.zeros(_MAX_CTAS, dtype=torch.int32, device=dev)
        self.use_nvls = bool(use_nvls and self.heap.mc_ptr)
        self._turn = 0
        self._mod = jit.load("gemm_comm_sm100")
        self.heap.barrier()

    def __call__(self, a: torch.Tensor, w: torch.Tensor, out: Optional[torch.Tensor] = None, two_shot: Optional[bool] = None,
                 bn: int = 0) -> torch.Tensor:
        M, K = a.shape
        N = w.shape[0]
        if N != self.n or M > self.max_m or a.dtype != self.dtype or w.dtype != self.dtype:
            raise ValueError("GemmAllReduce: shape / dtype does not match the communicator")
        if a.stride(1) != 1 or w.stride(1) != 1:
            a, w = a.contiguous(), w.contiguous()
        if two_shot is None:
            two_shot = M * N * a.element_size() > (4 << 20)
        self._turn ^= 1
        stage, soff, stab = self._stage[self._turn]
        if two_shot:
            target = self._out[:M]
        else:
            target = out if out is not None else torch.empty(M, N, dtype=a.dtype, device=a.device)
        mc = self.heap.mc if self.use_nvls else (lambda off: 0)
        self._mod.call("gemm_allreduce_nt", a, w, stage, target, M, N, K, a.stride(0), w.stride(0), N, dtype_code(a.dtype), stab,
                       self._flag_tab, self._out_tab if two_shot else None, self._done_tab, _ptr(mc(soff)), _ptr(mc(self._flag_off)),
                       _ptr(mc(self._out_off) if two_shot else 0), self._expect, self._done_epoch, self.rank, self.world,
                       1 if two_shot else 0, _MAX_TILES, bn, 0, 0, None, None, 1, stream_ptr(a))
        if two_shot:
            if out is not None:
                out.copy_(target)
                return out
            return target.clone()
        return target


    def reduce_scatter(self, a: torch.Tensor, w: torch.Tensor, residual: Optional[torch.Tensor] = None,
                       rms_weight: Optional[torch.Tensor] = None, eps: float = 1e-6, out: Optional[torch.Tensor] = None,
                       bn: int = 0, pipelined: Optional[int] = None):
        """``GEMM -> reduce-scatter (-> + residual -> RMSNorm)`` of sequence-parallel tensor parallelism, one GEMM kernel per
        rank: rank ``r`` ends up with rows ``[r * M / world, (r + 1) * M / world)`` of ``sum_ranks(a @ w.T)``.

        The GEMM epilogue stores tiles to the symmetric staging buffer; the same kernel's reduce warps pull ONLY the rows
        this rank owns through the switch (``multimem.ld_reduce``), add the ``residual`` shard, write the local shard and
        accumulate each row's sum of squares, so the RMSNorm that follows is a single scale pass (``rs_rmsnorm``).

        Returns ``shard`` (``[M / world, N]``; it is ``sum + residual`` when ``residual`` is given), or
        ``(normed, shard)`` when ``rms_weight`` is given.  BASELINE config "Llama-3-70B TP=8 GEMM->reduce-scatter +
        add-RMSNorm fusion"; reference analogue: flashinfer/comm/trtllm_ar.py allreduce-fusion patterns applied per shard."""
        M, K = a.shape
        N = w.shape[0]
        if N != self.n or M > self.max_m or a.dtype != self.dtype or w.dtype != self.dtype:
            raise ValueError("GemmAllReduce.reduce_scatter: shape / dtype does not match the communicator")
        if M % self.world:
            raise ValueError("reduce_scatter: M must be divisible by the group size")
        rpr = M // self.world
        if a.stride(1) != 1 or w.stride(1) != 1:
            a, w = a.contiguous(), w.contiguous()
        if residual is not None:
            if residual.shape != (rpr, N) or residual.dtype != a.dtype:
                raise ValueError("reduce_scatter: residual must be the local [M / world, N] shard")
            residual = residual.contiguous()
        shard = out if out is not None else torch.empty(rpr, N, dtype=a.dtype, device=a.device)
        if shard.stride(1) != 1 or shard.stride(0) != N:
            raise ValueError("reduce_scatter: out must be a contiguous [M / world, N] tensor")
        sumsq = None
        if rms_weight is not None:
            if getattr(self, "_sumsq", None) is None or self._sumsq.numel() < rpr:
                self._sumsq = torch.zeros(max(rpr, self.max_m // self.world + 1), dtype=torch.float32, device=a.device)
            sumsq = self._sumsq
        self._turn ^= 1
        stage, soff, stab = self._stage[self._turn]
        mc = self.heap.mc if self.use_nvls else (lambda off: 0)
        if pipelined is None:
            chunks = self._tuned_chunks(a, w, stage, soff, stab, shard, residual, sumsq, rpr, bn)
        else:
            chunks = int(pipelined) if pipelined else 0
        self._run_rs(chunks, a, w, stage, soff, stab, shard, residual, sumsq, rpr, bn)
        if rms_weight is None:
            return shard
        normed = torch.empty_like(shard)
        self._mod.call("rs_rmsnorm", shard, normed, rms_weight.to(a.dtype).contiguous(), sumsq, rpr, N, N, N, float(eps),
                       dtype_code(a.dtype), 1, stream_ptr(a))
        return normed, shard


    # ------------------------------------------------------------------ prefill sizes: chunk-pipelined GEMM / in-switch pull
    def _run_rs(self, chunks, a, w, stage, soff, stab, shard, residual, sumsq, rpr, bn) -> None:
        M, K = a.shape
        N = w.shape[0]
        if chunks:
            self._reduce_scatter_pipelined(a, w, stage, soff, stab, shard, residual, sumsq, rpr, chunks)
            return
        mc = self.heap.mc if self.use_nvls else (lambda off: 0)
        self._mod.call("gemm_allreduce_nt", a, w, stage, shard, M, N, K, a.stride(0), w.stride(0), N, dtype_code(a.dtype), stab,
                       self._flag_tab, None, self._done_tab, _ptr(mc(soff)), _ptr(mc(self._flag_off)), _ptr(0), self._expect,
                       self._done_epoch, self.rank, self.world, 2, _MAX_TILES, bn, rpr, N, residual, sumsq, 1, stream_ptr(a))

    def _chunk_candidates(self, M: int, N: int, K: int):
        """Schedules of GEMM -> reduce-scatter for this shape: 0 = the one-kernel version (tile-by-tile overlap inside one
        launch: wins at decode / small-prefill sizes), 1 = one full-size GEMM followed by one full-machine in-switch pull,
        c >= 2 = c pipeline chunks (the pull of chunk i on a side stream under the GEMM of chunk i + 1)."""
        rpr = M // self.world
        if M < 2048 or N % 256 or rpr % 128:
            return [0]
        # (without an NVLS multicast mapping of the staging buffer - e.g. a heap too large for the multicast object - the pull
        # kernel reads the `world` partials with direct peer loads: same bytes over NVLink, still far ahead of the one-kernel
        # version whose 4 reduce warps cannot keep a prefill-sized pull in flight: 4.4 ms at TP 8, M = 32768)
        cands = [c for c in (4, 2, 8, 1) if rpr % (c * 128) == 0 and rpr // c >= 256]
        cands.append(0)
        return cands

    def _pipeline_chunks(self, M: int, N: int, K: int) -> int:
        """Default schedule without tuning (first admissible candidate)."""
        return self._chunk_candidates(M, N, K)[0]

    def _tuned_chunks(self, a, w, stage, soff, stab, shard, residual, sumsq, rpr, bn) -> int:
        """Schedule for this (M, N, K): the candidates of :meth:`_chunk_candidates` are timed once per shape on the live
        tensors (device events, median of 3, max over ranks through one all-reduce so that every rank takes the same
        decision) and the winner is cached - the communication / compute balance depends on the group size and on K, and a
        wrong static choice costs integer factors at TP = 8 (profiles/tp_gemm_rs.md).  ``FIB200_RS_CHUNKS`` pins a schedule."""
        import os

        M, K = a.shape
        N = w.shape[0]
        key = (M, N, K, residual is not None, sumsq is not None)
        tuned = self.__dict__.setdefault("_rs_tuned", {})
        if key in tuned:
            return tuned[key]
        env = os.environ.get("FIB200_RS_CHUNKS")
        cands = self._chunk_candidates(M, N, K)
        if env is not None:
            tuned[key] = int(env)
            return tuned[key]
        if len(cands) == 1 or torch.cuda.is_current_stream_capturing():
            return cands[0]
        times = []
        for c in cands:
            ts = []
            for it in range(4):
                if sumsq is not None:
                    sumsq.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                self._run_rs(c, a, w, stage, soff, stab, shard, residual, sumsq, rpr, bn)
                e1.record()
                torch.cuda.synchronize()
                if it:
                    ts.append(e0.elapsed_time(e1))
            times.append(sorted(ts)[len(ts) // 2])
        t = torch.tensor(times, device=a.device, dtype=torch.float32)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        best = int(torch.argmin(t).item())
        tuned[key] = cands[best]
        self.__dict__.setdefault("_rs_tuning_log", {})[key] = {int(c): round(float(x), 4) for c, x in zip(cands, t.tolist())}
        if sumsq is not None:
            sumsq.zero_()
        return tuned[key]

    def _reduce_scatter_pipelined(self, a, w, stage, soff, stab, shard, residual, sumsq, rpr: int, chunks: int) -> None:
        from ..gemm.dense import linear

        M, K = a.shape
        N = w.shape[0]
        mc = self.heap.mc(soff) if self.use_nvls else 0
        main = torch.cuda.current_stream(a.device)
        if self._side is None:
            self._side = torch.cuda.Stream(a.device)
        side = self._side
        m_c = rpr // chunks
        if chunks == 1:  # one full-size GEMM, then one full-machine pull, same stream
            linear(a, w, out=stage[:M])
            lo = self.rank * rpr
            self._mod.call("rs_pull_rows", _ptr(mc), stab, self._rs_sig_tab, self._rs_epoch, self.rank, self.world, N, lo, rpr, N,
                           shard, N, residual, N, sumsq, _MAX_CTAS, dtype_code(a.dtype), main.cuda_stream)
            return
        evs = []
        for c in range(chunks):
            for r in range(self.world):  # chunk c of EVERY rank's row range: all ranks pull concurrently, links stay evenly loaded
                lo = r * rpr + c * m_c
                linear(a[lo: lo + m_c], w, out=stage[lo: lo + m_c])
            ev = torch.cuda.Event()
            ev.record(main)
            evs.append(ev)
            side.wait_event(ev)
            lo = self.rank * rpr + c * m_c
            self._mod.call("rs_pull_rows", _ptr(mc), stab, self._rs_sig_tab, self._rs_epoch, self.rank, self.world, N, lo, m_c, N,
                           shard[c * m_c: (c + 1) * m_c], N, residual[c * m_c: (c + 1) * m_c] if residual is not None else None,
                           N, sumsq[c * m_c:] if sumsq is not None else None, _MAX_CTAS, dtype_code(a.dtype), side.cuda_stream)
        main.wait_stream(side)
        for t in (a, w, shard, stage):
            t.record_stream(side)
        if residual is not None:
            residual.record_stream(side)


_CACHE: dict = {}


def gemm_allreduce(a: torch.Tensor, w: torch.Tensor, group: Optional[dist.ProcessGroup] = None, out: Optional[torch.Tensor] = None,
                   two_shot: Optional[bool] = None) -> torch.Tensor:
    """Functional form with a per-(group, N, dtype) cached communicator sized for ``max(M, 8192)`` rows."""
    g = group if group is not None else dist.group.WORLD
    key = (id(g), w.shape[0], a.dtype)
    comm = _CACHE.get(key)
    if comm is None or comm.max_m < a.shape[0]:
        comm = GemmAllReduce(g, max(a.shape[0], 8192), w.shape[0], a.dtype)
        _CACHE[key] = comm
    return comm(a, w, out, two_shot)


def gemm_reduce_scatter(a: torch.Tensor, w: torch.Tensor, group: Optional[dist.ProcessGroup] = None,
                        residual: Optional[torch.Tensor] = None, rms_weight: Optional[torch.Tensor] = None, eps: float = 1e-6,
                        out: Optional[torch.Tensor] = None):
    """Functional form of :meth:`GemmAllReduce.reduce_scatter` (cached communicator)."""
    g = group if group is not None else dist.group.WORLD
    key = (id(g), w.shape[0], a.dtype)
    comm = _CACHE.get(key)
    if comm is None or comm.max_m < a.shape[0]:
        comm = GemmAllReduce(g, max(a.shape[0], 8192), w.shape[0], a.dtype)
        _CACHE[key] = comm
    return comm.reduce_scatter(a, w, residual, rms_weight, eps, out)
