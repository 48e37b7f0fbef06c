"""MoE expert-parallel all-to-all (dispatch / combine) over the NVLink symmetric heap.

Parity: reference flashinfer/comm/trtllm_moe_alltoall.py (MoeAlltoAll :411-743, moe_a2a_* functional API :202-380).
Kernels: csrc/comm/moe_a2a.cu (push dispatch with in-kernel completion handshake, pull combine with fused reduction).

Protocol per MoE layer:  ``recv = a2a.dispatch(topk_ids, [hidden, topk_ids, topk_w], R)`` -> run the local experts on
``recv`` (rows ``>= recv_counts[src]`` are padding; ``invalid_token_expert_id`` marks them) writing the per-row result
into ``a2a.get_combine_payload_tensor_in_workspace(...)`` -> ``out = a2a.combine(payload, R, payload_in_workspace=True)``.
On CPU tensors (gloo groups) the same API runs on ``torch.distributed.all_to_all`` so the host logic is testable.
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.distributed as dist

from .. import jit
from ..utils import dtype_code, stream_ptr
from ._p2p import all_to_all_uneven
from .mapping import Mapping

_MAX_PAYLOADS = 4
_CTRL_BYTES = 3 * 16 * 4


def _pad(x: int, a: int = 128) -> int:
    return (x + a - 1) // a * a


def moe_a2a_get_workspace_size_per_rank(ep_size: int, max_num_tokens: int, hidden_size: int, top_k: int = 8,
                                        dtype: torch.dtype = torch.bfloat16, extra_payload_bytes_per_token: int = 0) -> int:
    """Bytes of symmetric memory per rank: control words + combine region + dispatch payloads
    (hidden + top-k ids + top-k weights + ``extra_payload_bytes_per_token``)."""
    esz = torch.empty(0, dtype=dtype).element_size()
    per_tok = _pad(hidden_size * esz, 16) + 2 * _pad(top_k * 4, 16) + _pad(extra_payload_bytes_per_token, 16)
    return _pad(_CTRL_BYTES, 1024) + _pad(ep_size * max_num_tokens * hidden_size * esz, 1024) + \
        _pad(ep_size * max_num_tokens * per_tok, 1024) + 4 * 1024


class MoeAlltoAll:
    def __init__(self, mapping: Mapping, max_num_tokens: int, top_k: int, num_experts: int,
                 workspace_size_per_rank: Optional[int] = None, hidden_size: Optional[int] = None, mnnvl_config=None,
                 group: Optional[dist.ProcessGroup] = None, dtype: torch.dtype = torch.bfloat16, device=None) -> None:
        self.mapping = mapping
        self.ep_size, self.ep_rank = mapping.moe_ep_size, mapping.moe_ep_rank
        self.max_num_tokens, self.top_k, self.num_experts = max_num_tokens, top_k, num_experts
        if num_experts % self.ep_size:
            raise ValueError("num_experts must be divisible by the EP size")
        self.experts_per_rank = num_experts // self.ep_size
        self.group = group if group is not None else dist.group.WORLD
        if dist.get_world_size(self.group) != self.ep_size:
            raise ValueError("the process group must span exactly the EP ranks")
        self.dtype = dtype
        self._cuda = torch.cuda.is_available() and dist.get_backend(self.group) != "gloo"
        if workspace_size_per_rank is None:
            if hidden_size is None:
                raise ValueError("either workspace_size_per_rank or hidden_size is required")
            workspace_size_per_rank = moe_a2a_get_workspace_size_per_rank(self.ep_size, max_num_tokens, hidden_size, top_k, dtype)
        self.workspace_size_per_rank = workspace_size_per_rank
        self.hidden_size = hidden_size
        self.phase = "idle"
        self.recv_counts: Optional[torch.Tensor] = None
        self._topk_ids: Optional[torch.Tensor] = None
        self._num_tokens = 0
        if self._cuda:
            from .symm import SymmetricHeap

            self.device = torch.device(device or torch.device("cuda", torch.cuda.current_device()))
            self.heap = SymmetricHeap(self.group, workspace_size_per_rank, self.device)
            self.workspace = self.heap.buffer
            self._peer = torch.tensor(self.heap.peer_ptrs, dtype=torch.int64)
            self._state = torch.zeros(4 + 16, dtype=torch.int32, device=self.device)
            self._token_slot = torch.full((max_num_tokens, top_k), -1, dtype=torch.int32, device=self.device)
            self.recv_counts = torch.zeros(self.ep_size, dtype=torch.int32, device=self.device)
            self._mod = jit.load("comm_alltoall")
            self._ctrl_off = 0
            self._combine_off = _pad(_CTRL_BYTES, 1024)
            self._payload_base = None
            self.heap.barrier()

    # ------------------------------------------------------------------ layout
    def _layout(self, payload_bytes: List[int], R: int, hidden_bytes: int):
        combine_bytes = _pad(self.ep_size * self.max_num_tokens * hidden_bytes, 1024) if hidden_bytes else 0
        if self.hidden_size is not None:
            esz = torch.empty(0, dtype=self.dtype).element_size()
            combine_bytes = max(combine_bytes, _pad(self.ep_size * self.max_num_tokens * self.hidden_size * esz, 1024))
        off = self._combine_off + combine_bytes
        offs = []
        for b in payload_bytes:
            offs.append(off)
            off += _pad(self.ep_size * R * b, 1024)
        if off > self.workspace_size_per_rank:
            raise MemoryError(f"MoE all-to-all workspace too small: need {off} bytes, have {self.workspace_size_per_rank}")
        lay = [self._ctrl_off, self._combine_off, R, len(payload_bytes)] + offs + [0] * (_MAX_PAYLOADS - len(offs)) + \
            payload_bytes + [0] * (_MAX_PAYLOADS - len(payload_bytes))
        return torch.tensor(lay, dtype=torch.int64), offs

    # ------------------------------------------------------------------ dispatch
    def dispatch(self, token_selected_experts: torch.Tensor, input_payloads: List[torch.Tensor],
                 runtime_max_tokens_per_rank: int, invalid_token_expert_id: Optional[int] = None,
                 expert_id_payload_index: Optional[int] = None) -> List[torch.Tensor]:
        """Send every token to the ranks that own its selected experts (once per rank).  Returns views
        ``[ep_size, runtime_max_tokens_per_rank, ...]`` of the received payloads; ``self.recv_counts[src]`` rows of
        block ``src`` are valid."""
        if self.phase != "idle":
            raise RuntimeError("dispatch called twice without combine")
        if len(input_payloads) > _MAX_PAYLOADS:
            raise ValueError(f"at most {_MAX_PAYLOADS} payloads")
        R = runtime_max_tokens_per_rank
        if R > self.max_num_tokens:
            raise ValueError("runtime_max_tokens_per_rank exceeds max_num_tokens")
        T, K = token_selected_experts.shape
        ids = token_selected_experts.to(torch.int32).contiguous()
        self._topk_ids, self._num_tokens, self._R = ids, T, R
        if not self._cuda:
            recv = self._dispatch_cpu(ids, input_payloads, R)
        else:
            flat = [p.contiguous().view(T, -1) for p in input_payloads]
            pbytes = [f.shape[1] * f.element_size() for f in flat]
            if any(b % 16 for b in pbytes):
                raise ValueError("payload rows must be multiples of 16 bytes")
            lay, offs = self._layout(pbytes, R, 0)
            self._lay = lay
            ptrs = flat + [None] * (_MAX_PAYLOADS - len(flat))
            self._mod.call("moe_a2a_dispatch", ids, T, K, self.experts_per_rank, self.ep_rank, self.ep_size, self._peer, lay,
                           ptrs[0], ptrs[1], ptrs[2], ptrs[3], self._token_slot, self._state, self.recv_counts, 1,
                           stream_ptr(ids))
            recv = []
            for p, f, off, b in zip(input_payloads, flat, offs, pbytes):
                v = self.workspace[off: off + self.ep_size * R * b].view(p.dtype)
                recv.append(v.view(self.ep_size, R, *p.shape[1:]))
        if invalid_token_expert_id is not None:
            if expert_id_payload_index is None:
                raise ValueError("expert_id_payload_index is required with invalid_token_expert_id")
            self.sanitize_expert_ids(recv[expert_id_payload_index], invalid_token_expert_id)
        self.phase = "dispatched"
        return recv

    def sanitize_expert_ids(self, expert_ids: torch.Tensor, invalid_id: int) -> None:
        ep, R = expert_ids.shape[:2]
        if expert_ids.is_cuda:
            self._mod.call("moe_a2a_sanitize", expert_ids, self.recv_counts, ep, R, expert_ids.shape[2], invalid_id,
                           stream_ptr(expert_ids))
        else:
            rows = torch.arange(R)[None, :, None]
            expert_ids.masked_fill_(rows >= self.recv_counts.view(-1, 1, 1), invalid_id)

    # ------------------------------------------------------------------ combine
    def get_combine_payload_tensor_in_workspace(self, runtime_max_tokens_per_rank: int, hidden_size: int,
                                                dtype: torch.dtype) -> torch.Tensor:
        """``[ep_size, R, hidden]`` view of the symmetric combine region: the expert GEMM can write its result rows
        there directly (zero-copy combine)."""
        R = runtime_max_tokens_per_rank
        if not self._cuda:
            return torch.zeros(self.ep_size, R, hidden_size, dtype=dtype)
        esz = torch.empty(0, dtype=dtype).element_size()
        if R != getattr(self, "_R", R):
            raise ValueError("runtime_max_tokens_per_rank differs from the dispatch call")
        n = self.ep_size * R * hidden_size * esz
        return self.workspace[self._combine_off: self._combine_off + n].view(dtype).view(self.ep_size, R, hidden_size)

    def combine(self, payload: torch.Tensor, runtime_max_tokens_per_rank: int, payload_in_workspace: bool = False) -> torch.Tensor:
        """``payload [ep_size, R, hidden]`` (row ``[src, i]`` = result for the i-th token received from ``src``)
        -> ``[local_num_tokens, hidden]``: the sum over the distinct ranks each token was sent to."""
        if self.phase != "dispatched":
            raise RuntimeError("combine called before dispatch")
        R = runtime_max_tokens_per_rank
        T, K = self._num_tokens, self.top_k
        hidden = payload.shape[-1]
        if not self._cuda:
            out = self._combine_cpu(payload, R)
        else:
            if not payload_in_workspace:
                self.get_combine_payload_tensor_in_workspace(R, hidden, payload.dtype).copy_(payload)
            out = torch.empty(T, hidden, dtype=payload.dtype, device=payload.device)
            lay = self._lay.clone()
            lay[2] = R
            self._mod.call("moe_a2a_combine", self._topk_ids, self._token_slot, out, T, K, hidden, self.experts_per_rank,
                           self.ep_rank, self.ep_size, self._peer, lay, self._state, dtype_code(payload.dtype), 1,
                           stream_ptr(payload))
        self.phase = "idle"
        return out

    # ------------------------------------------------------------------ gloo / CPU path (host-logic tests)
    def _targets(self, ids: torch.Tensor):
        r = torch.where((ids >= 0) & (ids < self.num_experts), ids // self.experts_per_rank, torch.full_like(ids, -1))
        per_rank = []
        for dst in range(self.ep_size):
            per_rank.append(torch.nonzero((r == dst).any(1)).flatten())
        return per_rank

    def _dispatch_cpu(self, ids, payloads, R):
        per_rank = self._targets(ids)
        self._sent_tokens = per_rank
        send_counts = torch.tensor([min(len(p), R) for p in per_rank], dtype=torch.int64)
        recv_counts = torch.zeros_like(send_counts)
        dist.all_to_all_single(recv_counts, send_counts, group=self.group)
        self.recv_counts = recv_counts.to(torch.int32)
        out = []
        for p in payloads:
            sends = [p[idx[:R]].contiguous() for idx in per_rank]
            recvs = [torch.empty(int(c), *p.shape[1:], dtype=p.dtype) for c in recv_counts]
            all_to_all_uneven(recvs, sends, self.group)
            buf = torch.zeros(self.ep_size, R, *p.shape[1:], dtype=p.dtype)
            for s, rcv in enumerate(recvs):
                buf[s, : rcv.shape[0]] = rcv
            out.append(buf)
        return out

    def _combine_cpu(self, payload, R):
        sends = [payload[s, : int(self.recv_counts[s])].contiguous() for s in range(self.ep_size)]
        recvs = [torch.empty(min(len(idx), R), payload.shape[-1], dtype=payload.dtype) for idx in self._sent_tokens]
        all_to_all_uneven(recvs, sends, self.group)
        out = torch.zeros(self._num_tokens, payload.shape[-1], dtype=torch.float32)
        for idx, rcv in zip(self._sent_tokens, recvs):
            out.index_add_(0, idx[:R], rcv.float())
        return out.to(payload.dtype)


# ------------------------------------------------------------------ functional API (reference names)
def moe_a2a_initialize(mapping: Mapping, max_num_tokens: int, top_k: int, num_experts: int, hidden_size: int, **kw) -> MoeAlltoAll:
    return MoeAlltoAll(mapping, max_num_tokens, top_k, num_experts, hidden_size=hidden_size, **kw)


def moe_a2a_dispatch(a2a: MoeAlltoAll, token_selected_experts, input_payloads, runtime_max_tokens_per_rank, **kw):
    return a2a.dispatch(token_selected_experts, input_payloads, runtime_max_tokens_per_rank, **kw)


def moe_a2a_combine(a2a: MoeAlltoAll, payload, runtime_max_tokens_per_rank, payload_in_workspace: bool = False):
    return a2a.combine(payload, runtime_max_tokens_per_rank, payload_in_workspace)


def moe_a2a_sanitize_expert_ids(a2a: MoeAlltoAll, expert_ids, invalid_id: int) -> None:
    a2a.sanitize_expert_ids(expert_ids, invalid_id)


def moe_a2a_wrap_payload_tensor_in_workspace(a2a: MoeAlltoAll, runtime_max_tokens_per_rank: int, hidden_size: int, dtype):
    return a2a.get_combine_payload_tensor_in_workspace(runtime_max_tokens_per_rank, hidden_size, dtype)
