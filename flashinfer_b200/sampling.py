"""Sorting-free sampling.  Parity: reference flashinfer/sampling.py:737-1957.

CUDA tensors run csrc/elementwise/sampling.cu (dual-pivot rejection sampling, Philox RNG keyed by
(seed, offset) so the ops are CUDA-graph friendly); CPU tensors use a plain torch implementation of the
same distributions (multinomial over the filtered / renormalised probabilities).
"""
from __future__ import annotations

from typing import Optional, Tuple, Union

import torch

from . import jit
from .utils import stream_ptr

_MAX_ROUNDS = 64


def get_seed_and_offset(increment: int, generator: Optional[torch.Generator] = None, device=None) -> Tuple[int, int]:
    """Draw a (seed, offset) pair from a torch generator and advance it (reference sampling.py:47)."""
    if generator is None:
        generator = torch.cuda.default_generators[torch.cuda.current_device()] if torch.cuda.is_available() else torch.default_generator
    state = generator.get_state()
    try:
        seed = int(generator.initial_seed())
        off_t = state.view(torch.int64)
        offset = int(off_t[-1]) if off_t.numel() >= 2 else 0
        off_t[-1] = offset + (increment + 3) // 4 * 4
        generator.set_state(state)
    except Exception:  # noqa: BLE001
        seed, offset = int(torch.randint(0, 2**31 - 1, (1,)).item()), 0
    return seed & 0x7FFFFFFFFFFFFFFF, offset & 0x7FFFFFFFFFFFFFFF


def _seed_offset(generator, seed, offset, increment):
    if seed is not None:
        s = int(seed.item()) if isinstance(seed, torch.Tensor) else int(seed)
        o = int(offset.item()) if isinstance(offset, torch.Tensor) else int(offset or 0)
        return s, o
    return get_seed_and_offset(increment, generator)


def _split(x, dtype):
    """tensor-or-scalar -> (tensor_or_None, scalar)."""
    if isinstance(x, torch.Tensor):
        return x.to(dtype).contiguous(), 0
    return None, x


def softmax(logits: torch.Tensor, temperature: Optional[Union[torch.Tensor, float]] = None,
            enable_pdl: Optional[bool] = None) -> torch.Tensor:
    """Safe softmax with temperature scaling over the last dim of ``[batch, vocab]`` logits."""
    if not logits.is_cuda:
        t = temperature if temperature is not None else 1.0
        t = t[:, None] if isinstance(t, torch.Tensor) else t
        return torch.softmax(logits.float() / t, dim=-1)
    x = logits.float().contiguous()
    probs = torch.empty_like(x)
    t_arr, t_val = _split(temperature if temperature is not None else 1.0, torch.float32)
    jit.load("sampling").call("softmax_run", x, probs, t_arr, float(t_val), x.shape[0], x.shape[1], stream_ptr(x))
    return probs


def _sample(probs, mode, indices=None, top_k=0, top_p=1.0, min_p=0.0, generator=None, seed=None, offset=None,
            return_valid=False):
    batch = indices.numel() if indices is not None else probs.shape[0]
    vocab = probs.shape[-1]
    if not probs.is_cuda:
        return _sample_cpu(probs, mode, indices, top_k, top_p, min_p, generator, return_valid)
    p = probs.float().contiguous()
    out = torch.empty(batch, dtype=torch.int32, device=p.device)
    valid = torch.empty(batch, dtype=torch.bool, device=p.device) if return_valid else None
    k_arr, k_val = _split(top_k, torch.int32)
    p_arr, p_val = _split(top_p, torch.float32)
    m_arr, m_val = _split(min_p, torch.float32)
    s, o = _seed_offset(generator, seed, offset, _MAX_ROUNDS * 4)
    jit.load("sampling").call(
        "sampling_run", p, out, indices.int() if indices is not None else None, p_arr, float(p_val), k_arr, int(k_val),
        m_arr, float(m_val), batch, vocab, mode, s, o, _MAX_ROUNDS, valid, stream_ptr(p),
    )
    return (out, valid) if return_valid else out


def _sample_cpu(probs, mode, indices, top_k, top_p, min_p, generator, return_valid):
    p = probs.float()
    if indices is not None:
        p = p[indices.long()]
    if mode in (1, 4):
        p = top_k_renorm_probs(p, top_k)
    if mode in (2, 4):
        p = top_p_renorm_probs(p, top_p)
    if mode == 3:
        thr = p.max(-1, keepdim=True).values * (min_p[:, None] if isinstance(min_p, torch.Tensor) else min_p)
        p = torch.where(p >= thr, p, torch.zeros_like(p))
    out = torch.multinomial(p / p.sum(-1, keepdim=True), 1, generator=generator)[:, 0].int()
    if return_valid:
        return out, torch.ones_like(out, dtype=torch.bool)
    return out


def _check_nan(probs, check_nan: bool) -> None:
    """``check_nan=True`` of the reference's samplers: refuse NaN probabilities (one device reduction + a host sync, opt-in)."""
    if check_nan and torch.isnan(probs).any():
        raise ValueError("Input probs contains NaN.")


def sampling_from_probs(probs, indices=None, deterministic=True, generator=None, check_nan=False, seed=None,
                        offset=None, return_valid=False):
    """Category sampling from ``probs [batch, vocab]`` (inverse-CDF)."""
    _check_nan(probs, check_nan)
    return _sample(probs, 0, indices, generator=generator, seed=seed, offset=offset, return_valid=return_valid)


def sampling_from_logits(logits, indices=None, deterministic=True, generator=None, check_nan=False, seed=None,
                         offset=None, return_valid=False):
    """Category sampling from logits (softmax fused upstream)."""
    return sampling_from_probs(softmax(logits), indices, deterministic, generator, check_nan, seed, offset, return_valid)


def top_p_sampling_from_probs(probs, top_p, indices=None, deterministic=True, generator=None, check_nan=False,
                              seed=None, offset=None, return_valid=False):
    _check_nan(probs, check_nan)
    return _sample(probs, 2, indices, top_p=top_p, generator=generator, seed=seed, offset=offset,
                   return_valid=return_valid)


def top_k_sampling_from_probs(probs, top_k, indices=None, deterministic=True, generator=None, check_nan=False,
                              seed=None, offset=None, return_valid=False):
    _check_nan(probs, check_nan)
    return _sample(probs, 1, indices, top_k=top_k, generator=generator, seed=seed, offset=offset,
                   return_valid=return_valid)


def min_p_sampling_from_probs(probs, min_p, indices=None, deterministic=True, generator=None, check_nan=False,
                              seed=None, offset=None, return_valid=False):
    _check_nan(probs, check_nan)
    return _sample(probs, 3, indices, min_p=min_p, generator=generator, seed=seed, offset=offset,
                   return_valid=return_valid)


def top_k_top_p_sampling_from_probs(probs, top_k, top_p, indices=None, filter_apply_order="top_k_first",
                                    deterministic=True, generator=None, check_nan=False, seed=None, offset=None,
                                    return_valid=False):
    if filter_apply_order == "top_k_first":
        renorm = top_k_renorm_probs(probs, top_k)
        return top_p_sampling_from_probs(renorm, top_p, indices, deterministic, generator, check_nan, seed, offset,
                                         return_valid)
    if filter_apply_order == "joint":
        _check_nan(probs, check_nan)
        return _sample(probs, 4, indices, top_k=top_k, top_p=top_p, generator=generator, seed=seed, offset=offset,
                       return_valid=return_valid)
    raise ValueError(f"Invalid filter_apply_order: {filter_apply_order}")


def top_k_top_p_sampling_from_logits(logits, top_k, top_p, indices=None, filter_apply_order="top_k_first",
                                     deterministic=True, generator=None, check_nan=False, seed=None, offset=None,
                                     return_valid=False):
    if filter_apply_order == "top_k_first":
        masked = top_k_mask_logits(logits, top_k)
        probs = softmax(masked)
        return top_p_sampling_from_probs(probs, top_p, indices, deterministic, generator, check_nan, seed, offset,
                                         return_valid)
    if filter_apply_order == "joint":
        return top_k_top_p_sampling_from_probs(softmax(logits), top_k, top_p, indices, "joint", deterministic,
                                               generator, check_nan, seed, offset, return_valid)
    raise ValueError(f"Invalid filter_apply_order: {filter_apply_order}")


def _renorm(x, mode, top_p=1.0, top_k=0):
    if not x.is_cuda:
        xf = x.float()
        if mode == 0:
            sp, si = xf.sort(-1, descending=True)
            cum = sp.cumsum(-1)
            tp = top_p[:, None] if isinstance(top_p, torch.Tensor) else top_p
            keep_sorted = (cum - sp) < tp
            thr = torch.where(keep_sorted, sp, torch.full_like(sp, float("inf"))).min(-1, keepdim=True).values
            kept = torch.where(xf >= thr, xf, torch.zeros_like(xf))
            return kept / kept.sum(-1, keepdim=True)
        k = top_k if isinstance(top_k, torch.Tensor) else torch.full((xf.shape[0],), int(top_k))
        k = k.clamp(min=1, max=xf.shape[-1]).long()
        kth = xf.sort(-1, descending=True).values.gather(-1, (k - 1)[:, None])
        if mode == 1:
            kept = torch.where(xf >= kth, xf, torch.zeros_like(xf))
            return kept / kept.sum(-1, keepdim=True)
        return torch.where(xf >= kth, xf, torch.full_like(xf, float("-inf")))
    xf = x.float().contiguous()
    out = torch.empty_like(xf)
    p_arr, p_val = _split(top_p, torch.float32)
    k_arr, k_val = _split(top_k, torch.int32)
    jit.load("sampling").call("renorm_run", xf, out, p_arr, float(p_val), k_arr, int(k_val), xf.shape[0], xf.shape[1],
                              mode, stream_ptr(xf))
    return out


def top_p_renorm_probs(probs, top_p, is_deterministic: bool = False):
    """Keep the smallest set of entries whose mass reaches ``top_p`` and renormalise.  ``is_deterministic`` selects integer
    histograms in the reference's radix search; the threshold search here has a fixed reduction order - always deterministic."""
    return _renorm(probs, 0, top_p=top_p)


def top_k_renorm_probs(probs, top_k):
    """Keep the ``top_k`` largest entries (ties kept) and renormalise."""
    return _renorm(probs, 1, top_k=top_k)


def top_k_mask_logits(logits, top_k):
    """Set everything but the ``top_k`` largest logits to ``-inf``."""
    return _renorm(logits, 2, top_k=top_k)


def chain_speculative_sampling(draft_probs, draft_token_ids, target_probs, maybe_output_accepted_token_num=None,
                               maybe_output_emitted_draft_token_num=None, deterministic=True, generator=None,
                               seed=None, offset=None):
    """Speculative-decoding verification (Leviathan et al.): accept draft token ``i`` with probability
    ``min(1, q/p)``, on the first rejection resample from ``relu(q - p)``; returns ``[batch, n+1]`` token ids
    padded with ``-1`` (plus the two optional counters updated in place)."""
    b, n, v = draft_probs.shape
    out = torch.empty(b, n + 1, dtype=torch.int32, device=draft_probs.device)
    acc = maybe_output_accepted_token_num
    emi = maybe_output_emitted_draft_token_num
    if acc is None:
        acc = torch.zeros(b, dtype=torch.int32, device=draft_probs.device)
    if emi is None:
        emi = torch.zeros(b, dtype=torch.int32, device=draft_probs.device)
    if not draft_probs.is_cuda:
        for i in range(b):
            emitted, pos, rejected = 0, 0, False
            for pos in range(n):
                tok = int(draft_token_ids[i, pos])
                q, p = float(target_probs[i, pos, tok]), float(draft_probs[i, pos, tok])
                if float(torch.rand(1, generator=generator)) * p < q:
                    out[i, pos] = tok
                    emitted += 1
                else:
                    rejected = True
                    break
            else:
                pos = n
            dist_ = (target_probs[i, pos] - draft_probs[i, pos]).clamp(min=0) if rejected else target_probs[i, pos]
            out[i, pos] = int(torch.multinomial(dist_ / dist_.sum(), 1, generator=generator))
            out[i, pos + 1 :] = -1
            emi[i] += emitted
            acc[i] += emitted
        return out, acc, emi
    s, o = _seed_offset(generator, seed, offset, (n + 2) * 2)
    jit.load("sampling").call(
        "chain_speculative_sampling_run", draft_probs.float().contiguous(), draft_token_ids.int().contiguous(),
        target_probs.float().contiguous(), out, acc, emi, b, n, v, 1 if deterministic else 0, s, o,
        stream_ptr(draft_probs),
    )
    return out, acc, emi


from . import jit as _jit_acc  # noqa: E402

get_sampling_module = _jit_acc.module_accessor("sampling")
