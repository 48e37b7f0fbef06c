"""Minimal serving loop around the op-by-op engines: a page allocator for the paged KV cache and greedy continuous generation
(prefill the prompts of a batch, then decode step by step, growing every request's page list on demand).  It is the glue an inference
server wraps around this library's ``plan`` / ``run`` calls, reduced to what a test or an example needs."""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import torch


class PagedKVAllocator:
    """Free-list page pool with per-request page tables; emits the ``(kv_indptr, kv_indices, kv_last_page_len)`` triple the wrappers plan
    with.  Pages are handed out in a shuffled order (nothing may depend on contiguity)."""

    def __init__(self, num_pages: int, page_size: int, seed: int = 0) -> None:
        self.page_size = page_size
        g = torch.Generator().manual_seed(seed)
        self._free: List[int] = torch.randperm(num_pages, generator=g).tolist()
        self._tables: Dict[int, List[int]] = {}
        self._lens: Dict[int, int] = {}

    @property
    def free_pages(self) -> int:
        return len(self._free)

    def add_request(self, rid: int) -> None:
        if rid in self._tables:
            raise KeyError(f"request {rid} already exists")
        self._tables[rid], self._lens[rid] = [], 0

    def grow(self, rid: int, new_tokens: int) -> None:
        """Reserve room for ``new_tokens`` more tokens of request ``rid`` (raises MemoryError when the pool is exhausted)."""
        want = self._lens[rid] + new_tokens
        need = -(-want // self.page_size) - len(self._tables[rid])
        if need > len(self._free):
            raise MemoryError(f"KV pool exhausted: request {rid} needs {need} pages, {len(self._free)} free")
        for _ in range(need):
            self._tables[rid].append(self._free.pop())
        self._lens[rid] = want

    def truncate(self, rid: int, length: int) -> None:
        """Forget the tokens beyond ``length`` (rejected speculative tokens); their pages stay reserved for the request."""
        if length > self._lens[rid]:
            raise ValueError("truncate cannot grow a request")
        self._lens[rid] = length

    def release(self, rid: int) -> None:
        self._free.extend(self._tables.pop(rid))
        self._lens.pop(rid)

    def length(self, rid: int) -> int:
        return self._lens[rid]

    def tables(self, rids: Sequence[int]) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        counts = [len(self._tables[r]) for r in rids]
        indptr = torch.tensor([0] + list(torch.tensor(counts).cumsum(0).tolist()) if counts else [0], dtype=torch.int32)
        indices = torch.tensor([p for r in rids for p in self._tables[r]], dtype=torch.int32)
        last = torch.tensor([(self._lens[r] - 1) % self.page_size + 1 if self._lens[r] else 0 for r in rids], dtype=torch.int32)
        return indptr, indices, last


def _num_pages(engine) -> int:
    """Pages of an engine's paged cache (GQA engines: ``k_cache``, the MLA engine: ``ckv_cache``)."""
    layer = engine.layers[0]
    return (layer["k_cache"] if "k_cache" in layer else layer["ckv_cache"]).shape[0]


def generate(engine, prompts: Sequence[Sequence[int]], max_new_tokens: int, allocator: PagedKVAllocator = None, eos_token: int = -1) -> List[List[int]]:
    """Greedy generation with an engine that has ``prefill`` / ``plan`` / ``step`` over a paged cache
    (:class:`~flashinfer_b200.models.transformer.TransformerDecodeEngine`, :class:`~flashinfer_b200.models.deepseek.DeepSeekDecodeEngine`): one batched prefill of all prompts,
    then batched decode steps; a request that emits ``eos_token`` leaves the batch and returns its pages.  Returns the generated tokens
    (without the prompts)."""
    alloc = allocator or PagedKVAllocator(_num_pages(engine), engine.page_size)
    rids = list(range(len(prompts)))
    for r, p in zip(rids, prompts):
        alloc.add_request(r)
        alloc.grow(r, len(p))
    qo = torch.tensor([0] + list(torch.tensor([len(p) for p in prompts]).cumsum(0).tolist()), dtype=torch.int32)
    flat = torch.tensor([t for p in prompts for t in p], dtype=torch.int64)
    nxt = engine.prefill(flat, qo, *alloc.tables(rids)).tolist()
    out: List[List[int]] = [[] for _ in prompts]
    active = list(rids)
    for _ in range(max_new_tokens):
        emitted = dict(zip(active, nxt))
        for r, t in emitted.items():
            out[r].append(int(t))
        active = [r for r in active if emitted[r] != eos_token]
        for r in list(emitted):
            if emitted[r] == eos_token:
                alloc.release(r)
        if not active or len(out[active[0]]) >= max_new_tokens:
            break
        for r in active:
            alloc.grow(r, 1)
        engine.plan(*alloc.tables(active))
        engine.tokens.copy_(torch.tensor([emitted[r] for r in active], dtype=torch.int64))
        nxt = engine.step().tolist()
    return out


def _probs(logits: torch.Tensor, temperature: float) -> torch.Tensor:
    """Sampling distribution of a logits row: softmax at ``temperature``, the one-hot argmax at temperature 0 (greedy)."""
    if temperature <= 0:
        return torch.nn.functional.one_hot(logits.argmax(-1), logits.shape[-1]).float()
    return torch.softmax(logits.float() / temperature, -1)


def speculative_generate(target, draft, prompts: Sequence[Sequence[int]], max_new_tokens: int, num_draft_tokens: int = 3,
                         temperature: float = 0.0, generator=None) -> Tuple[List[List[int]], float]:
    """Speculative decoding with two :class:`~flashinfer_b200.models.transformer.TransformerDecodeEngine` s (same vocabulary): per
    round the draft proposes ``num_draft_tokens`` tokens with single-token decode steps, the target scores the pending token and all
    proposals in ONE multi-token pass over its paged cache, and ``sampling.chain_speculative_sampling`` accepts a prefix and emits one
    corrected / bonus token.  Rejected tokens are dropped from both caches by truncating the requests' lengths.  Returns the generated
    tokens and the mean number of tokens emitted per round.  At ``temperature`` 0 the output equals greedy decoding of the target,
    whatever the draft proposes."""
    from ..sampling import chain_speculative_sampling

    k, b = num_draft_tokens, len(prompts)
    pools = {e: PagedKVAllocator(_num_pages(e), e.page_size) for e in (target, draft)}
    rids = list(range(b))
    qo = torch.tensor([0] + list(torch.tensor([len(p) for p in prompts]).cumsum(0).tolist()), dtype=torch.int32)
    flat = torch.tensor([t for p in prompts for t in p], dtype=torch.int64)
    pending = None
    for e in (draft, target):                                                 # both caches hold the prompts; the target picks the first token
        for r, p in zip(rids, prompts):
            pools[e].add_request(r)
            pools[e].grow(r, len(p))
        e.prefill(flat, qo, *pools[e].tables(rids))
        pending = torch.multinomial(_probs(e.logits, temperature), 1, generator=generator)[:, 0]
    out: List[List[int]] = [[int(t)] for t in pending]
    rounds = emitted_total = 0
    while min(len(o) for o in out) < max_new_tokens:
        base = [pools[target].length(r) for r in rids]                        # tokens in the caches; ``pending`` is not among them yet
        # ---- draft: k proposals (+ one more step that only writes the last proposal into the draft cache)
        cur, d_ids, d_probs = pending.clone(), [], []
        for i in range(k + 1):
            for r in rids:
                pools[draft].grow(r, 1)
            draft.plan(*pools[draft].tables(rids))
            draft.tokens.copy_(cur)
            draft.step()
            if i < k:
                pr = _probs(draft.logits, temperature)
                cur = torch.multinomial(pr, 1, generator=generator)[:, 0]
                d_ids.append(cur.clone())
                d_probs.append(pr)
        d_ids, d_probs = torch.stack(d_ids, 1), torch.stack(d_probs, 1)       # [b, k], [b, k, vocab]
        # ---- target: score [pending, d_1 .. d_k] in one pass
        for r in rids:
            pools[target].grow(r, k + 1)
        seq = torch.cat([pending[:, None], d_ids], 1).reshape(-1)
        target.prefill(seq, torch.arange(0, (b + 1) * (k + 1), k + 1, dtype=torch.int32), *pools[target].tables(rids), all_logits=True)
        t_probs = _probs(target.logits, temperature).view(b, k + 1, -1)
        tokens = chain_speculative_sampling(d_probs.float(), d_ids.int(), t_probs, generator=generator)[0]   # [b, k + 1], -1 padded
        n_emit = (tokens >= 0).sum(1)
        for i, r in enumerate(rids):
            new = tokens[i, : int(n_emit[i])].tolist()
            out[i].extend(int(t) for t in new)
            keep = base[i] + int(n_emit[i])                                   # pending + the accepted proposals stay cached
            pools[target].truncate(r, keep)
            pools[draft].truncate(r, keep)
        pending = torch.stack([tokens[i, int(n_emit[i]) - 1] for i in range(b)]).long()
        rounds += 1
        emitted_total += int(n_emit.sum())
    return [o[:max_new_tokens] for o in out], emitted_total / max(rounds * b, 1)
