"""Intra-kernel profiler host side.  Parity: reference flashinfer/profiler/__init__.py:28-101 (decode tags ->
Perfetto).  The device side is ``csrc/include/fib200/profiler.cuh``; here we allocate the buffer, decode it and write a
Chrome/Perfetto ``traceEvents`` JSON (no external dependency)."""
from __future__ import annotations

import json
from enum import Enum
from typing import Dict, List, Optional, Sequence

import torch

START, END, INSTANT = 0, 1, 2


class EventType(Enum):
    """Low two bits of a tag (reference profiler/__init__.py EventType)."""
    kBegin = 0
    kEnd = 1
    kInstant = 2


def decode_tag(tag: int, num_blocks: int = 0, num_groups: int = 0):
    """``(block_idx, group_idx, event_idx, event_type, sm_id)`` of one 32-bit tag.

    The reference packs block / group / SM into the tag itself (bits 12-23 and 24-31); tags written by ``profiler.cuh``
    carry only ``event << 2 | type`` - block, group and SM id live in the header word of each (block, group) slot - so
    both layouts decode here: fields that a fib200 tag does not hold come back as 0."""
    event_type = tag & 0x3
    if tag >> 12:                                   # reference layout
        bg = (tag >> 12) & 0xFFF
        return (bg // num_groups if num_groups else 0, bg % num_groups if num_groups else bg, (tag >> 2) & 0x3FF, event_type,
                (tag >> 24) & 0xFF)
    return 0, 0, tag >> 2, event_type, 0


def alloc_profiler_buffer(num_blocks: int, num_groups: int, max_events_per_group: int = 256, device="cuda") -> torch.Tensor:
    n = 1 + num_blocks * num_groups * (max_events_per_group + 1)
    return torch.zeros(n, dtype=torch.int64, device=device)


def decode_profiler_buffer(buf: torch.Tensor, max_events_per_group: int = 256) -> List[Dict]:
    """-> list of {block, group, sm, event, type, t_ns}"""
    b = buf.detach().cpu().numpy().astype("uint64")
    hdr = int(b[0])
    nblocks, ngroups = hdr >> 32, hdr & 0xFFFFFFFF
    out = []
    stride = max_events_per_group + 1
    for blk in range(nblocks):
        for g in range(ngroups):
            base = 1 + (blk * ngroups + g) * stride
            h = int(b[base])
            count = h & 0xFFFFFFFF
            sm = (h >> 48) & 0xFFFF
            for i in range(min(count, max_events_per_group)):
                e = int(b[base + 1 + i])
                tag, t = e >> 32, e & 0xFFFFFFFF
                out.append({"block": blk, "group": g, "sm": sm, "event": tag >> 2, "type": tag & 3, "t_ns": t})
    return out


def export_to_perfetto_trace(buf: torch.Tensor, event_names: Sequence[str], file_name: str,
                             max_events_per_group: int = 256, group_names: Optional[Sequence[str]] = None) -> int:
    """Writes a Chrome-trace JSON (open in ui.perfetto.dev).  pid = SM, tid = (block, group)."""
    ev = decode_profiler_buffer(buf, max_events_per_group)
    if not ev:
        with open(file_name, "w") as f:
            json.dump({"traceEvents": []}, f)
        return 0
    t0 = min(e["t_ns"] for e in ev)
    trace = []
    for e in ev:
        name = event_names[e["event"]] if e["event"] < len(event_names) else f"event{e['event']}"
        tid = e["block"] * 64 + e["group"]
        ph = {START: "B", END: "E", INSTANT: "i"}.get(e["type"], "i")
        # globaltimer_lo wraps every ~4.29 s; a single kernel never spans a wrap in practice
        rec = {"name": name, "ph": ph, "ts": ((e["t_ns"] - t0) & 0xFFFFFFFF) / 1e3, "pid": e["sm"], "tid": tid}
        if ph == "i":
            rec["s"] = "t"
        trace.append(rec)
    meta = []
    for sm in sorted({e["sm"] for e in ev}):
        meta.append({"name": "process_name", "ph": "M", "pid": sm, "args": {"name": f"SM {sm}"}})
    for e in {(e["sm"], e["block"], e["group"]) for e in ev}:
        gname = group_names[e[2]] if group_names and e[2] < len(group_names) else f"group {e[2]}"
        meta.append({"name": "thread_name", "ph": "M", "pid": e[0], "tid": e[1] * 64 + e[2], "args": {"name": f"cta {e[1]} / {gname}"}})
    with open(file_name, "w") as f:
        json.dump({"traceEvents": meta + trace, "displayTimeUnit": "ns"}, f)
    return len(trace)
