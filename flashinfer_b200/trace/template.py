"""Declarative op schemas -> "benchmark definition" JSON files, one per unique (op, constant axes) pair.

Parity: reference flashinfer/trace/template.py:1-634 and flashinfer/fi_trace.py:88-285.  A template names the axes of
an op (``Const`` axes identify a definition, ``Var`` axes vary per call), the tensors in terms of those axes and a
reference implementation; ``@fi_trace(template)`` records a definition the first time each constant-axis combination
is seen when ``FLASHINFER_TRACE_DIR`` is set, and is a no-op otherwise.
"""
from __future__ import annotations

import functools
import inspect
import json
import os
import threading
from dataclasses import dataclass, field
from typing import Any, Callable, Dict, List, Optional, Sequence, Tuple

import torch

_TRACE_DIR = os.environ.get("FLASHINFER_TRACE_DIR")
_seen: set = set()
_lock = threading.Lock()
_TEMPLATES: Dict[str, "TraceTemplate"] = {}


@dataclass(frozen=True)
class Const:
    name: str
    description: str = ""


@dataclass(frozen=True)
class Var:
    name: str
    description: str = ""


@dataclass(frozen=True)
class Tensor:
    name: str
    axes: Tuple[str, ...]
    dtype: Optional[str] = None  # None = taken from the live tensor
    optional: bool = False
    description: str = ""


@dataclass(frozen=True)
class Scalar:
    name: str
    dtype: str = "float32"
    description: str = ""


@dataclass
class TraceTemplate:
    op_type: str
    name_fmt: str  # e.g. "rmsnorm_h{hidden}"
    axes: Sequence[Any]
    inputs: Sequence[Any]
    outputs: Sequence[Tensor]
    reference: Optional[Callable] = None
    tags: Sequence[str] = field(default_factory=tuple)
    description: str = ""

    def __post_init__(self):
        _TEMPLATES[self.op_type + ":" + self.name_fmt] = self

    def resolve_axes(self, bound: Dict[str, Any]) -> Dict[str, int]:
        sizes: Dict[str, int] = {}
        for spec in self.inputs:
            if isinstance(spec, Tensor):
                t = bound.get(spec.name)
                if isinstance(t, torch.Tensor):
                    for ax, n in zip(spec.axes, t.shape[-len(spec.axes):] if spec.axes else ()):
                        sizes.setdefault(ax, int(n))
        for ax in self.axes:
            if ax.name in bound and isinstance(bound[ax.name], int):
                sizes.setdefault(ax.name, bound[ax.name])
        return sizes

    def definition(self, bound: Dict[str, Any]) -> Dict[str, Any]:
        sizes = self.resolve_axes(bound)
        consts = {a.name: sizes.get(a.name) for a in self.axes if isinstance(a, Const)}
        name = self.name_fmt.format(**{k: v for k, v in consts.items() if v is not None})
        axes = {}
        for a in self.axes:
            if isinstance(a, Const):
                axes[a.name] = {"type": "const", "value": sizes.get(a.name), "description": a.description}
            else:
                axes[a.name] = {"type": "var", "description": a.description}

        def tdesc(spec):
            if isinstance(spec, Scalar):
                return {"shape": None, "dtype": spec.dtype, "description": spec.description}
            live = bound.get(spec.name)
            dt = spec.dtype or (str(live.dtype).replace("torch.", "") if isinstance(live, torch.Tensor) else "unknown")
            return {"shape": list(spec.axes), "dtype": dt, "optional": spec.optional, "description": spec.description}

        ref_src = None
        if self.reference is not None:
            try:
                ref_src = inspect.getsource(self.reference)
            except (OSError, TypeError):
                ref_src = None
        return {"name": name, "op_type": self.op_type, "description": self.description, "tags": list(self.tags), "axes": axes,
                "inputs": {s.name: tdesc(s) for s in self.inputs}, "outputs": {s.name: tdesc(s) for s in self.outputs},
                "reference": ref_src}


def registered_templates() -> Dict[str, TraceTemplate]:
    return dict(_TEMPLATES)


def fi_trace(template: TraceTemplate, trace_dir: Optional[str] = None):
    """Decorator: emit ``<trace_dir>/<op_type>/<name>.json`` once per unique definition."""

    def deco(fn):
        out_dir = trace_dir or _TRACE_DIR
        if not out_dir:
            fn.__fi_trace_template__ = template
            return fn
        sig = inspect.signature(fn)

        @functools.wraps(fn)
        def wrapper(*args, **kwargs):
            try:
                bound = dict(sig.bind(*args, **kwargs).arguments)
                d = template.definition(bound)
                key = (d["op_type"], d["name"])
                with _lock:
                    fresh = key not in _seen
                    _seen.add(key)
                if fresh:
                    p = os.path.join(out_dir, d["op_type"])
                    os.makedirs(p, exist_ok=True)
                    with open(os.path.join(p, d["name"] + ".json"), "w") as f:
                        json.dump(d, f, indent=1)
            except Exception:  # noqa: BLE001 - tracing must never break the op
                pass
            return fn(*args, **kwargs)

        wrapper.__fi_trace_template__ = template
        return wrapper

    return deco
