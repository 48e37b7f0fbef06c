"""Declarative op schemas -> "benchmark definition" JSON files, one per unique (op, constant axes) pair.

Parity: reference flashinfer/trace/template.py:1-634 and flashinfer/fi_trace.py:88-285.  A template names the axes of an op
(``Const`` axes identify a definition, ``Var`` axes vary per call), the tensors in terms of those axes, a plain-PyTorch
reference implementation and an input builder.  ``bindings.py`` attaches templates to the public functions; the definition
of a call is available as ``fn.fi_trace(**kwargs)`` at any time, and every traced call writes its definition once per
unique name while dumping is on (``FLASHINFER_TRACE_DUMP=1`` + ``FLASHINFER_TRACE_DUMP_DIR``, the older
``FLASHINFER_TRACE_DIR``, or :func:`enable`).  With dumping off the public functions are the undecorated originals.

What is different from the reference: descriptors carry their own names (ordered lists, not dicts), the reference
implementation is *callable through the template* (``run_reference`` maps API arguments onto it, ``make_inputs`` builds
arguments from axis sizes), so one generic test can check every template against the API it is bound to on any device.
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

_lock = threading.Lock()
_seen: set = set()
_TEMPLATES: Dict[str, "TraceTemplate"] = {}


def _env_dir() -> Optional[str]:
    d = os.environ.get("FLASHINFER_TRACE_DIR")
    if d:
        return d
    if os.environ.get("FLASHINFER_TRACE_DUMP", "0") not in ("", "0", "false", "False"):
        return os.environ.get("FLASHINFER_TRACE_DUMP_DIR", "fi_trace_out")
    return None


class _State:
    dump_dir: Optional[str] = _env_dir()


def dump_dir() -> Optional[str]:
    return _State.dump_dir


def _dtype_str(dt) -> str:
    return str(dt).replace("torch.", "")


@dataclass(frozen=True)
class Const:
    """Axis whose value is part of the definition's identity (``abbrev``: prefix in the file name; ``""`` leaves it out)."""
    name: str
    description: str = ""
    abbrev: Optional[str] = None


@dataclass(frozen=True)
class Var:
    """Axis that varies from call to call (batch size, sequence length ...)."""
    name: str
    description: str = ""


@dataclass(frozen=True)
class Tensor:
    """``name`` is the reference function's argument; ``param`` (default: same) the API argument it is read from,
    ``tuple_idx`` picks a member when that argument is a tuple (``paged_kv_cache=(k, v)``).  Outputs take their dtype from
    ``dtype`` or from the input named by ``dtype_from``."""
    name: str
    axes: Tuple[str, ...]
    dtype: Optional[str] = None
    optional: bool = False
    description: str = ""
    param: Optional[str] = None        # API argument, or "self.<attr>" for state a wrapper captured in plan()
    tuple_idx: Optional[int] = None
    dtype_from: Optional[str] = None

    @property
    def source(self) -> str:
        return self.param or self.name


@dataclass(frozen=True)
class Scalar:
    name: str
    dtype: str = "float32"
    description: str = ""
    param: Optional[str] = None
    optional: bool = False

    @property
    def source(self) -> str:
        return self.param or self.name


def _pick(bound: Dict[str, Any], spec) -> Any:
    src = spec.source
    if src.startswith("self."):                      # plan()-time state of a wrapper object: "self._kv_indptr_host"
        v = bound.get("self")
        for part in src[5:].split("."):                # a numeric part indexes a list / tuple held by the wrapper
            if part.isdigit():
                v = v[int(part)] if isinstance(v, (list, tuple)) and len(v) > int(part) else None
            else:
                v = getattr(v, part, None)
    else:
        v = bound.get(src)
    idx = getattr(spec, "tuple_idx", None)
    if idx is not None:
        v = v[idx] if isinstance(v, (tuple, list)) and len(v) > idx else None
    return v


@dataclass
class TraceTemplate:
    op_type: str
    name_fmt: str                      # "rmsnorm_h{hidden_size}"; "" = op_type + the abbreviated const axes
    axes: Sequence[Any]
    inputs: Sequence[Any]
    outputs: Sequence[Tensor]
    reference: Optional[Callable] = None
    init: Optional[Callable] = None    # init(**axis sizes, device=, seed=) -> kwargs of the API
    tags: Sequence[str] = field(default_factory=tuple)
    description: str = ""
    constraints: Sequence[str] = field(default_factory=tuple)
    tolerance: str = "bf16"            # row of tests/test_trace_templates.py::TOLERANCE ("exact", "fp16", "bf16", "cos", ...)
    fi_api: str = ""                   # filled by bindings: dotted name of the bound function
    derive: Optional[Callable] = None  # derive(sizes) -> extra axis sizes computed from the resolved ones
    test_sizes: Optional[Dict[str, int]] = None   # small axis sizes for the generic reference-correctness test
    compare: Optional[Callable] = None  # compare(got, expected, kwargs): op-specific check replacing the tolerance class
    helpers: Sequence[Callable] = field(default_factory=tuple)   # functions ``reference`` calls: their source is emitted in front of it

    def __post_init__(self):
        _TEMPLATES[self.key] = self

    @property
    def key(self) -> str:
        return self.op_type + ":" + (self.name_fmt or self.op_type)

    # ---- axes
    def resolve_axes(self, bound: Dict[str, Any]) -> Dict[str, int]:
        sizes: Dict[str, int] = {}
        for spec in self.inputs:
            if isinstance(spec, Tensor):
                t = _pick(bound, spec)
                if isinstance(t, torch.Tensor) and t.dim() >= len(spec.axes):
                    dims = t.shape[t.dim() - len(spec.axes):] if spec.axes else ()
                    for ax, n in zip(spec.axes, dims):
                        sizes.setdefault(ax, int(n))
        names = {a.name for a in self.axes}
        for spec in self.inputs:                          # integer scalars named like an axis (also plan()-time state: "self.chunk_size")
            if isinstance(spec, Scalar) and spec.name in names:
                v = _pick(bound, spec)
                if isinstance(v, int) and not isinstance(v, bool):
                    sizes.setdefault(spec.name, v)
        for ax in self.axes:
            v = bound.get(ax.name)
            if isinstance(v, int) and not isinstance(v, bool):
                sizes.setdefault(ax.name, v)
        if self.derive is not None:
            for k, v in self.derive(sizes).items():
                sizes.setdefault(k, v)
        return sizes

    def definition_name(self, sizes: Dict[str, int]) -> str:
        consts = {a.name: sizes.get(a.name) for a in self.axes if isinstance(a, Const)}
        if self.name_fmt:
            try:
                return self.name_fmt.format(**{k: v for k, v in consts.items() if v is not None})
            except KeyError:
                pass
        parts = []
        for a in self.axes:
            if isinstance(a, Const) and consts.get(a.name) is not None and a.abbrev != "":
                parts.append(f"{a.abbrev if a.abbrev is not None else a.name}{consts[a.name]}")
        base = self.name_fmt.split("{")[0].rstrip("_") if self.name_fmt else self.op_type
        return "_".join([base] + parts)

    # ---- definition JSON
    def definition(self, bound: Dict[str, Any]) -> Dict[str, Any]:
        sizes = self.resolve_axes(bound)
        axes: Dict[str, Any] = {}
        for a in self.axes:
            entry: Dict[str, Any] = {"type": "const" if isinstance(a, Const) else "var"}
            if isinstance(a, Const):
                entry["value"] = sizes.get(a.name)
            if a.description:
                entry["description"] = a.description
            axes[a.name] = entry

        def describe(spec, is_output: bool) -> Dict[str, Any]:
            if isinstance(spec, Scalar):
                live = bound.get(spec.source)
                e = {"shape": None, "dtype": _dtype_str(live.dtype) if isinstance(live, torch.Tensor) else spec.dtype}
            else:
                live = None if is_output else _pick(bound, spec)
                dt = spec.dtype
                if is_output and spec.dtype_from:
                    src = bound.get(spec.dtype_from)
                    dt = _dtype_str(src.dtype) if isinstance(src, torch.Tensor) else dt
                if dt is None:
                    dt = _dtype_str(live.dtype) if isinstance(live, torch.Tensor) else "unknown"
                e = {"shape": list(spec.axes), "dtype": dt}
            if spec.optional:
                e["optional"] = True
            if spec.description:
                e["description"] = spec.description
            return e

        out: Dict[str, Any] = {"name": self.definition_name(sizes), "description": self.description, "op_type": self.op_type,
                               "tags": ([f"fi_api:{self.fi_api}"] if self.fi_api else []) + list(self.tags), "axes": axes}
        if self.constraints:
            out["constraints"] = list(self.constraints)
        out["inputs"] = {s.name: describe(s, False) for s in self.inputs}
        out["outputs"] = {s.name: describe(s, True) for s in self.outputs}
        for label, fn in (("reference", self.reference), ("init", self.init)):
            if fn is not None:
                try:
                    parts = [inspect.getsource(h) for h in self.helpers] if label == "reference" else []
                    out[label] = "\n\n".join(parts + [inspect.getsource(fn)])
                except (OSError, TypeError):
                    out[label] = None
        return out

    def dump(self, bound: Dict[str, Any], save_dir: Optional[str] = None) -> Dict[str, Any]:
        d = self.definition(bound)
        target = save_dir if save_dir is not None else (_State.dump_dir or os.environ.get("FLASHINFER_TRACE_DUMP_DIR"))
        if target:
            p = os.path.join(str(target), d["op_type"])
            os.makedirs(p, exist_ok=True)
            with open(os.path.join(p, d["name"] + ".json"), "w") as f:
                json.dump(d, f, indent=1)
        return d

    # ---- executable side
    def make_inputs(self, device="cpu", seed: int = 0, **sizes) -> Dict[str, Any]:
        """API keyword arguments for the given axis sizes (defaults come from the template's ``init``)."""
        if self.init is None:
            raise ValueError(f"template {self.key} has no input builder")
        accepted = inspect.signature(self.init).parameters
        return self.init(device=device, seed=seed, **{k: v for k, v in sizes.items() if k in accepted})

    def reference_kwargs(self, bound: Dict[str, Any]) -> Dict[str, Any]:
        """The arguments of ``reference`` picked out of API keyword arguments (optional inputs that are absent are dropped)."""
        accepted = set(inspect.signature(self.reference).parameters) if self.reference is not None else set()
        kw = {}
        for spec in self.inputs:
            v = _pick(bound, spec)
            if v is None and spec.optional:
                continue
            if spec.name in accepted:
                kw[spec.name] = v
        return kw

    def collect_outputs(self, result: Any, bound: Dict[str, Any]) -> List[Any]:
        """The API's outputs in template order: outputs with ``param`` are read back from that argument (in-place and
        ``out=`` style APIs), the others are taken in order from the returned value."""
        returned = list(result) if isinstance(result, (tuple, list)) else [result]
        outs = []
        for spec in self.outputs:
            if spec.param is not None:
                outs.append(_pick(bound, spec))
            else:
                outs.append(returned.pop(0) if returned else None)
        return outs

    def run_reference(self, bound: Dict[str, Any]):
        if self.reference is None:
            raise ValueError(f"template {self.key} has no reference implementation")
        with torch.no_grad():
            return self.reference(**self.reference_kwargs(bound))

    def build_fi_trace_fn(self, fi_api: str = "") -> Callable[..., Dict[str, Any]]:
        if fi_api:
            self.fi_api = fi_api

        def trace(save_dir: Optional[str] = None, **kwargs) -> Dict[str, Any]:
            return self.dump(kwargs, save_dir) if (save_dir or _State.dump_dir) else self.definition(kwargs)

        return trace


class TemplateDispatch:
    """Several templates behind one API: ``select(bound)`` names the member that describes a given call (reference
    flashinfer/api_logging.py:2182-2291, the ``trace=<callable>`` form - e.g. one MoE entry point whose definition depends
    on ``routing_method_type``).  Offers the subset of :class:`TraceTemplate` the binding / tracing code uses."""

    def __init__(self, templates: Sequence["TraceTemplate"], select: Callable[[Dict[str, Any]], "TraceTemplate"]):
        self.templates = tuple(templates)
        self._select = select

    def pick(self, bound: Dict[str, Any]) -> "TraceTemplate":
        tpl = self._select(bound)
        if tpl not in self.templates:
            raise ValueError("dispatch returned a template that is not one of its members")
        return tpl

    @property
    def fi_api(self) -> str:
        return self.templates[0].fi_api

    @fi_api.setter
    def fi_api(self, value: str) -> None:
        for t in self.templates:
            t.fi_api = value

    def definition(self, bound: Dict[str, Any]) -> Dict[str, Any]:
        return self.pick(bound).definition(bound)

    def dump(self, bound: Dict[str, Any], save_dir: Optional[str] = None) -> Dict[str, Any]:
        return self.pick(bound).dump(bound, save_dir)

    def build_fi_trace_fn(self, fi_api: str = "") -> Callable[..., Dict[str, Any]]:
        if fi_api:
            self.fi_api = fi_api

        def trace(save_dir: Optional[str] = None, **kwargs) -> Dict[str, Any]:
            return self.dump(kwargs, save_dir) if (save_dir or _State.dump_dir) else self.definition(kwargs)

        return trace


def concrete_templates(tpl) -> Tuple["TraceTemplate", ...]:
    """The member templates of a dispatch, or the template itself."""
    return tuple(tpl.templates) if isinstance(tpl, TemplateDispatch) else (tpl,)


def registered_templates() -> Dict[str, TraceTemplate]:
    return dict(_TEMPLATES)


def _emit_once(template, bound: Dict[str, Any], out_dir: str) -> None:
    try:
        if isinstance(template, TemplateDispatch):
            template = template.pick(bound)
        sizes = template.resolve_axes(bound)
        key = (template.op_type, template.definition_name(sizes), out_dir)
        with _lock:
            fresh = key not in _seen
            _seen.add(key)
        if fresh:
            template.dump(bound, out_dir)
    except Exception:  # noqa: BLE001 - tracing must never break the op
        pass


def traced(fn: Callable, template: TraceTemplate) -> Callable:
    """``fn`` wrapped so every call records its definition (once per name) while dumping is on."""
    sig = inspect.signature(fn)

    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        out_dir = _State.dump_dir
        if out_dir:
            try:
                bound = dict(sig.bind(*args, **kwargs).arguments)
            except TypeError:
                bound = None
            if bound is not None:
                _emit_once(template, bound, out_dir)
        return fn(*args, **kwargs)

    wrapper.__fi_trace_template__ = template
    wrapper.fi_trace = template.build_fi_trace_fn()
    wrapper.__wrapped_untraced__ = fn
    return wrapper


def fi_trace(template: TraceTemplate, trace_dir: Optional[str] = None):
    """Decorator form: attach ``template`` to a function.  With dumping off (and no ``trace_dir``) the function itself is
    returned, carrying ``.fi_trace`` / ``.__fi_trace_template__``."""

    def deco(fn):
        if trace_dir:
            sig = inspect.signature(fn)

            @functools.wraps(fn)
            def pinned(*args, **kwargs):
                try:
                    _emit_once(template, dict(sig.bind(*args, **kwargs).arguments), trace_dir)
                except TypeError:
                    pass
                return fn(*args, **kwargs)

            pinned.__fi_trace_template__ = template
            pinned.fi_trace = template.build_fi_trace_fn()
            return pinned
        if _State.dump_dir:
            return traced(fn, template)
        try:
            fn.__fi_trace_template__ = template
            fn.fi_trace = template.build_fi_trace_fn()
        except AttributeError:
            pass
        return fn

    return deco
