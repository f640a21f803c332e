"""Inference API — host-side mirror of ``src/inference_engine.jl`` and ``src/dependencies.jl``.

Same names, argument meaning and error behaviour as the reference; the work is done behind the
C ABI (``include/cortex_b200.h``): graph ingestion through the 7 backend generics, dependency
resolution, ``request_inference_for`` / ``scan_inference_request`` / ``update_marginals!``.
Rules are *registered kernels keyed by factor type* (``RuleProcessor``) instead of Julia methods.
"""
from __future__ import annotations

import ctypes as C
import time
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import _capi as capi
from .inference_signal import (IndividualMarginal, JointMarginal, MessageToFactor, MessageToVariable, NoRuleError,
                               ProductOfMessages, Signal, SignalStore, UndefValue, get_value, get_values, get_variant,
                               is_computed, _ids)
from .model_engine import (B200ModelEngine, Connection, Factor, Variable, backend_get_connected_factor_ids,
                           backend_get_connected_variable_ids, backend_get_connection, backend_get_factor,
                           backend_get_factor_ids, backend_get_variable, backend_get_variable_ids,
                           throw_if_engine_unsupported)


# ---- dependency resolvers: src/dependencies.jl:1-3 -------------------------------------------------
class AbstractDependencyResolver:
    """A resolver is either one of the built-in wirings (``resolver_kind``: run inside the library) or a user subclass
    that overrides ``resolve_factor_dependencies`` / ``resolve_variable_dependencies`` and calls ``add_dependency``
    itself, as the reference's tests do (test/inference_engine_tests.jl:597-621, 811-907)."""
    resolver_kind = capi.RESOLVER_NONE

    def resolve_factor_dependencies(self, engine, factor_id):  # src/dependencies.jl:17
        if self.resolver_kind == capi.RESOLVER_NONE:
            raise NotImplementedError
        engine.store.check(engine.api.resolve_factor_dependencies(engine.store.h, self.resolver_kind, int(factor_id)))

    def resolve_variable_dependencies(self, engine, variable_id):  # src/dependencies.jl:33
        if self.resolver_kind == capi.RESOLVER_NONE:
            raise NotImplementedError
        engine.store.check(engine.api.resolve_variable_dependencies(engine.store.h, self.resolver_kind, int(variable_id)))

    def resolve_dependencies(self, engine):
        """resolve_dependencies!(resolver, engine), src/dependencies.jl:5-15: every factor, then every variable."""
        if self.resolver_kind != capi.RESOLVER_NONE:
            engine.store.check(engine.api.resolve_dependencies(engine.store.h, self.resolver_kind))
            return
        for factor_id in get_factor_ids(engine):
            self.resolve_factor_dependencies(engine, factor_id)
        for variable_id in get_variable_ids(engine):
            self.resolve_variable_dependencies(engine, variable_id)


class DefaultDependencyResolver(AbstractDependencyResolver):
    """Sum-product BP wiring incl. the >5-neighbour segment tree, src/dependencies.jl:17-173."""
    resolver_kind = capi.RESOLVER_DEFAULT_BP


class MeanFieldResolver(AbstractDependencyResolver):
    """The weak-dependency resolver of test/inference_engine_tests.jl:597-621."""
    resolver_kind = capi.RESOLVER_MEAN_FIELD


# ---- request processors: src/inference_engine.jl:331 -----------------------------------------------
class AbstractInferenceRequestProcessor:
    value_dim = 1
    family = capi.FAMILY_SUM
    rules: Dict[Any, Tuple[int, Sequence[float]]] = {}


class InferenceRequestScanner(AbstractInferenceRequestProcessor):
    """Default processor of the reference constructor (src/inference_engine.jl:528-537): collects only."""


class RuleProcessor(AbstractInferenceRequestProcessor):
    """Rules registered by factor type (``Factor.functional_form``): ``{form: (CXB_RULE_*, params)}``.

    Replaces user methods of compute_message_to_variable! (per factor type) and the
    compute_message_to_factor! / compute_product_of_messages! / compute_individual_marginal!
    family (``family`` = how values of this model multiply, CXB_FAMILY_*).
    """

    def __init__(self, rules: Dict[Any, Tuple[int, Sequence[float]]], family: int, value_dim: int):
        self.rules = dict(rules)
        self.family = family
        self.value_dim = value_dim


class _LazyNeighbours:
    """variable id -> tuple of connected factor ids, computed on demand (ProductOfMessages variants of a lazy model engine)."""

    def __init__(self, model_engine):
        self.me = model_engine

    def get(self, v, default=()):
        try:
            return tuple(self.me.get_connected_factor_ids(v))
        except KeyError:
            return default

    def __setitem__(self, k, v):
        pass


class RuleProgram:
    """A user-defined message-to-variable rule for fixed-size values (``CXB_RULE_PROGRAM``): the counterpart of writing a new
    method of ``compute_message_to_variable!`` (src/inference_engine.jl:351-361) - a new factor type needs no rebuild of the
    library. Build the program in postfix order, e.g. the random-walk rule (L, h) -> (L / (1 + q L), h / (1 + q L)):

        p = RuleProgram(default_param=1.0)
        p.const(1.0).param().dep(0, 0).mul().add().tset(0)          # den = 1 + q * L
        p.dep(0, 0).tget(0).div().store(0).dep(0, 1).tget(0).div().store(1)
        RuleProcessor({"transition": p.rule()}, family=FAMILY_GAUSS_CANON, value_dim=2)
    """

    def __init__(self, default_param: float = 1.0):
        self.consts: List[float] = [float(default_param)]  # consts[0] is the default of the per-factor parameter
        self.code: List[float] = []

    def _emit(self, *xs):
        self.code.extend(float(x) for x in xs)
        return self

    def dep(self, i, k):
        return self._emit(capi.OP_DEP, i, k)

    def const(self, c):
        c = float(c)
        if c not in self.consts[1:]:
            self.consts.append(c)
        return self._emit(capi.OP_CONST, 1 + self.consts[1:].index(c))

    def param(self):
        return self._emit(capi.OP_PARAM)

    def ndeps(self):
        return self._emit(capi.OP_NDEPS)

    def add(self):
        return self._emit(capi.OP_ADD)

    def sub(self):
        return self._emit(capi.OP_SUB)

    def mul(self):
        return self._emit(capi.OP_MUL)

    def div(self):
        return self._emit(capi.OP_DIV)

    def neg(self):
        return self._emit(capi.OP_NEG)

    def exp(self):
        return self._emit(capi.OP_EXP)

    def log(self):
        return self._emit(capi.OP_LOG)

    def sqrt(self):
        return self._emit(capi.OP_SQRT)

    def store(self, k):
        return self._emit(capi.OP_STORE, k)

    def tset(self, j):
        return self._emit(capi.OP_TSET, j)

    def tget(self, j):
        return self._emit(capi.OP_TGET, j)

    def rule(self) -> Tuple[int, List[float]]:
        """``(CXB_RULE_PROGRAM, params)`` for a ``RuleProcessor``."""
        return capi.RULE_PROGRAM, [float(len(self.consts))] + self.consts + self.code


@dataclass
class InferenceEngineWarning:  # src/inference_engine.jl:11-14
    description: str
    context: Any


# ---- tracing: src/inference_engine.jl:650-754 ---------------------------------------------------------
@dataclass
class TracedInferenceExecution:
    engine: Any
    variable_id: Any
    signal: Signal
    total_time_in_ns: int
    value_before_execution: Any
    value_after_execution: Any


@dataclass
class TracedInferenceRound:
    engine: Any
    total_time_in_ns: int
    executions: List[TracedInferenceExecution]


@dataclass
class TracedInferenceRequest:
    engine: Any
    total_time_in_ns: int
    request: Any
    rounds: List[TracedInferenceRound]


@dataclass
class InferenceEngineTracer:
    inference_requests: List[TracedInferenceRequest] = field(default_factory=list)


@dataclass
class InferenceRequest:  # src/inference_engine.jl:265-270
    engine: Any
    variable_ids: Tuple[int, ...]
    marginals: List[Signal]


class InferenceEngine:
    """InferenceEngine(; model_engine, dependency_resolver, inference_request_processor,
    prepare_signals_metadata, resolve_dependencies, trace), src/inference_engine.jl:53-90."""

    def __init__(self, *, model_engine, dependency_resolver=None, inference_request_processor=None,
                 prepare_signals_metadata: bool = True, resolve_dependencies: bool = True, trace: bool = False,
                 dtype: int = capi.F64, device: int = 0, api: Optional[capi.CApi] = None):
        self.model_engine = throw_if_engine_unsupported(model_engine)
        resolver = DefaultDependencyResolver() if dependency_resolver is None else dependency_resolver
        processor = InferenceRequestScanner() if inference_request_processor is None else inference_request_processor
        if not isinstance(resolver, AbstractDependencyResolver):  # convert(...) has no methods, :69
            raise TypeError("dependency_resolver must be an AbstractDependencyResolver")
        if not isinstance(processor, AbstractInferenceRequestProcessor):  # :70
            raise TypeError("inference_request_processor must be an AbstractInferenceRequestProcessor")
        self.dependency_resolver = resolver
        self.inference_request_processor = processor
        self.tracer = InferenceEngineTracer() if trace else None
        self.warnings: List[InferenceEngineWarning] = []
        self.api = api if api is not None else capi.default_api()
        self.store = SignalStore(self.api, processor.value_dim, processor.family, dtype, device)
        self._ingest()
        self._register_rules()
        if hasattr(processor, "install"):  # processors that hook into the engine themselves (tests/oracle_frontend.py)
            processor.install(self)
        if trace:
            self.store.check(self.api.trace_enable(self.store.h, 1))
        if resolve_dependencies:
            resolver.resolve_dependencies(self)
            n = self.api.get_warnings(self.store.h, None, 0)
            if n > 0:
                buf = np.zeros(n, dtype=np.int64)
                self.api.get_warnings(self.store.h, buf.ctypes.data_as(capi.i64p), n)
                for v in buf:  # src/dependencies.jl:40-43
                    self.warnings.append(InferenceEngineWarning("Variable has no connected factors", int(v)))

    # -- graph ingestion through the 7 generics (src/model_engine.jl:329-391) ---------------------------
    def _ingest_lazy(self):
        """B200ModelEngine: the graph arrives as flat arrays; no Variable / Factor / Connection object is created here - the
        model engine materialises views on demand (mirror of julia/CortexB200.jl)."""
        me, st = self.model_engine, self.store
        self._type_of_form: Dict[Any, int] = {}
        ftype = np.zeros(max(me.n_ids, 1), dtype=np.int32)
        forms = me.functional_forms
        for f in np.flatnonzero(me.is_factor):
            ftype[f] = self._type_of_form.setdefault(forms[int(f)], len(self._type_of_form))
        st.check(self.api.graph_build(st.h, me.n_ids, me.is_factor.ctypes.data_as(capi.u8p), ftype.ctypes.data_as(capi.i32p),
                                      len(me.edge_variable), me.edge_variable.ctypes.data_as(capi.i64p),
                                      me.edge_factor.ctypes.data_as(capi.i64p)))
        me._engine = self
        st._neighbours = _LazyNeighbours(me)

    def _ingest(self):
        me = self.model_engine
        self._links: Dict[int, List[Signal]] = {}
        if isinstance(me, B200ModelEngine):
            return self._ingest_lazy()
        vids = [int(v) for v in backend_get_variable_ids(me)]
        fids = [int(f) for f in backend_get_factor_ids(me)]
        n_ids = (max(vids + fids) + 1) if (vids or fids) else 0
        is_factor = np.zeros(max(n_ids, 1), dtype=np.uint8)
        ftype = np.zeros(max(n_ids, 1), dtype=np.int32)
        self._type_of_form: Dict[Any, int] = {}
        for f in fids:
            is_factor[f] = 1
            form = backend_get_factor(me, f).functional_form
            ftype[f] = self._type_of_form.setdefault(form, len(self._type_of_form))
        edges = me.edges() if hasattr(me, "edges") else [
            (int(v), f) for f in fids for v in backend_get_connected_variable_ids(me, f)]
        ev = np.ascontiguousarray([e[0] for e in edges], dtype=np.int64)
        ef = np.ascontiguousarray([e[1] for e in edges], dtype=np.int64)
        st = self.store
        st.check(self.api.graph_build(st.h, n_ids, is_factor.ctypes.data_as(capi.u8p),
                                      ftype.ctypes.data_as(capi.i32p), len(edges),
                                      ev.ctypes.data_as(capi.i64p), ef.ctypes.data_as(capi.i64p)))
        self._variable_ids, self._factor_ids = vids, fids
        for v in vids:  # bind the signal references (set_signals_variants!, :228-247, is done by graph_build)
            var = backend_get_variable(me, v)
            var.marginal = Signal(st, self.api.signal_id(st.h, capi.KIND_MARGINAL, v, -1))
            var._engine, var._id = self, v
            st._neighbours[v] = tuple(backend_get_connected_factor_ids(me, v))
        for (v, f) in edges:
            c = backend_get_connection(me, v, f)
            c.message_to_variable = Signal(st, self.api.signal_id(st.h, capi.KIND_M2V, v, f))
            c.message_to_factor = Signal(st, self.api.signal_id(st.h, capi.KIND_M2F, v, f))

    def _register_rules(self):
        st = self.store
        for form, (kind, params) in getattr(self.inference_request_processor, "rules", {}).items():
            if form not in self._type_of_form:
                continue
            p = np.ascontiguousarray(np.asarray(params, dtype=np.float64).ravel())
            st.check(self.api.register_rule(st.h, self._type_of_form[form], int(kind),
                                            p.ctypes.data_as(capi.f64p) if p.size else None, p.size))

    def __repr__(self):  # src/inference_engine.jl:92-100
        return f"InferenceEngine(trace = {'true' if self.tracer is not None else 'false'})"


# ---- accessors, src/inference_engine.jl:119-205 ---------------------------------------------------------
def get_model_engine(engine: InferenceEngine):
    return engine.model_engine


def get_inference_request_processor(engine: InferenceEngine):
    return engine.inference_request_processor


def get_trace(engine: InferenceEngine):
    return engine.tracer


def get_warnings(engine: InferenceEngine):
    return engine.warnings


def add_warning(engine: InferenceEngine, description: str, context) -> None:
    """add_warning!(engine, description, context), src/inference_engine.jl:127-129."""
    engine.warnings.append(InferenceEngineWarning(description, context))


def resolve_dependencies(resolver: AbstractDependencyResolver, engine: InferenceEngine) -> None:
    """resolve_dependencies!(resolver, engine), src/dependencies.jl:5-15."""
    resolver.resolve_dependencies(engine)


def format_time_ns(ns: int) -> str:
    """format_time_ns(ns), src/utils.jl:2-16 (used when a trace is printed)."""
    ns = int(ns)
    if ns < 1_000:
        return f"{ns} ns"
    for limit, div, unit in ((1_000_000, 1_000, "μs"), (1_000_000_000, 1_000_000, "ms"), (60_000_000_000, 1_000_000_000, "s"),
                             (3_600_000_000_000, 60_000_000_000, "min"), (None, 3_600_000_000_000, "hr")):
        if limit is None or ns < limit:
            return f"{float(round(ns / div * 100) / 100)!r} {unit}"  # round(x, digits = 2), shortest float repr as in Julia
    raise AssertionError("unreachable")


def get_variable(engine: InferenceEngine, variable_id: int) -> Variable:
    return backend_get_variable(engine.model_engine, variable_id)


def get_variable_ids(engine: InferenceEngine):
    return backend_get_variable_ids(engine.model_engine)


def get_factor(engine: InferenceEngine, factor_id: int) -> Factor:
    return backend_get_factor(engine.model_engine, factor_id)


def get_factor_ids(engine: InferenceEngine):
    return backend_get_factor_ids(engine.model_engine)


def get_connection(engine: InferenceEngine, variable_id: int, factor_id: int) -> Connection:
    return backend_get_connection(engine.model_engine, variable_id, factor_id)


def get_connection_message_to_variable(engine_or_connection, variable_id=None, factor_id=None) -> Signal:
    if isinstance(engine_or_connection, Connection):
        return engine_or_connection.message_to_variable
    return get_connection(engine_or_connection, variable_id, factor_id).message_to_variable


def get_connection_message_to_factor(engine_or_connection, variable_id=None, factor_id=None) -> Signal:
    if isinstance(engine_or_connection, Connection):
        return engine_or_connection.message_to_factor
    return get_connection(engine_or_connection, variable_id, factor_id).message_to_factor


def get_connected_variable_ids(engine: InferenceEngine, factor_id: int):
    return backend_get_connected_variable_ids(engine.model_engine, factor_id)


def get_connected_factor_ids(engine: InferenceEngine, variable_id: int):
    return backend_get_connected_factor_ids(engine.model_engine, variable_id)


def create_inference_signal(engine: InferenceEngine) -> Signal:
    """create_inference_signal(), src/inference_signal.jl:140-142 (signals live in the engine's store here)."""
    return engine.store.Signal()


def set_variable_families(engine: InferenceEngine, variable_ids, families) -> None:
    """Value type of the signals of each variable in a model that mixes value types (``CXB_FAMILY_*``). In the
    reference the type travels with the Julia value (NormalMeanPrecision / Gamma / Float64, test/runtests.jl:52-99)."""
    ids = np.ascontiguousarray(_as_ids(variable_ids), dtype=np.int64)
    fam = np.ascontiguousarray(np.broadcast_to(np.asarray(families, dtype=np.int32), ids.shape))
    engine.store.check(engine.api.set_variable_families(engine.store.h, len(ids), ids.ctypes.data_as(capi.i64p),
                                                        fam.ctypes.data_as(capi.i32p)))


def link_signal_to_variable(variable: Variable, signal: Signal) -> None:
    """link_signal_to_variable!(variable, signal), src/model_engine.jl:80-83."""
    variable.linked_signals.append(signal)
    eng = getattr(variable, "_engine", None)
    if eng is None:
        raise RuntimeError("link_signal_to_variable!: the variable is not bound to an InferenceEngine yet")
    eng._links.setdefault(variable._id, []).append(signal)  # views of a lazy model engine are rebuilt from this
    eng.store.check(eng.api.link_signal(eng.store.h, variable._id, signal.sid))


def _as_ids(variable_id_or_ids) -> Tuple[int, ...]:
    if isinstance(variable_id_or_ids, (list, tuple, np.ndarray)):
        return tuple(int(v) for v in variable_id_or_ids)
    return (int(variable_id_or_ids),)


def request_inference_for(engine: InferenceEngine, variable_id_or_ids) -> InferenceRequest:
    """request_inference_for(engine, ids), src/inference_engine.jl:294-323."""
    ids = _as_ids(variable_id_or_ids)
    arr, p = _ids(ids)
    engine.store.check(engine.api.request_inference(engine.store.h, len(ids), p))
    return InferenceRequest(engine, ids, [get_variable(engine, v).marginal for v in ids])


def scan_inference_request(request: InferenceRequest, order: str = "dfs") -> List[Signal]:
    """scan_inference_request(request), src/inference_engine.jl:540-546.

    ``order="dfs"`` (default): the reference's own order - the depth-first visit order of ``process_dependencies!``,
    duplicates included (the sequential device traversal, ``cxb_scan_dfs``);
    ``order="id"``: the same set in ascending signal id, de-duplicated (the breadth-first frontier scan).
    """
    eng = request.engine
    fn = eng.api.scan if order == "id" else eng.api.scan_dfs
    n = fn(eng.store.h, None, 0)
    if n < 0:
        eng.store.check(capi.ERR_STATE)
    buf = np.zeros(max(n, 1), dtype=np.int64)
    # scanning is idempotent (is_pending only caches), so the sizing call above is harmless
    n = fn(eng.store.h, buf.ctypes.data_as(capi.i64p), n)
    return [Signal(eng.store, s) for s in buf[:n]]


SCHEDULES = {"auto": capi.SCHEDULE_AUTO, "lvl": capi.SCHEDULE_LEVEL, "seq": capi.SCHEDULE_SEQUENTIAL}


def update_marginals(engine: InferenceEngine, variable_id_or_ids, schedule: str = "auto") -> capi.UpdateStats:
    """update_marginals!(engine, ids), src/inference_engine.jl:553-632.

    ``schedule`` (``cxb_set_schedule``): ``"auto"`` (default) always answers as the reference does - the level-synchronous
    frontier schedule where its contract holds (graphs wired by the built-in resolvers; memoised / closed-form plans when
    the request was seen before), the literal sequential loop on the device for hand-wired graphs and for requests the
    level schedule refuses; ``"lvl"``: the level schedule alone (``OutOfContractError`` when it would differ from the
    reference, the engine left untouched); ``"seq"``: the sequential loop alone.
    Returns the per-call statistics (the reference returns ``nothing``).
    """
    prepared = variable_id_or_ids if isinstance(variable_id_or_ids, PreparedRequest) else None
    ids = prepared.variable_ids if prepared else _as_ids(variable_id_or_ids)
    stats = capi.UpdateStats()
    engine.store.check(engine.api.set_schedule(engine.store.h, SCHEDULES[schedule]))
    engine._callback_error = None
    before = _snapshot(engine) if engine.tracer is not None else None
    t0 = time.perf_counter_ns()
    if prepared:
        status = engine.api.update_marginals_prepared(engine.store.h, prepared.handle, C.byref(stats))
    else:
        arr, p = _ids(ids)
        status = engine.api.update_marginals(engine.store.h, len(ids), p, C.byref(stats))
    t1 = time.perf_counter_ns()
    if status != capi.OK and getattr(engine, "_callback_error", None) is not None:
        raise engine._callback_error
    engine.store.check(status)
    if engine.tracer is not None:
        engine.tracer.inference_requests.append(_collect_trace(engine, ids, t1 - t0, before))
    return stats


class PreparedSignals:
    """Handle of ``prepare_signals`` (``cxb_prepare_signals``)."""

    def __init__(self, engine, handle, n):
        self.engine, self.handle, self.n = engine, int(handle), int(n)


class PreparedRequest:
    """Handle of ``prepare_request`` (``cxb_prepare_request``): the ids of an inference request, validated once."""

    def __init__(self, engine, handle, ids):
        self.engine, self.handle, self.variable_ids = engine, int(handle), tuple(ids)


def engine_np_dtype(engine: InferenceEngine):
    return np.float32 if engine.store.dtype == capi.F32 else np.float64


def prepare_signals(engine: InferenceEngine, signals: Sequence[Signal]) -> PreparedSignals:
    """Validate and upload a list of signals once; ``set_values_prepared`` then sets all of them per call without any per-call
    host work on ids. The signals must be distinct and must not depend on each other (e.g. all observations)."""
    arr, p = _ids([s.sid for s in signals])
    h = engine.api.prepare_signals(engine.store.h, len(arr), p)
    if h < 0:
        engine.store.check(capi.ERR_BAD_ARG)
    return PreparedSignals(engine, h, len(arr))


def set_values_prepared(prepared: PreparedSignals, values, device_pointer: Optional[int] = None) -> None:
    """Bulk ``set_value!`` of a prepared list. ``values``: array ``[n][value_dim]`` (converted to the engine dtype if needed);
    or pass ``device_pointer`` (engine dtype, same layout, device memory): the call is then asynchronous on the engine's
    stream and nothing crosses the host."""
    e = prepared.engine
    if device_pointer is not None:
        e.store.check(e.api.set_values_prepared(e.store.h, prepared.handle, C.c_void_p(int(device_pointer)), 1))
        return
    v = np.ascontiguousarray(np.asarray(values).reshape(prepared.n, -1), dtype=engine_np_dtype(e) if e.api.is_device else np.float64)
    if v.shape[1] != e.store.value_dim:
        w = np.zeros((prepared.n, e.store.value_dim), dtype=v.dtype)
        w[:, : v.shape[1]] = v
        v = w
    e.store.check(e.api.set_values_prepared(e.store.h, prepared.handle, v.ctypes.data_as(C.c_void_p), 0))


def get_values_prepared(prepared: PreparedSignals, device_pointer: Optional[int] = None):
    """Bulk ``get_value`` of a prepared list in the engine dtype: returns ``[n][value_dim]`` (or fills ``device_pointer``)."""
    e = prepared.engine
    if device_pointer is not None:
        e.store.check(e.api.get_values_prepared(e.store.h, prepared.handle, C.c_void_p(int(device_pointer)), 1))
        return None
    out = np.zeros((prepared.n, e.store.value_dim), dtype=engine_np_dtype(e) if e.api.is_device else np.float64)
    e.store.check(e.api.get_values_prepared(e.store.h, prepared.handle, out.ctypes.data_as(C.c_void_p), 0))
    return out


def prepare_request(engine: InferenceEngine, variable_id_or_ids) -> PreparedRequest:
    """The ids of a request, validated once (``update_marginals(engine, prepared_request)``): the counterpart of keeping the
    reference's ``InferenceRequest`` object (src/inference_engine.jl:265-323) around."""
    ids = _as_ids(variable_id_or_ids)
    arr, p = _ids(ids)
    h = engine.api.prepare_request(engine.store.h, len(ids), p)
    if h < 0:
        engine.store.check(capi.ERR_BAD_ARG)
    return PreparedRequest(engine, h, ids)


def last_schedule(engine: InferenceEngine) -> int:
    """Which path answered the last ``update_marginals``: ``capi.SCHEDULE_LEVEL`` / ``SCHEDULE_SEQUENTIAL`` / ``RAN_REPLAY`` /
    ``RAN_PLAN`` (``cxb_last_schedule``)."""
    return int(engine.api.last_schedule(engine.store.h))


def _snapshot(engine: InferenceEngine):
    """Values and computed flags of every signal before a traced request (tracing is a debugging aid, as in the
    reference: src/inference_engine.jl:650-657 records `value_before_execution` per execution)."""
    n = engine.store.n_signals()
    if n == 0:
        return np.zeros((0, engine.store.value_dim)), np.zeros(0, dtype=bool)
    sigs = [Signal(engine.store, i) for i in range(n)]
    computed = np.array([is_computed(s) for s in sigs], dtype=bool)
    return get_values(sigs), computed


def _collect_trace(engine: InferenceEngine, ids, total_ns: int, before=None) -> TracedInferenceRequest:
    n = engine.api.trace_get(engine.store.h, None, None, 0)
    lv = np.zeros(max(n, 1), dtype=np.int64)
    sg = np.zeros(max(n, 1), dtype=np.int64)
    engine.api.trace_get(engine.store.h, lv.ctypes.data_as(capi.i64p), sg.ctypes.data_as(capi.i64p), n)
    ns = np.zeros(max(n, 1), dtype=np.int64)  # measured per execution (oracle) / per level, shared evenly (device)
    engine.api.trace_get_times(engine.store.h, ns.ctypes.data_as(capi.i64p), n)
    var = np.zeros(max(n, 1), dtype=np.int64)
    if engine.api.trace_get_variables(engine.store.h, var.ctypes.data_as(capi.i64p), n) != n:
        var = None
    rounds: List[TracedInferenceRound] = []
    cur_key, cur = None, None
    seen_after: Dict[int, Any] = {}
    for k in range(n):
        key = int(lv[k]) if lv[k] >= 0 else -1  # the final phase (marginals, then linked signals) is ONE round, :610-628
        if key != cur_key:
            cur = TracedInferenceRound(engine, 0, [])
            rounds.append(cur)  # only rounds with >= 1 execution are recorded, :818
            cur_key = key
        s = Signal(engine.store, int(sg[k]))
        variant = get_variant(s)
        vid = int(var[k]) if var is not None else getattr(variant, "variable_id", None)
        value_before = None
        if before is not None:  # first execution of the signal in this request: its old value is the snapshot's
            vals, computed = before
            if s.sid in seen_after:  # executed before in this request (sequential schedule only): the value it left then
                value_before = seen_after[s.sid]
            elif not computed[s.sid]:
                value_before = UndefValue()
            else:
                value_before = float(vals[s.sid][0]) if engine.store.value_dim == 1 else vals[s.sid].copy()
        value_after = get_value(s)  # of the signal's LAST execution in the request
        cur.executions.append(TracedInferenceExecution(engine, vid, s, max(int(ns[k]), 1), value_before, value_after))
        seen_after[s.sid] = value_after
        cur.total_time_in_ns += max(int(ns[k]), 1)
    return TracedInferenceRequest(engine, max(total_ns, 1), InferenceRequest(engine, tuple(ids), []), rounds)
