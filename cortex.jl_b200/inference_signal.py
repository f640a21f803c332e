"""Signals and their variants — host-side mirror of ``src/signal.jl`` and ``src/inference_signal.jl``.

A ``Signal`` here is a *reference* (engine handle + dense signal id): all state — value,
``(is_potentially_pending, is_pending)`` props, the 4-bit per-dependency nibbles, dependency and
listener lists — lives in the engine behind the C ABI (device memory for the product library).
Function names follow the reference (``set_value!`` -> ``set_value`` ...).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _capi as capi


class CortexError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(message)
        self.status = status


class NotPendingError(ValueError):
    """ArgumentError of compute! on a non-pending signal, src/signal.jl:399-405."""


class NoRuleError(RuntimeError):
    """error(...) of an unimplemented compute_* rule, src/inference_engine.jl:358-360."""


class OutOfContractError(RuntimeError):
    """The level-synchronous schedule would not reproduce the reference order."""


def raise_for_status(status: int, message: str):
    if status == capi.OK:
        return
    if status == capi.ERR_NOT_PENDING:
        raise NotPendingError(message)
    if status == capi.ERR_NO_RULE:
        raise NoRuleError(message)
    if status == capi.ERR_OUT_OF_CONTRACT:
        raise OutOfContractError(message)
    if status == capi.ERR_BAD_ARG:
        raise ValueError(message)
    raise CortexError(status, message)


# ---- variants: src/inference_signal.jl:16-96 ---------------------------------------------------
@dataclass(frozen=True)
class Unspecified:
    pass


@dataclass(frozen=True)
class MessageToFactor:
    variable_id: int
    factor_id: int


@dataclass(frozen=True)
class MessageToVariable:
    variable_id: int
    factor_id: int


@dataclass(frozen=True)
class ProductOfMessages:
    variable_id: int
    range: Tuple[int, int]  # inclusive 0-based positions into factors_connected_to_variable
    factors_connected_to_variable: Tuple[int, ...] = ()


@dataclass(frozen=True)
class IndividualMarginal:
    variable_id: int


@dataclass(frozen=True)
class JointMarginal:
    factor_id: int
    variable_ids: Tuple[int, ...] = ()


class SignalStore:
    """Owner of one engine handle; the thing every ``Signal`` reference points into."""

    def __init__(self, api: capi.CApi, value_dim: int = 1, family: int = capi.FAMILY_SUM, dtype: int = capi.F64,
                 device: int = 0):
        self.api = api
        self.value_dim = int(value_dim)
        self.family = family
        self.dtype = dtype
        h = C.c_void_p()
        st = api.create(device, dtype, self.value_dim, family, C.byref(h))
        if st != capi.OK or not h:
            raise CortexError(st, "cxb_create failed (is a CUDA device present? there is no CPU fallback)")
        self.h = h
        self._neighbours = {}
        self._joint_vars = {}

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.api.destroy(self.h)
                self.h = None
        except Exception:
            pass

    def check(self, status: int):
        if status != capi.OK:
            msg = self.api.last_error(self.h)
            raise_for_status(status, msg.decode() if msg else f"status {status}")

    # Signal() / Signal(value), src/signal.jl:107-114 ; create_inference_signal, src/inference_signal.jl:140
    def Signal(self, value=None) -> "Signal":
        sid = self.api.create_signal(self.h)
        if sid < 0:
            self.check(capi.ERR_STATE)
        s = Signal(self, sid)
        if value is not None:
            set_value(s, value)
        return s

    def n_signals(self) -> int:
        return int(self.api.n_signals(self.h))


class Signal:
    """Reference to one signal (src/signal.jl:82-115). Equality = identity of (store, id) (``===``)."""

    __slots__ = ("store", "sid")

    def __init__(self, store: SignalStore, sid: int):
        self.store = store
        self.sid = int(sid)

    def __eq__(self, other):
        return isinstance(other, Signal) and other.store is self.store and other.sid == self.sid

    def __hash__(self):
        return hash((id(self.store), self.sid))

    def __repr__(self):  # Base.show, src/signal.jl:360-370
        val = repr(get_value(self)) if is_computed(self) else "#undef"
        return f"Signal(value={val}, pending={'true' if is_pending(self) else 'false'}, variant={get_variant(self)!r})"


def _ids(seq: Sequence[int]):
    arr = np.ascontiguousarray(np.asarray(seq, dtype=np.int64))
    return arr, arr.ctypes.data_as(capi.i64p)


def _as_value(store: SignalStore, value) -> np.ndarray:
    v = np.zeros(store.value_dim, dtype=np.float64)
    a = np.atleast_1d(np.asarray(value, dtype=np.float64)).ravel()
    if a.size > store.value_dim:
        raise ValueError(f"value has {a.size} components, engine value_dim is {store.value_dim}")  # typed signals throw, test/signal_tests.jl:19-20
    v[: a.size] = a
    return v


def set_value(signal: Signal, value) -> None:
    """set_value!(signal, value), src/signal.jl:232-253."""
    st = signal.store
    v = _as_value(st, value)
    arr, p = _ids([signal.sid])
    st.check(st.api.set_values(st.h, 1, p, v.ctypes.data_as(capi.f64p), st.value_dim))


def set_values(signals: Sequence[Signal], values) -> None:
    """Bulk set_value! (one H2D copy + one notification kernel on the device engine)."""
    if not signals:
        return
    st = signals[0].store
    vals = np.zeros((len(signals), st.value_dim), dtype=np.float64)
    a = np.asarray(values, dtype=np.float64).reshape(len(signals), -1)
    vals[:, : a.shape[1]] = a
    arr, p = _ids([s.sid for s in signals])
    st.check(st.api.set_values(st.h, len(signals), p, vals.ctypes.data_as(capi.f64p), st.value_dim))


class UndefValue:
    """UndefValue(), src/signal.jl:9-14: the value of a signal that has never been computed or set."""
    _instance = None

    def __new__(cls):
        if cls._instance is None:
            cls._instance = super().__new__(cls)
        return cls._instance

    def __eq__(self, other):
        return isinstance(other, UndefValue)

    def __hash__(self):
        return hash("UndefValue")

    def __repr__(self):
        return "UndefValue()"


def get_values(signals: Sequence[Signal]) -> np.ndarray:
    st = signals[0].store
    out = np.zeros((len(signals), st.value_dim), dtype=np.float64)
    arr, p = _ids([s.sid for s in signals])
    st.check(st.api.get_values(st.h, len(signals), p, out.ctypes.data_as(capi.f64p), st.value_dim))
    return out


def get_value(signal: Signal):
    """get_value(signal), src/signal.jl:171. Scalars come back as float, vectors as ndarray."""
    v = get_values([signal])[0]
    return float(v[0]) if signal.store.value_dim == 1 else v


def is_pending(signal: Signal) -> bool:
    r = signal.store.api.is_pending(signal.store.h, signal.sid)
    if r < 0:
        signal.store.check(capi.ERR_BAD_ARG)
    return bool(r)


def is_computed(signal: Signal) -> bool:
    r = signal.store.api.is_computed(signal.store.h, signal.sid)
    if r < 0:
        signal.store.check(capi.ERR_BAD_ARG)
    return bool(r)


def add_dependency(signal: Signal, dependency: Signal, *, weak: bool = False, listen: bool = True,
                   check_computed: bool = True, intermediate: bool = False) -> None:
    """add_dependency!(signal, dependency; weak, listen, check_computed, intermediate), src/signal.jl:286-337."""
    if signal.store is not dependency.store:
        raise TypeError("signals of different engines cannot depend on each other")  # test/signal_tests.jl:137-162
    flags = ((capi.DEP_WEAK if weak else 0) | (0 if listen else capi.DEP_NO_LISTEN)
             | (0 if check_computed else capi.DEP_NO_CHECK_COMPUTED) | (capi.DEP_INTERMEDIATE if intermediate else 0))
    st = signal.store
    st.check(st.api.add_dependency(st.h, signal.sid, dependency.sid, flags))


def compute(signal: Signal, *, force: bool = False, skip_if_no_listeners: bool = False) -> None:
    """compute!(strategy, signal; force, skip_if_no_listeners), src/signal.jl:392-410.

    The strategy is the rule registered with the engine (kernel by factor type / value family)."""
    st = signal.store
    st.check(st.api.compute(st.h, signal.sid, int(force), int(skip_if_no_listeners)))


def process_dependencies(f, signal: Signal, *, retry: bool = False, visited: Optional[list] = None) -> bool:
    """process_dependencies!(f, signal; retry), src/signal.jl:466-490, run on the device in the reference's depth-first
    order (``cxb_process_dependencies_table``).

    ``f`` is ``None`` (the scanner's callback: ``is_pending``, with its caching side effect) or a PURE predicate on
    signals: it is tabulated over every signal of the engine before the traversal starts and the traversal reads the
    table. The visit sequence (every call of ``f``, retries included) is appended to ``visited``."""
    st = signal.store
    n = st.n_signals()
    table = None
    if f is not None:
        table = np.ascontiguousarray([1 if f(Signal(st, i)) else 0 for i in range(n)], dtype=np.uint8)
    cap = 8 * max(n, 1) + 1024
    out = np.zeros(cap, dtype=np.int64)
    processed = C.c_int32(0)
    cnt = st.api.process_dependencies_table(st.h, signal.sid, int(retry), table.ctypes.data_as(capi.u8p) if table is not None else None,
                                            out.ctypes.data_as(capi.i64p), cap, C.byref(processed))
    if cnt < 0:
        st.check(capi.ERR_BAD_ARG)
    if visited is not None:
        visited.extend(Signal(st, int(i)) for i in out[:min(cnt, cap)])
    return bool(processed.value)


def _list(fn, signal: Signal):
    st = signal.store
    n = fn(st.h, signal.sid, None, None, 0)
    if n < 0:
        st.check(capi.ERR_BAD_ARG)
    ids = np.zeros(max(n, 1), dtype=np.int64)
    aux = np.zeros(max(n, 1), dtype=np.uint8)
    fn(st.h, signal.sid, ids.ctypes.data_as(capi.i64p), aux.ctypes.data_as(capi.u8p), n)
    return ids[:n], aux[:n]


def get_dependencies(signal: Signal) -> List[Signal]:
    ids, _ = _list(signal.store.api.get_dependencies, signal)
    return [Signal(signal.store, i) for i in ids]


def get_dependency_props(signal: Signal) -> List[int]:
    """The 4-bit (F C W I) nibble of every dependency, src/signal.jl:17-45, 507-526."""
    _, nib = _list(signal.store.api.get_dependencies, signal)
    return [int(x) for x in nib]


def get_listeners(signal: Signal) -> List[Signal]:
    ids, _ = _list(signal.store.api.get_listeners, signal)
    return [Signal(signal.store, i) for i in ids]


def get_listenmask(signal: Signal) -> List[bool]:
    _, m = _list(signal.store.api.get_listeners, signal)
    return [bool(x) for x in m]


def get_variant(signal: Signal):
    """get_variant(signal), src/signal.jl:180 -> one of the InferenceSignalVariants."""
    st = signal.store
    out = (C.c_int64 * 5)()
    st.check(st.api.signal_info(st.h, signal.sid, out))
    kind, var, fac, r0, r1 = (int(x) for x in out)
    if kind == capi.KIND_M2F:
        return MessageToFactor(var, fac)
    if kind == capi.KIND_M2V:
        return MessageToVariable(var, fac)
    if kind == capi.KIND_MARGINAL:
        return IndividualMarginal(var)
    if kind == capi.KIND_PRODUCT:
        return ProductOfMessages(var, (r0, r1), tuple(st._neighbours.get(var, ())))
    if kind == capi.KIND_JOINT:
        return JointMarginal(fac, tuple(st._joint_vars.get(signal.sid, ())))
    return Unspecified()


def set_variant(signal: Signal, variant) -> None:
    """set_variant!(signal, variant), src/signal.jl:185-192."""
    st = signal.store
    if isinstance(variant, MessageToFactor):
        args = (capi.KIND_M2F, variant.variable_id, variant.factor_id)
    elif isinstance(variant, MessageToVariable):
        args = (capi.KIND_M2V, variant.variable_id, variant.factor_id)
    elif isinstance(variant, IndividualMarginal):
        args = (capi.KIND_MARGINAL, variant.variable_id, -1)
    elif isinstance(variant, JointMarginal):
        args = (capi.KIND_JOINT, -1, variant.factor_id)
        st._joint_vars[signal.sid] = tuple(variant.variable_ids)
    elif isinstance(variant, Unspecified):
        args = (capi.KIND_UNSPECIFIED, -1, -1)
    else:
        raise TypeError(f"set_variant!: unsupported variant {variant!r}")
    st.check(st.api.set_signal_variant(st.h, signal.sid, *args))


def isa_variant(signal: Signal, T) -> bool:
    return isinstance(get_variant(signal), T)
