"""Model-engine interface — host-side mirror of ``src/model_engine.jl`` and of the
``BipartiteFactorGraphsExt`` backend (``ext/BipartiteFactorGraphsExt/BipartiteFactorGraphsExt.jl``).

``BipartiteFactorGraph`` below plays the role of BipartiteFactorGraphs.jl (third party, not vendored
by the reference): one shared id space for variables and factors, ids handed out in creation order
(0-based here, 1-based in Julia), neighbours iterated in ascending id order.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional, Tuple

from .inference_signal import Signal


@dataclass
class Variable:  # src/model_engine.jl:30-35
    name: str
    index: Any = None
    marginal: Optional[Signal] = None  # bound when the InferenceEngine is built
    linked_signals: List[Signal] = field(default_factory=list)

    def __repr__(self):  # src/model_engine.jl:85-91
        return f"Variable(name = {self.name}" + (f", index = {self.index})" if self.index is not None else ")")


@dataclass
class Factor:  # src/model_engine.jl:119-122
    functional_form: Any
    local_marginals: List[Signal] = field(default_factory=list)

    def __repr__(self):
        return f"Factor(functional_form = {self.functional_form})"


@dataclass
class Connection:  # src/model_engine.jl:181-186
    label: str
    index: int = 0
    message_to_variable: Optional[Signal] = None
    message_to_factor: Optional[Signal] = None

    def __repr__(self):
        return f"Connection(label = {self.label}" + (f", index = {self.index})" if self.index else ")")


def get_variable_name(v: Variable):
    return v.name


def get_variable_index(v: Variable):
    return v.index


def get_variable_marginal(v: Variable) -> Signal:
    return v.marginal


def get_variable_linked_signals(v: Variable) -> List[Signal]:
    return v.linked_signals


def get_factor_functional_form(f: Factor):
    return f.functional_form


def get_factor_local_marginals(f: Factor):
    return f.local_marginals


def add_local_marginal_to_factor(f: Factor, s: Signal) -> None:
    f.local_marginals.append(s)


def get_connection_label(c: Connection):
    return c.label


def get_connection_index(c: Connection):
    return c.index


class UnsupportedModelEngineError(Exception):  # src/model_engine.jl:252-266
    def __init__(self, model_engine, missing_function=None):
        self.model_engine = model_engine
        self.missing_function = missing_function
        if missing_function is None:
            msg = f"The model engine of type `{type(model_engine).__name__}` is not supported."
        else:
            msg = (f"The model engine of type `{type(model_engine).__name__}` does not implement the function "
                   f"`{missing_function}`.")
        super().__init__(msg)


class SupportedModelEngine:
    pass


class UnsupportedModelEngine:
    pass


def is_engine_supported(engine) -> Any:  # trait, src/model_engine.jl:310
    fn = getattr(engine, "is_engine_supported", None)
    return fn() if fn else UnsupportedModelEngine()


def throw_if_engine_unsupported(engine):  # src/model_engine.jl:319-321
    if not isinstance(is_engine_supported(engine), SupportedModelEngine):
        raise UnsupportedModelEngineError(engine, None)
    return engine


def _generic(name):
    def call(engine, *args):
        fn = getattr(engine, name, None)
        if fn is None:
            raise UnsupportedModelEngineError(engine, name)  # src/model_engine.jl:329-391
        return fn(*args)

    call.__name__ = name
    return call


backend_get_variable = _generic("get_variable")
backend_get_factor = _generic("get_factor")
backend_get_variable_ids = _generic("get_variable_ids")
backend_get_factor_ids = _generic("get_factor_ids")
backend_get_connection = _generic("get_connection")
backend_get_connected_variable_ids = _generic("get_connected_variable_ids")
backend_get_connected_factor_ids = _generic("get_connected_factor_ids")


class BipartiteFactorGraph:
    """Stand-in for ``BipartiteFactorGraph{Variable,Factor,Connection}``; implements the 7 generics."""

    def __init__(self):
        self._variables: Dict[int, Variable] = {}
        self._factors: Dict[int, Factor] = {}
        self._edges: Dict[Tuple[int, int], Connection] = {}
        self._edge_order: List[Tuple[int, int]] = []
        self._nbr: Dict[int, List[int]] = {}
        self._next = 0

    def is_engine_supported(self):  # ext/...Ext.jl:18-20
        return SupportedModelEngine()

    def add_variable(self, v: Variable) -> int:
        i = self._next
        self._next += 1
        self._variables[i] = v
        self._nbr[i] = []
        return i

    def add_factor(self, f: Factor) -> int:
        i = self._next
        self._next += 1
        self._factors[i] = f
        self._nbr[i] = []
        return i

    def add_edge(self, variable_id: int, factor_id: int, c: Connection) -> None:
        if variable_id not in self._variables or factor_id not in self._factors:
            raise KeyError("add_edge!: unknown variable or factor id")
        if (variable_id, factor_id) in self._edges:
            raise ValueError("add_edge!: duplicate edge")
        self._edges[(variable_id, factor_id)] = c
        self._edge_order.append((variable_id, factor_id))
        self._nbr[variable_id].append(factor_id)
        self._nbr[variable_id].sort()
        self._nbr[factor_id].append(variable_id)
        self._nbr[factor_id].sort()

    # the 7 generics, ext/...Ext.jl:22-48
    def get_variable(self, variable_id: int) -> Variable:
        return self._variables[variable_id]

    def get_factor(self, factor_id: int) -> Factor:
        return self._factors[factor_id]

    def get_variable_ids(self):
        return sorted(self._variables)

    def get_factor_ids(self):
        return sorted(self._factors)

    def get_connection(self, variable_id: int, factor_id: int) -> Connection:
        return self._edges[(variable_id, factor_id)]

    def get_connected_variable_ids(self, factor_id: int):
        return list(self._nbr[factor_id])

    def get_connected_factor_ids(self, variable_id: int):
        return list(self._nbr[variable_id])

    # helpers for the engine constructor
    def n_ids(self) -> int:
        return self._next

    def edges(self):
        return list(self._edge_order)


def add_variable(graph: BipartiteFactorGraph, v: Variable) -> int:
    return graph.add_variable(v)


def add_factor(graph: BipartiteFactorGraph, f: Factor) -> int:
    return graph.add_factor(f)


def add_edge(graph: BipartiteFactorGraph, variable_id: int, factor_id: int, c: Connection) -> None:
    graph.add_edge(variable_id, factor_id, c)


class B200ModelEngine:
    """A model-engine BACKEND whose signals live on the device (mirror of ``julia/CortexB200.jl::B200ModelEngine``, SURVEY
    8b / 8f.3): implements the trait and the seven generics of ``src/model_engine.jl:269-391`` LAZILY. The graph is handed to
    the library as flat arrays; nothing is allocated per variable, factor or connection up front — ``get_variable`` /
    ``get_connection`` materialise a ``Variable`` / ``Connection`` view on demand whose signals are references into the
    device state, so the reference-style read-out

        get_value(get_variable_marginal(get_variable(engine, v)))

    fetches the current value from the device and no signal id ever shows in user code.

    ``B200ModelEngine(n_ids, is_factor, functional_forms, edge_variable, edge_factor)`` takes the arrays directly (graphs that
    never exist as host objects); ``B200ModelEngine.from_engine(source)`` walks any supported model engine ONCE.
    """

    def __init__(self, n_ids, is_factor, functional_forms, edge_variable, edge_factor, names=None, labels=None):
        import numpy as np

        self.n_ids = int(n_ids)
        self.is_factor = np.ascontiguousarray(is_factor, dtype=np.uint8)
        if self.is_factor.shape != (self.n_ids,):
            raise ValueError("is_factor must have one entry per id")
        self.functional_forms = functional_forms  # sequence / dict: factor id -> Factor.functional_form
        self.edge_variable = np.ascontiguousarray(edge_variable, dtype=np.int64)
        self.edge_factor = np.ascontiguousarray(edge_factor, dtype=np.int64)
        if self.edge_variable.shape != self.edge_factor.shape:
            raise ValueError("edge arrays differ in length")
        self._names = names or {}
        self._labels = labels or {}
        # adjacency in CSR form, neighbours in ascending id (the iteration-order contract of ext/BipartiteFactorGraphsExt)
        deg = np.bincount(np.concatenate([self.edge_variable, self.edge_factor]), minlength=self.n_ids)
        self._adj_off = np.concatenate([[0], np.cumsum(deg)])
        ends = np.concatenate([self.edge_variable, self.edge_factor])
        others = np.concatenate([self.edge_factor, self.edge_variable])
        order = np.lexsort((others, ends))
        self._adj = others[order]
        self._engine = None  # bound by InferenceEngine.__init__

    @classmethod
    def from_engine(cls, source):
        throw_if_engine_unsupported(source)
        vids = sorted(int(v) for v in backend_get_variable_ids(source))
        fids = sorted(int(f) for f in backend_get_factor_ids(source))
        n_ids = (max(vids + fids) + 1) if (vids or fids) else 0
        is_factor = [0] * n_ids
        forms = {}
        for f in fids:
            is_factor[f] = 1
            forms[f] = backend_get_factor(source, f).functional_form
        names = {v: (backend_get_variable(source, v).name, backend_get_variable(source, v).index) for v in vids}
        edges = source.edges() if hasattr(source, "edges") else [(int(v), f) for f in fids for v in backend_get_connected_variable_ids(source, f)]
        labels = {}
        for (v, f) in edges:
            c = backend_get_connection(source, v, f)
            labels[(v, f)] = (c.label, c.index)
        return cls(n_ids, is_factor, forms, [e[0] for e in edges], [e[1] for e in edges], names=names, labels=labels)

    def is_engine_supported(self):
        return SupportedModelEngine()

    # ---- the seven generics, src/model_engine.jl:329-391 -------------------------------------------------------------------
    def get_variable_ids(self):
        import numpy as np

        return [int(i) for i in np.flatnonzero(self.is_factor == 0)]

    def get_factor_ids(self):
        import numpy as np

        return [int(i) for i in np.flatnonzero(self.is_factor == 1)]

    def _neighbours(self, i):
        return [int(x) for x in self._adj[self._adj_off[i]:self._adj_off[i + 1]]]

    def get_connected_variable_ids(self, factor_id: int):
        self._check(factor_id, factor=True)
        return self._neighbours(factor_id)

    def get_connected_factor_ids(self, variable_id: int):
        self._check(variable_id, factor=False)
        return self._neighbours(variable_id)

    def _check(self, i, factor):
        if not (0 <= int(i) < self.n_ids) or bool(self.is_factor[int(i)]) != factor:
            raise KeyError(f"not a {'factor' if factor else 'variable'} id: {i}")

    def _signal(self, kind, v, f=-1):
        eng = self._engine
        if eng is None:
            raise RuntimeError("the model engine is not bound to an InferenceEngine yet")
        sid = eng.api.signal_id(eng.store.h, kind, int(v), int(f))
        if sid < 0:
            raise KeyError(f"no such signal: kind {kind}, variable {v}, factor {f}")
        return Signal(eng.store, sid)

    def get_variable(self, variable_id: int) -> Variable:
        from . import _capi as capi

        self._check(variable_id, factor=False)
        name, index = self._names.get(int(variable_id), ("variable", int(variable_id)))
        var = Variable(name=name, index=index, marginal=self._signal(capi.KIND_MARGINAL, variable_id),
                       linked_signals=list(self._engine._links.get(int(variable_id), ())))
        var._engine, var._id = self._engine, int(variable_id)
        return var

    def get_factor(self, factor_id: int) -> Factor:
        self._check(factor_id, factor=True)
        forms = self.functional_forms
        return Factor(functional_form=forms[int(factor_id)])

    def get_connection(self, variable_id: int, factor_id: int) -> Connection:
        from . import _capi as capi

        label, index = self._labels.get((int(variable_id), int(factor_id)), ("edge", 0))
        return Connection(label=label, index=index, message_to_variable=self._signal(capi.KIND_M2V, variable_id, factor_id),
                          message_to_factor=self._signal(capi.KIND_M2F, variable_id, factor_id))

    def edges(self):
        return list(zip((int(v) for v in self.edge_variable), (int(f) for f in self.edge_factor)))
