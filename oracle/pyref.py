"""Second, structurally independent CPU witness of the reference's signal semantics (TEST INFRASTRUCTURE ONLY).

`oracle/cortex_oracle.cpp` restates Cortex.jl with integer signal ids and flat arrays. This file restates the same
functions the way the reference itself is built - one heap object per signal, dependencies and listeners as lists of
object references, the per-dependency nibbles in a list of 64-bit chunks addressed with the reference's 1-based
index arithmetic - so that the two restatements share no data structure and no code. `tests/test_pyref_witness.py`
drives both with the same random scripts and asserts identical observable state after every operation.

Pure-Python loops: small graphs only. Only `tests/` may import this module (same rule as everything under oracle/).

Reference: ReactiveBayes/Cortex.jl v0.3.0, `src/signal.jl` and `src/inference_engine.jl` (file:line on each function).
"""
from __future__ import annotations

M64 = (1 << 64) - 1
INTERMEDIATE, WEAK, COMPUTED, FRESH = 0x1, 0x2, 0x4, 0x8  # src/signal.jl:507-510 (single-nibble masks)
ALL_W, ALL_C, ALL_F = 0x2222_2222_2222_2222, 0x4444_4444_4444_4444, 0x8888_8888_8888_8888  # :512-515
PASS_TARGET = 0x1111_1111_1111_1111  # :519


class _Undef:
    def __repr__(self):
        return "#undef"


UNDEF = _Undef()  # UndefValue(), src/signal.jl:7


class DependenciesProps:
    """SignalDependenciesProps, src/signal.jl:36-45."""

    def __init__(self):
        self.length = 0
        self.chunks = [0]

    @staticmethod
    def offset(index):  # signal_dependencies_props_get_offset, :522-526 (index is 1-based)
        return (index - 1) // 16 + 1, ((index - 1) % 16) << 2

    def add(self):  # add_dependency!(props), :529-544
        self.length += 1
        if len(self.chunks) < (4 * self.length - 1) // 64 + 1:
            self.chunks.append(0)
        return self.length

    def test(self, index, mask):  # is_dependency, :546-549
        c, off = self.offset(index)
        return (self.chunks[c - 1] & ((mask << off) & M64)) != 0

    def set(self, index, mask):  # set_dependency!, :567-571
        c, off = self.offset(index)
        self.chunks[c - 1] |= (mask << off) & M64

    def unset_all_fresh(self):  # unset_all_dependencies!, :634-639, with the fresh mask (:653-655)
        for i in range(len(self.chunks)):
            self.chunks[i] &= ~ALL_F & M64

    def nibble(self, index):
        c, off = self.offset(index)
        return (self.chunks[c - 1] >> off) & 0xF

    def meets_pending_criteria(self):  # is_meeting_pending_criteria, :668-730
        if self.length == 0:
            return False
        for chunk in self.chunks[:-1]:
            w, c, f = (chunk & ALL_W) >> 1, (chunk & ALL_C) >> 2, (chunk & ALL_F) >> 3
            if (c & (w | f)) != PASS_TARGET:
                return False
        last_chunk, off = self.offset(self.length)
        chunk = self.chunks[last_chunk - 1] | ((M64 << (off + 4)) & M64)
        w, c, f = (chunk & ALL_W) >> 1, (chunk & ALL_C) >> 2, (chunk & ALL_F) >> 3
        return (c & (w | f)) == PASS_TARGET


class Signal:
    """Signal{D,V}, src/signal.jl:82-115."""

    def __init__(self, value=UNDEF, variant=None):
        self.value = value
        self.variant = variant
        self.potentially_pending = False
        self.pending = False
        self.dependencies_props = DependenciesProps()
        self.dependencies = []
        self.listenmask = []
        self.listeners = []


def is_pending(s):  # src/signal.jl:141-154
    if s.pending:
        return True
    if s.potentially_pending:
        new = s.dependencies_props.meets_pending_criteria()
        s.potentially_pending, s.pending = False, new
        return new
    return False


def is_computed(s):  # :162-164
    return s.value is not UNDEF


def notify_listener(listener, signal, update_potentially_pending=False):  # :339-356
    if update_potentially_pending:
        listener.potentially_pending, listener.pending = True, False
    for i, dependency in enumerate(listener.dependencies, start=1):
        if dependency is signal:
            listener.dependencies_props.set(i, FRESH)
            listener.dependencies_props.set(i, COMPUTED)
            break


def set_value(signal, value):  # :232-253
    signal.value = value
    signal.dependencies_props.unset_all_fresh()
    signal.potentially_pending, signal.pending = False, False
    for is_listening, listener in zip(signal.listenmask, signal.listeners):
        notify_listener(listener, signal, update_potentially_pending=is_listening)


def add_dependency(signal, dependency, weak=False, listen=True, check_computed=True, intermediate=False):  # :286-337
    if signal is dependency:
        return
    props = signal.dependencies_props
    index = props.add()
    if weak:
        props.set(index, WEAK)
    if intermediate:
        props.set(index, INTERMEDIATE)
    signal.dependencies.append(dependency)
    dependency.listenmask.append(bool(listen))
    dependency.listeners.append(signal)
    if check_computed and is_computed(dependency):
        props.set(index, COMPUTED)
        if not is_computed(signal):
            props.set(index, FRESH)
        signal.potentially_pending, signal.pending = True, False
    elif check_computed and not is_computed(dependency):
        signal.potentially_pending, signal.pending = False, False


def compute(strategy, signal, force=False, skip_if_no_listeners=False):  # :392-410
    if skip_if_no_listeners and not signal.listeners:
        return
    if not force and not is_pending(signal):
        raise ValueError("Signal is not pending. Cannot compute a non-pending signal.")
    set_value(signal, strategy(signal, signal.dependencies))


def process_dependencies(f, signal, retry=False):  # :466-490
    processed_at_least_once = False
    for i, dependency in enumerate(signal.dependencies, start=1):
        processed = f(dependency)
        if not processed:
            if signal.dependencies_props.test(i, INTERMEDIATE):
                intermediate_processed = process_dependencies(f, dependency, retry=retry)
                if intermediate_processed and retry:
                    processed = f(dependency)
                processed_at_least_once = processed_at_least_once or intermediate_processed
        processed_at_least_once = processed_at_least_once or processed
    return processed_at_least_once


def update_marginals(marginals, linked_signals, strategy):
    """update_marginals!(engine, ids), src/inference_engine.jl:559-632, with request_inference_for (:298-323) inlined.
    marginals[i] / linked_signals[i]: the marginal signal and the linked signals of the i-th requested variable;
    strategy(signal, dependencies) -> value is what process! (:479-509) dispatches to. Returns the executed signals."""
    executed = []
    for marginal, linked in zip(marginals, linked_signals):  # :305-318
        for dependency in marginal.dependencies:
            dependency.potentially_pending, dependency.pending = True, False
        for ls in linked:
            ls.potentially_pending, ls.pending = True, False
    ready = [False] * len(marginals)

    def f(dependency):  # process_inference_request, :512-525
        if is_pending(dependency):
            compute(strategy, dependency)
            executed.append(dependency)
            return True
        return False

    indices = list(range(len(marginals)))
    is_reverse = False
    should_continue = True
    while should_continue:  # :577-607
        cont = False
        for i in (reversed(indices) if is_reverse else indices):
            if not ready[i]:
                processed = process_dependencies(f, marginals[i], retry=True)
                if is_pending(marginals[i]):
                    ready[i] = True
                cont = cont or processed
        is_reverse = not is_reverse
        should_continue = cont
    for marginal, linked in zip(marginals, linked_signals):  # :610-628
        if is_pending(marginal):
            compute(strategy, marginal)
            executed.append(marginal)
        for ls in linked:
            if not is_pending(ls):
                continue
            compute(strategy, ls)
            executed.append(ls)
    return executed


# ---- DefaultDependencyResolver, src/dependencies.jl:5-173 --------------------------------------------------------------
class Model:
    """What the resolver needs of an InferenceEngine: ids, adjacency in backend iteration order, and the signals owned by
    variables (marginal) and connections (message_to_variable / message_to_factor), src/model_engine.jl:30-35,181-186."""

    def __init__(self, variable_ids, factor_ids, edges):
        self.variable_ids = list(variable_ids)
        self.factor_ids = list(factor_ids)
        self.factors_of = {v: [] for v in self.variable_ids}
        self.variables_of = {f: [] for f in self.factor_ids}
        self.marginal = {v: Signal(variant=("marginal", v)) for v in self.variable_ids}
        self.m2v, self.m2f = {}, {}
        self.products = []
        self.warnings = []
        for v, f in edges:
            self.factors_of[v].append(f)
            self.variables_of[f].append(v)
            self.m2v[(v, f)] = Signal(variant=("m2v", v, f))
            self.m2f[(v, f)] = Signal(variant=("m2f", v, f))

    def signals(self):
        return list(self.marginal.values()) + list(self.m2v.values()) + list(self.m2f.values()) + self.products


def resolve_dependencies(model):  # src/dependencies.jl:7-15
    for f in model.factor_ids:
        resolve_factor_dependencies(model, f)
    for v in model.variable_ids:
        resolve_variable_dependencies(model, v)


def resolve_factor_dependencies(model, f):  # :17-31
    vs = model.variables_of[f]
    for v1 in vs:
        for v2 in vs:
            if v1 != v2:
                add_dependency(model.m2v[(v1, f)], model.m2f[(v2, f)])


def resolve_variable_dependencies(model, v):  # :33-126
    fs = model.factors_of[v]
    marginal = model.marginal[v]
    n = len(fs)
    if n == 0:
        model.warnings.append(("Variable has no connected factors", v))  # :40-43
        return
    if n < 2:
        add_dependency(marginal, model.m2v[(v, fs[0])], intermediate=True)
        return
    if n <= 5:
        for f in fs:
            add_dependency(marginal, model.m2v[(v, f)], intermediate=True)
            to_factor = model.m2f[(v, f)]
            if to_factor.listeners:
                for other in fs:
                    if other != f:
                        add_dependency(to_factor, model.m2v[(v, other)], intermediate=True)
        return
    middle = n // 2
    left_range, right_range = range(1, middle + 1), range(middle + 1, n + 1)  # 1-based, inclusive like the reference
    left = form_segment_tree_dependency(model, left_range, fs, v)
    right = form_segment_tree_dependency(model, right_range, fs, v)
    for i in left_range:
        to_factor = model.m2f[(v, fs[i - 1])]
        if to_factor.listeners:
            add_dependency(to_factor, right, intermediate=True)
    for i in right_range:
        to_factor = model.m2f[(v, fs[i - 1])]
        if to_factor.listeners:
            add_dependency(to_factor, left, intermediate=True)
    add_dependency(marginal, left, intermediate=True)
    add_dependency(marginal, right, intermediate=True)


def form_segment_tree_dependency(model, rng, fs, v):  # :128-173
    assert len(rng) >= 1
    if len(rng) == 1:
        return model.m2v[(v, fs[rng[0] - 1])]
    middle = len(rng) // 2
    left_range, right_range = rng[:middle], rng[middle:]
    left = form_segment_tree_dependency(model, left_range, fs, v)
    right = form_segment_tree_dependency(model, right_range, fs, v)
    for i in left_range:
        to_factor = model.m2f[(v, fs[i - 1])]
        if to_factor.listeners:
            add_dependency(to_factor, right, intermediate=True)
    for i in right_range:
        to_factor = model.m2f[(v, fs[i - 1])]
        if to_factor.listeners:
            add_dependency(to_factor, left, intermediate=True)
    product = Signal(variant=("product", v, rng[0], rng[-1]))  # ProductOfMessages(variable_id, range, factors)
    model.products.append(product)
    add_dependency(product, left, intermediate=True)
    add_dependency(product, right, intermediate=True)
    return product


def request_inference_for(marginals, linked_signals):  # src/inference_engine.jl:298-323
    for marginal, linked in zip(marginals, linked_signals):
        for dependency in marginal.dependencies:
            dependency.potentially_pending, dependency.pending = True, False
        for ls in linked:
            ls.potentially_pending, ls.pending = True, False


def scan_inference_request(marginals):  # :528-546 — the scanner's process! only collects, nothing is computed
    found = []

    def f(dependency):
        if is_pending(dependency):
            found.append(dependency)
            return True
        return False

    for marginal in marginals:
        process_dependencies(f, marginal, retry=True)
    return found
