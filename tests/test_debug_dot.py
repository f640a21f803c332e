"""Ports of test/ext/graph_viz_ext_tests.jl against `signal_to_dot` (the DOT-source counterpart of the GraphViz extension,
SURVEY §8f.4): same display options, same summaries, same colour / style conventions, read back through the C ABI."""
import pytest

from tests._pkg import pkg

C = pkg
cap = pkg.capi


@pytest.fixture
def store(backend):
    return C.SignalStore(backend, value_dim=1, family=cap.FAMILY_SUM)


def test_dot_is_a_digraph_with_value_and_variant(store):  # :11-85
    s = store.Signal(42)
    dot = C.signal_to_dot(s)
    assert dot.startswith("digraph G {") and dot.rstrip().endswith("}")
    assert "MainSignal" in dot and "Current value: 42" in dot and "Variant: " in dot and "Unspecified" in dot
    empty = store.Signal()
    assert "UndefValue()" in C.signal_to_dot(empty)


def test_dot_dependencies(store):  # :87-132
    s = store.Signal()
    assert "No dependencies" in C.signal_to_dot(s)
    dep1, dep2 = store.Signal(11), store.Signal(22)
    C.add_dependency(s, dep1)
    C.add_dependency(s, dep2)
    dot = C.signal_to_dot(s)
    assert "dependency 1" in dot and "dependency 2" in dot and "Current value: 11" in dot and "Current value: 22" in dot
    shallow = C.signal_to_dot(s, max_depth=0)
    assert "2 dependencies" in shallow and "Use `max_depth` to render more dependencies" in shallow
    assert "Current value: 11" not in shallow


def test_dot_reflects_pending_state(store):  # :134-153
    source, dependent = store.Signal(1), store.Signal()
    C.add_dependency(dependent, source)
    assert C.is_pending(dependent)
    pending = C.signal_to_dot(dependent)
    assert 'fillcolor="orange"' in pending.split("main [")[1].split("];")[0]
    C.set_value(dependent, 42)
    assert not C.is_pending(dependent)
    not_pending = C.signal_to_dot(dependent)
    assert pending != not_pending and 'fillcolor="palegreen"' in not_pending.split("main [")[1].split("];")[0]


def test_dot_dependency_styles_differ(store):  # :155-191
    dots = []
    for intermediate, weak in ((True, True), (True, False), (False, True), (False, False)):
        s = store.Signal(1)
        dep1, dep2 = store.Signal(), store.Signal()
        C.add_dependency(s, dep1, intermediate=intermediate, weak=weak)
        C.add_dependency(s, dep2, weak=weak)
        dep3, dep4 = store.Signal(3), store.Signal(4)
        C.add_dependency(dep1, dep3)
        C.add_dependency(dep1, dep4, intermediate=intermediate, weak=weak)
        C.add_dependency(dep2, store.Signal())
        dots.append(C.signal_to_dot(s, max_depth=100))
    assert len(set(dots)) == 4
    assert 'style="dashed" color="gray"' in dots[0] and 'style="solid" color="gray"' in dots[1]
    assert 'style="dashed" color="black"' in dots[2] and 'style="solid" color="blue"' in dots[3]  # dep3 is computed: fresh


def test_dot_display_options(store):  # :193-242
    s = store.Signal(42)
    assert "Current value: " in C.signal_to_dot(s, show_value=True) and "42" in C.signal_to_dot(s, show_value=True)
    assert "Current value: " not in C.signal_to_dot(s, show_value=False) and "42" not in C.signal_to_dot(s, show_value=False)
    assert "Variant: " in C.signal_to_dot(s, show_variant=True) and "Variant: " not in C.signal_to_dot(s, show_variant=False)
    custom = C.signal_to_dot(s, variant_to_string_fn=lambda v: f"CUSTOM_{v}")
    assert custom != C.signal_to_dot(s, variant_to_string_fn=str) and "CUSTOM_" in custom
    source, dependent = store.Signal(1), store.Signal(2)
    C.add_dependency(dependent, source)
    minimal = C.signal_to_dot(dependent, show_value=False, show_variant=False)
    assert "Current value: " not in minimal and "Variant: " not in minimal
    full = C.signal_to_dot(dependent, show_value=True, show_variant=True)
    assert full.count("Current value: ") == 2 and full.count("Variant: ") == 2
    assert C.signal_to_dot(dependent, variant_to_string_fn=lambda v: f"CUSTOM_{v}").count("CUSTOM_") == 2  # propagates


def test_dot_max_dependencies_statistics(store):  # :244-282
    main = store.Signal(1)
    deps = [store.Signal(i) for i in range(1, 16)]
    for i, dep in enumerate(deps, start=1):
        C.add_dependency(main, dep, weak=i % 2 == 0, intermediate=i % 3 == 0)
        if i % 4 == 0:
            C.set_value(dep, i)
    default = C.signal_to_dot(main)
    assert "5 more dependencies" in default and "Use `max_dependencies` to show more dependencies" in default
    assert "2 weak" in default and "2 intermediate" in default
    custom = C.signal_to_dot(main, max_dependencies=5)
    assert "10 more dependencies" in custom and "Use `max_dependencies` to show more dependencies" in custom
    everything = C.signal_to_dot(main, max_dependencies=20)
    assert "more dependencies" not in everything and "Use `max_dependencies`" not in everything


def test_dot_listeners_and_their_styles(store):  # :284-306
    main = store.Signal(1)
    active, inactive = store.Signal(2), store.Signal(3)
    C.add_dependency(active, main, listen=True)
    C.add_dependency(inactive, main, listen=False)
    shown, hidden = C.signal_to_dot(main, show_listeners=True), C.signal_to_dot(main, show_listeners=False)
    assert shown != hidden and "Listener" in shown and "Listener" not in hidden
    assert 'main -> listener1 [style="solid" color="black"]' in shown
    assert 'main -> listener2 [style="dotted" color="gray40"]' in shown


def test_dot_max_listeners_statistics(store):  # :308-340
    main = store.Signal(1)
    for i in range(1, 16):
        C.add_dependency(store.Signal(i), main, listen=i % 2 == 0)
    default = C.signal_to_dot(main, show_listeners=True)
    assert "5 more listeners" in default and "Use `max_listeners` to show more listeners" in default
    assert "2 active" in default and "3 inactive" in default
    assert "more listeners" not in C.signal_to_dot(main, max_listeners=20)
