"""Port of the reference's test/signal_tests.jl (pending-state truth table, compute! guards,
process_dependencies! traversal order). Runs against the oracle and, on a GPU box, the device
engine through the same C ABI. Citations are test/signal_tests.jl line ranges."""
import pytest

from tests._pkg import pkg

C = pkg
add_dependency, set_value, get_value = C.add_dependency, C.set_value, C.get_value
is_pending, is_computed = C.is_pending, C.is_computed
get_dependencies, get_listeners = C.get_dependencies, C.get_listeners


@pytest.fixture
def pool(backend):
    return C.SignalStore(backend, value_dim=1, family=C.capi.FAMILY_SUM, dtype=C.capi.F64)


def test_basic_signal_operations(pool):  # :1-21
    s = pool.Signal(42)
    assert get_value(s) == 42
    set_value(s, 100)
    assert get_value(s) == 100
    with pytest.raises(Exception):
        set_value(s, [1.0, 2.0])  # typed signal rejects a value of the wrong shape (:19-20)
    with pytest.raises(Exception):
        set_value(s, "abc")


def test_empty_signal_creation(pool):  # :69-88
    s = pool.Signal()
    assert isinstance(C.get_variant(s), C.Unspecified)
    assert get_dependencies(s) == [] and get_listeners(s) == []
    assert not is_pending(s) and not is_computed(s)


def test_signal_creation_with_value_sets_computed(pool):  # :90-97
    s = pool.Signal(10)
    assert get_value(s) == 10 and is_computed(s) and not is_pending(s)


def test_add_dependency_basic(pool):  # :99-135
    a, b = pool.Signal(1), pool.Signal(2)
    assert not is_pending(a) and not is_pending(b)
    add_dependency(a, b)
    assert get_dependencies(a) == [b] and get_listeners(a) == []
    assert get_dependencies(b) == [] and get_listeners(b) == [a]
    assert not is_pending(a) and not is_pending(b)
    set_value(b, 3)
    assert is_pending(a) and not is_pending(b)


def test_add_dependency_of_different_engine_throws(backend):  # :137-162
    p1 = C.SignalStore(backend, 1)
    p2 = C.SignalStore(backend, 1)
    a, b = p1.Signal(1), p2.Signal(2)
    for kw in ({}, {"weak": True}, {"listen": False}, {"check_computed": False}, {"intermediate": True}):
        with pytest.raises(Exception):
            add_dependency(a, b, **kw)
        with pytest.raises(Exception):
            add_dependency(b, a, **kw)


def test_add_dependency_initialized(pool):  # :164-182
    dep, s = pool.Signal(1), pool.Signal()
    assert not is_pending(s) and not is_computed(s) and is_computed(dep)
    add_dependency(s, dep)
    assert get_dependencies(s) == [dep] and get_listeners(dep) == [s]
    assert is_pending(s) and not is_computed(s)


def test_single_non_initialized_weak_dependency(pool):  # :184-202
    s1, s2 = pool.Signal(), pool.Signal()
    add_dependency(s2, s1, weak=True)
    assert not is_pending(s2) and not is_computed(s2)
    set_value(s1, 10)
    assert is_pending(s2) and not is_computed(s2)


def test_single_initialized_weak_dependency(pool):  # :204-222
    s1, s2 = pool.Signal(1), pool.Signal()
    add_dependency(s2, s1, weak=True)
    assert is_pending(s2) and not is_computed(s2)
    set_value(s1, 10)
    assert is_pending(s2)


def test_initialized_dependency_without_check_computed(pool):  # :224-242
    s1, s2 = pool.Signal(1), pool.Signal()
    add_dependency(s2, s1, check_computed=False)
    assert not is_pending(s2) and not is_computed(s2)
    set_value(s1, 10)
    assert is_pending(s2)


def test_many_dependencies_all_strong(pool):  # :244-285
    s1, s2, s3, d = pool.Signal(), pool.Signal(), pool.Signal(), pool.Signal()
    for s in (s1, s2, s3):
        add_dependency(d, s)
    assert get_dependencies(d) == [s1, s2, s3]
    assert get_listeners(s1) == [d] and get_listeners(s2) == [d] and get_listeners(s3) == [d]
    assert not is_pending(d) and not is_computed(d)
    set_value(s1, 1)
    assert not is_pending(d) and is_computed(s1)
    set_value(s2, 2)
    assert not is_pending(d)
    set_value(s3, 3)
    assert is_pending(d) and not is_computed(d)
    set_value(d, 10)
    assert not is_pending(d) and is_computed(d)


@pytest.mark.parametrize("init", [False, True])
def test_update_dependency_marks_pending(pool, init):  # :287-331
    s1, s2 = (pool.Signal(1), pool.Signal(2)) if init else (pool.Signal(), pool.Signal())
    assert not is_pending(s1) and not is_pending(s2)
    add_dependency(s1, s2)
    assert not is_pending(s1) and not is_pending(s2)
    set_value(s2, 3)
    assert is_pending(s1) and not is_pending(s2)
    assert is_computed(s1) == init and is_computed(s2)


def test_weak_dependencies_basic(pool):  # :333-366
    weak, strong, d = pool.Signal(1), pool.Signal(2), pool.Signal()
    add_dependency(d, weak, weak=True)
    add_dependency(d, strong)
    assert get_dependencies(d) == [weak, strong]
    assert is_pending(d) and not is_computed(d)
    set_value(d, 10)
    assert not is_pending(d) and is_computed(d)
    set_value(strong, 3)
    assert is_pending(d)
    set_value(d, 11)
    assert not is_pending(d)
    set_value(weak, 4)
    assert not is_pending(d)
    set_value(strong, 5)
    assert is_pending(d)


def test_many_weak_dependencies(pool):  # :368-440
    w1, w2, s1, d = pool.Signal(), pool.Signal(), pool.Signal(), pool.Signal()
    add_dependency(d, w1, weak=True)
    add_dependency(d, w2, weak=True)
    add_dependency(d, s1)
    assert get_dependencies(d) == [w1, w2, s1] and not is_pending(d)
    set_value(s1, 10)
    assert not is_pending(d)
    set_value(w1, 1)
    assert not is_pending(d)
    set_value(w2, 2)
    assert is_pending(d) and not is_computed(d)
    set_value(d, 100)
    assert not is_pending(d) and is_computed(d)
    set_value(s1, 11)
    assert is_pending(d)
    set_value(d, 101)
    assert not is_pending(d)
    set_value(w1, 3)
    assert not is_pending(d)
    set_value(s1, 333)
    assert is_pending(d)


def test_duplicate_dependencies_never_notified(pool):  # :442-465
    s1, s2 = pool.Signal(), pool.Signal()
    add_dependency(s1, s2)
    add_dependency(s1, s2)
    assert get_dependencies(s1) == [s2, s2] and get_listeners(s2) == [s1, s1]
    assert not is_pending(s1)
    set_value(s2, 1)
    assert not is_pending(s1)


def test_circular_dependencies(pool):  # :467-507
    s1, s2 = pool.Signal(), pool.Signal()
    add_dependency(s1, s2)
    add_dependency(s2, s1)
    assert get_dependencies(s1) == [s2] and get_listeners(s1) == [s2]
    assert not is_pending(s1) and not is_pending(s2)
    set_value(s1, 1)
    assert not is_pending(s1) and is_pending(s2)
    set_value(s2, 2)
    assert is_pending(s1) and not is_pending(s2)
    set_value(s2, 3)
    assert is_pending(s1) and not is_pending(s2)
    set_value(s1, 4)
    assert not is_pending(s1) and is_pending(s2)


def test_self_dependency_does_nothing(pool):  # :509-521
    s1 = pool.Signal()
    add_dependency(s1, s1)
    assert get_dependencies(s1) == [] and get_listeners(s1) == [] and not is_pending(s1)


def test_pending_state_logic_coverage(pool):  # :523-591
    d, s = pool.Signal(), pool.Signal()
    add_dependency(d, s)
    assert not is_pending(d)
    set_value(d, 1)
    assert not is_pending(d)
    d, s = pool.Signal(), pool.Signal()
    add_dependency(d, s)
    set_value(s, 10)
    assert is_pending(d)
    # weak dep not computed
    d, w = pool.Signal(), pool.Signal()
    add_dependency(d, w, weak=True)
    assert not is_pending(d)
    set_value(d, 1)
    assert not is_pending(d)
    set_value(w, 10)
    assert is_pending(d)
    # strong dep computed but not older
    d, s = pool.Signal(1), pool.Signal(10)
    add_dependency(d, s)
    assert not is_pending(d)
    set_value(d, 100)
    assert not is_pending(d)
    set_value(s, 101)
    assert is_pending(d)
    set_value(d, 102)
    assert not is_pending(d)
    set_value(s, 103)
    assert is_pending(d)
    # mixed
    d, w, s = pool.Signal(), pool.Signal(), pool.Signal()
    add_dependency(d, w, weak=True)
    add_dependency(d, s)
    assert not is_pending(d)
    set_value(w, 1)
    assert not is_pending(d)
    set_value(s, 2)
    assert is_pending(d)


def test_chain_of_signals(pool):  # :593-637
    s1, s2, s3 = pool.Signal(1), pool.Signal(), pool.Signal()
    add_dependency(s2, s1)
    add_dependency(s3, s2)
    assert [is_pending(s) for s in (s1, s2, s3)] == [False, True, False]
    for base in (2, 5):
        set_value(s1, base)
        assert [is_pending(s) for s in (s1, s2, s3)] == [False, True, False]
        set_value(s2, base + 1)
        assert [is_pending(s) for s in (s1, s2, s3)] == [False, False, True]
        set_value(s3, base + 2)
        assert [is_pending(s) for s in (s1, s2, s3)] == [False, False, False]


def test_not_listening_dependency(pool):  # :639-710
    s1, s2 = pool.Signal(1), pool.Signal(2)
    add_dependency(s2, s1, listen=False)
    assert not is_pending(s2)
    set_value(s1, 10)
    assert not is_pending(s2)
    s1, s2 = pool.Signal(1), pool.Signal(2)
    add_dependency(s2, s1, listen=False, weak=True)
    assert is_pending(s2)
    set_value(s1, 10)
    assert is_pending(s2)
    s1, s2 = pool.Signal(1), pool.Signal(2)
    add_dependency(s2, s1, listen=False, check_computed=False)
    assert not is_pending(s2)
    set_value(s1, 10)
    assert not is_pending(s2)
    s1, s2, s3 = pool.Signal(), pool.Signal(), pool.Signal()
    add_dependency(s3, s1, listen=False)
    add_dependency(s3, s2)
    assert not is_pending(s3)
    set_value(s2, 10)
    assert not is_pending(s3)
    set_value(s1, 10)
    assert not is_pending(s3)  # does not listen to s1
    set_value(s2, 30)
    assert is_pending(s3)


def test_adding_computed_then_uncomputed_unsets_pending(pool):  # :712-749
    s1, s2, d = pool.Signal(1), pool.Signal(), pool.Signal()
    add_dependency(d, s1)
    assert is_pending(d)
    add_dependency(d, s2)
    assert not is_pending(d)
    s1, s2, d = pool.Signal(1), pool.Signal(), pool.Signal()
    add_dependency(d, s1, check_computed=True)
    assert is_pending(d)
    add_dependency(d, s2, check_computed=False)
    assert is_pending(d)


def test_signal_representation(pool):  # :751-788
    s = pool.Signal()
    assert repr(s).startswith("Signal(value=#undef, pending=false")
    s1 = pool.Signal(1)
    sp = pool.Signal()
    add_dependency(sp, s1)
    assert "pending=true" in repr(sp)
    set_value(sp, 50)
    assert "value=50.0, pending=false" in repr(sp)
    set_value(s1, 2)
    assert "value=50.0, pending=true" in repr(sp)


def test_compute_with_sum_strategy(pool):  # :834-871 (strategy = sum of dependency values)
    s1, s2, s3 = pool.Signal(1), pool.Signal(2), pool.Signal()
    add_dependency(s3, s1)
    add_dependency(s3, s2)
    assert is_pending(s3) and not is_computed(s3)
    C.compute(s3)
    assert is_computed(s3) and not is_pending(s3) and get_value(s3) == 3
    with pytest.raises(C.NotPendingError):  # ArgumentError, :857
        C.compute(s3)
    C.compute(s3, force=True)
    assert get_value(s3) == 3 and not is_pending(s3)
    set_value(s1, 10)
    set_value(s2, 20)
    assert is_pending(s3)
    C.compute(s3)
    assert is_computed(s3) and not is_pending(s3) and get_value(s3) == 30


def test_compute_pyramid(pool):  # :873-916
    s01, s02, s11, s12 = pool.Signal(1), pool.Signal(2), pool.Signal(3), pool.Signal(4)
    s21, s22, s3 = pool.Signal(), pool.Signal(), pool.Signal()
    add_dependency(s21, s01)
    add_dependency(s21, s02)
    add_dependency(s22, s11)
    add_dependency(s22, s12)
    add_dependency(s3, s21)
    add_dependency(s3, s22)
    assert is_pending(s21) and is_pending(s22) and not is_pending(s3)
    C.compute(s21)
    C.compute(s22)
    assert not is_pending(s21) and not is_pending(s22) and is_pending(s3) and not is_computed(s3)
    C.compute(s3)
    assert is_computed(s3) and not is_pending(s3) and get_value(s3) == 10


def test_compute_skip_if_no_listeners(pool):  # :1116-1133
    s = pool.Signal(1)
    C.compute(s, skip_if_no_listeners=True)
    assert get_value(s) == 1
    with pytest.raises(Exception):  # forced compute of a signal without dependencies has nothing to reduce
        C.compute(s, force=True, skip_if_no_listeners=False)


def test_intermediate_dependencies_can_be_added(pool):  # :919-931
    src, mid, d = pool.Signal(), pool.Signal(), pool.Signal()
    add_dependency(mid, src)
    add_dependency(d, mid, intermediate=True)
    assert get_dependencies(d) == [mid] and get_dependencies(mid) == [src]
    assert C.get_dependency_props(d) == [C.capi.NIB_INTERMEDIATE]


# ---- process_dependencies! traversal: the reference's visit order, on the oracle and on the device ------------------
# (the callback is a pure predicate, tabulated before the traversal: cxb_process_dependencies_table)
def _chain(p, intermediate):
    src, mid, d = p.Signal(), p.Signal(), p.Signal()
    add_dependency(mid, src)
    add_dependency(d, mid, intermediate=intermediate)
    return src, mid, d


@pytest.mark.parametrize("retry", [False, True])
def test_process_dependencies_steps_down_callback_false(pool, retry):  # :943-973
    src, mid, d = _chain(pool, True)
    seen = []
    r = C.process_dependencies(lambda dep: False, d, retry=retry, visited=seen)
    assert seen == [mid, src] and not r


@pytest.mark.parametrize("retry", [False, True])
def test_process_dependencies_callback_true(pool, retry):  # :975-999
    src, mid, d = _chain(pool, True)
    seen = []
    r = C.process_dependencies(lambda dep: True, d, retry=retry, visited=seen)
    assert seen == [mid] and r


def test_process_dependencies_retry_order(pool):  # :1001-1028
    src, mid, d = _chain(pool, True)
    seen = []
    r = C.process_dependencies(lambda dep: dep != mid, d, retry=False, visited=seen)
    assert seen == [mid, src] and r
    seen = []
    r = C.process_dependencies(lambda dep: dep != mid, d, retry=True, visited=seen)
    assert seen == [mid, src, mid] and r  # retried on the intermediate


@pytest.mark.parametrize("retry", [False, True])
@pytest.mark.parametrize("ret", [False, True])
def test_process_dependencies_not_intermediate(pool, retry, ret):  # :1031-1059
    src, mid, d = _chain(pool, False)
    seen = []
    r = C.process_dependencies(lambda dep: ret, d, retry=retry, visited=seen)
    assert seen == [mid] and r == ret


@pytest.mark.parametrize("retry", [False, True])
def test_process_dependencies_returns_true_if_any(pool, retry):  # :1061-1082
    src, mid, d = _chain(pool, True)
    seen = []
    assert C.process_dependencies(lambda dep: dep == src, d, retry=retry, visited=seen)
    assert len(seen) >= 1


def test_process_dependencies_scanner_callback_and_arbitrary_python_callback(pool, oracle_api):
    """f = None is the scanner's is_pending callback; the oracle additionally accepts any Python callable
    (tests/oracle_frontend.py) and must visit in the same order."""
    from tests.oracle_frontend import process_dependencies_callback

    def build(p):
        a, b, c, top = p.Signal(), p.Signal(), p.Signal(), p.Signal()
        add_dependency(b, a)
        add_dependency(c, b, intermediate=True)
        add_dependency(top, c, intermediate=True)
        add_dependency(top, a, intermediate=True)
        C.set_value(a, 1.0)
        return a, b, c, top

    a, b, c, top = build(pool)
    seen = []
    r = C.process_dependencies(None, top, retry=True, visited=seen)
    # c is not pending -> descend; b is pending (a is fresh) -> processed; retry asks c again; a has no dependencies
    assert seen == [c, b, c, a] and r
    op = C.SignalStore(oracle_api, 1)
    a2, b2, c2, top2 = build(op)
    seen2 = []
    r2 = process_dependencies_callback(lambda dep: (seen2.append(dep), C.is_pending(dep))[1], top2, retry=True)
    assert [s.sid for s in seen2] == [s.sid for s in seen] and r2 == r
