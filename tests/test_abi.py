"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/cortex_b200.h declares; without a CUDA device every compute entry point fails loudly
(no CPU fallback, no routing through the oracle)."""
import ctypes
import re

import pytest

from tests._pkg import ROOT, pkg

cap = pkg.capi


def test_library_exports_every_declared_symbol():
    api = pkg.default_api()
    header = (ROOT / "include" / "cortex_b200.h").read_text()
    declared = set(re.findall(r"\b(cxb_[a-z0-9_]+)\s*\(", header))
    assert declared, "header declares no functions?"
    for name in sorted(declared):
        assert hasattr(api.lib, name), f"{name} is declared in include/cortex_b200.h but not exported"
    assert declared == set(pkg.exported_symbols()), "ctypes binding table and header disagree"
    assert b"sm_100a" in api.version()


def test_product_library_does_not_link_the_oracle():
    import subprocess

    out = subprocess.run(["nm", "-D", "--defined-only", pkg.default_api().path], capture_output=True, text=True).stdout
    assert "cxo_" not in out
    ldd = subprocess.run(["ldd", pkg.default_api().path], capture_output=True, text=True).stdout
    assert "liboracle" not in ldd


def test_sass_is_sm_100a():
    import subprocess

    out = subprocess.run(["cuobjdump", "-lelf", pkg.default_api().path], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_no_cpu_fallback_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    api = pkg.default_api()
    h = ctypes.c_void_p()
    assert api.create(0, cap.F64, 1, cap.FAMILY_SUM, ctypes.byref(h)) == cap.ERR_CUDA
    with pytest.raises(pkg.CortexError):
        pkg.SignalStore(api, 1)
    with pytest.raises(pkg.CortexError):
        pkg.GaussianChainBatch(4, 8)
    with pytest.raises(pkg.CortexError):
        pkg.PottsGrid(4, 4, 4, 0.5)
    with pytest.raises(pkg.CortexError):
        pkg.HmmBatch(2, 8, 4, 3)


@pytest.mark.gpu
def test_cxx_exceptions_do_not_cross_the_abi(device_api):
    """A host-side C++ exception (here: std::length_error from a negative id count in cxb_graph_build) is stopped at the C
    boundary and reported as a status; without the function-try-blocks it would terminate the calling Julia / Python
    process."""
    import numpy as np

    store = pkg.SignalStore(device_api, value_dim=1, family=pkg.capi.FAMILY_SUM, dtype=pkg.capi.F64)
    isf = np.zeros(1, dtype=np.uint8)
    st = device_api.graph_build(store.h, -1, isf.ctypes.data_as(pkg.capi.u8p), None, 0, None, None)
    assert st in (pkg.capi.ERR_INTERNAL, pkg.capi.ERR_BAD_ARG)


def test_graph_build_refuses_negative_sizes(oracle_api):
    """Same answer as the device library (host_graph.hpp): CXB_ERR_BAD_ARG, and the handle stays usable."""
    import numpy as np

    store = pkg.SignalStore(oracle_api, value_dim=1, family=pkg.capi.FAMILY_SUM, dtype=pkg.capi.F64)
    isf = np.zeros(1, dtype=np.uint8)
    assert oracle_api.graph_build(store.h, -1, isf.ctypes.data_as(pkg.capi.u8p), None, 0, None, None) == pkg.capi.ERR_BAD_ARG
    assert oracle_api.graph_build(store.h, 1, isf.ctypes.data_as(pkg.capi.u8p), None, -3, None, None) == pkg.capi.ERR_BAD_ARG
    assert oracle_api.graph_build(store.h, 1, isf.ctypes.data_as(pkg.capi.u8p), None, 0, None, None) == pkg.capi.OK
