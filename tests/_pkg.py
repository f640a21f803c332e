"""Loads the product package (`cortex.jl_b200/`, registered as module cortex_jl_b200) for the tests."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

import __graft_entry__ as entry  # noqa: E402

pkg = entry.load_package()
ORACLE_LIB = ROOT / "oracle" / "liboracle.so"
