"""TEST INFRASTRUCTURE ONLY -- loader of the unmodified reference (kitrofimov/srp) built into
oracle/_ref by oracle/Makefile, behind the same ctypes mirror of the srp C API that drives the
product (srp_b200.host.SrpLibrary).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may
import this module; nothing under srp_b200/ does, and the product path (srp_b200.host.load_product)
never loads anything from oracle/.
"""
from __future__ import annotations

from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
REFERENCE_SO = ROOT / "oracle" / "_ref" / "libref_host.so"

_loaded = {}


def available() -> bool:
    return REFERENCE_SO.exists()


def load_oracle_reference():
    """the reference library as an SrpLibrary (is_product=False)"""
    if "reference" not in _loaded:
        if not REFERENCE_SO.exists():
            raise FileNotFoundError(f"{REFERENCE_SO} is missing (it is built from /root/reference by oracle/Makefile)")
        from srp_b200.host import SrpLibrary
        _loaded["reference"] = SrpLibrary(REFERENCE_SO, is_product=False)
    return _loaded["reference"]
