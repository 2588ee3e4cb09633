"""Import the UNMODIFIED reference (ptwt) from /root/reference -- only possible in the build
container, never on the GPU box.  TEST INFRASTRUCTURE ONLY (used by make_golden.py and by the CPU
tests that cross-check the oracle against the real reference when it is present)."""
from __future__ import annotations

import importlib
import sys
from pathlib import Path

REFERENCE_SRC = Path("/root/reference/src")
SHIMS = Path(__file__).resolve().parent / "shims"


def reference_available() -> bool:
    return (REFERENCE_SRC / "ptwt" / "__init__.py").exists()


def import_reference():
    """Returns the reference ``ptwt`` module (with the pywt / more_itertools shims if needed)."""
    if not reference_available():
        raise RuntimeError("/root/reference is not present on this machine")
    try:
        importlib.import_module("pywt")
    except Exception:  # noqa: BLE001
        if str(SHIMS) not in sys.path:
            sys.path.insert(0, str(SHIMS))
    try:
        importlib.import_module("more_itertools")
    except Exception:  # noqa: BLE001
        if str(SHIMS) not in sys.path:
            sys.path.insert(0, str(SHIMS))
    if str(REFERENCE_SRC) not in sys.path:
        sys.path.insert(0, str(REFERENCE_SRC))
    return importlib.import_module("ptwt")
