"""CPU: the C-ABI library builds, loads and exports every symbol include/gmt_c_api.h declares.
No compute call is made here (there is no GPU); the only call that touches CUDA checks that the
library FAILS LOUDLY without a device instead of falling back to anything."""
import ctypes as C
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "gmt_c_api.h")).read()
    return sorted(set(re.findall(r"GMT_API[^;(]*?\b(gmt_\w+)\s*\(", text)))


def test_header_declares_the_survey_boundary():
    names = _declared()
    for must in ("gmt_create", "gmt_destroy", "gmt_plan", "gmt_get_route", "gmt_get_parent", "gmt_get_cost",
                 "gmt_get_trace", "gmt_plan_batch", "gmt_sample_states", "gmt_build_neighbors",
                 "gmt_edge_costs", "gmt_last_error"):
        assert must in names


def test_library_exports_every_declared_symbol(gmt_lib):
    import _lib
    names = _declared()
    assert sorted(_lib.PROTOTYPES) == names          # the ctypes table and the header agree
    for name in names:
        assert hasattr(gmt_lib, name), name
    assert gmt_lib.gmt_version() >= 100


def test_library_exports_nothing_the_header_does_not_declare():
    """exports of libgmt_b200.so with the gmt_ prefix == the header's declarations (no undeclared debug hooks)."""
    import subprocess
    lib = os.path.join(ROOT, "robot-parallel-motion-planning_b200", "libgmt_b200.so")
    out = subprocess.run(["nm", "-D", "--defined-only", lib], stdout=subprocess.PIPE, text=True, check=True).stdout
    exported = sorted({l.split()[-1] for l in out.splitlines() if l.split() and l.split()[-1].startswith("gmt_")})
    assert exported == _declared()


def test_no_cpu_fallback_without_a_device(gmt_lib):
    import torch
    if torch.cuda.is_available():
        return
    states = np.zeros((4, 3), np.float32)
    cnt = np.zeros(4, np.int32)
    h = C.c_void_p()
    rc = gmt_lib.gmt_create(states.ctypes.data, 4, None, cnt.ctypes.data, 0, C.byref(h))
    assert rc == -2 and not h.value                   # GMT_ERR_CUDA
    assert b"no CUDA device" in gmt_lib.gmt_last_error()
    out = np.zeros((4, 3), np.float32)
    assert gmt_lib.gmt_sample_states(1, 4, C.c_float(10.0), 0, 0, out.ctypes.data) == -2


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "robot-parallel-motion-planning_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f
                assert "libgmt_oracle" not in text and "libref_host" not in text, f
