"""The C-ABI library loads without a GPU, exports every symbol include/b200sqp.h declares, and the header is plain C whose struct
layouts match the ctypes mirror."""
import ctypes
import os
import re
import subprocess
import sys
import tempfile

import pytest

from control_box_rst_b200 import _abi as abi
from control_box_rst_b200 import problems, solver

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "b200sqp.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200sqp_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = solver.load_library()
    names = declared_symbols()
    assert len(names) >= 25
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/b200sqp.h but not exported by libb200sqp.so"
    assert sorted(solver.ABI_SYMBOLS) == names


def test_header_is_plain_c_and_struct_sizes_match_ctypes():
    src = '#include <stdio.h>\n#include "b200sqp.h"\nint main(void){printf("%zu %zu %zu\\n", sizeof(b200sqp_ocp), sizeof(b200sqp_lm_options), sizeof(b200sqp_dims));return 0;}\n'
    with tempfile.TemporaryDirectory() as tmp:
        c = os.path.join(tmp, "t.c")
        open(c, "w").write(src)
        exe = os.path.join(tmp, "t")
        subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), c, "-o", exe], check=True)
        sizes = [int(x) for x in subprocess.run([exe], check=True, capture_output=True, text=True).stdout.split()]
    assert sizes == [ctypes.sizeof(abi.Ocp), ctypes.sizeof(abi.LmOptions), ctypes.sizeof(abi.Dims)]


def test_no_cpu_fallback_without_device():
    """Without a usable GPU the compute entry points fail loudly (B200SQP_ERR_NO_DEVICE); structure queries still work."""
    if solver.device_available():
        pytest.skip("a GPU is present")
    assert solver.dims_of(problems.van_der_pol(20)).n_params == 57
    with pytest.raises(solver.B200SqpError) as e:
        solver.BatchedLevenbergMarquardt(problems.van_der_pol(20), 4)
    assert e.value.code == abi.ERR_NO_DEVICE


def test_unsupported_and_invalid_descriptors_are_rejected():
    ocp = problems.van_der_pol(20)
    ocp.dynamics = 99
    with pytest.raises(solver.B200SqpError) as e:
        solver.dims_of(ocp)
    assert e.value.code == abi.ERR_UNSUPPORTED  # closed functor registry -> SolverStatus::Error on the reference side
    ocp = problems.van_der_pol(20)
    ocp.n_grid = 1
    with pytest.raises(solver.B200SqpError) as e:
        solver.dims_of(ocp)
    assert e.value.code == abi.ERR_INVALID
    ocp = problems.van_der_pol(20)
    ocp.nx = 3
    with pytest.raises(solver.B200SqpError) as e:
        solver.dims_of(ocp)
    assert e.value.code == abi.ERR_INVALID
    ocp = problems.van_der_pol(20)
    ocp.zero_u_ref = 0  # quadratic_cost.cpp:161: lsq form with a non-zero control reference is not a vector lsq term
    with pytest.raises(solver.B200SqpError) as e:
        solver.dims_of(ocp)
    assert e.value.code == abi.ERR_NOT_LSQ
    ocp = problems.van_der_pol(20)
    ocp.grid = 7
    with pytest.raises(solver.B200SqpError) as e:
        solver.dims_of(ocp)
    assert e.value.code == abi.ERR_UNSUPPORTED


def test_product_never_imports_the_oracle():
    """The shipped package must not reference oracle/ (a product path through the oracle would void every parity claim)."""
    pkg = os.path.join(ROOT, "control_box_rst_b200")
    for dirpath, _, files in os.walk(pkg):
        if "_build" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cpp", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "sqp_oracle" not in text and "corbo_ref" not in text and "import oracle" not in text and "from oracle" not in text, f


def test_every_golden_combination_is_compiled_into_the_library():
    """b200sqp_create looks the (dynamics, defect, grid) combination up in the kernel tables before it asks for a device: without a
    GPU a compiled-in combination answers B200SQP_ERR_NO_DEVICE, one missing from kernels_*.cu B200SQP_ERR_UNSUPPORTED.  Every case
    of the golden table must be compiled in; a combination outside the tables must not be."""
    if solver.device_available():
        pytest.skip("a GPU is present (the GPU parity suite creates a handle for every case)")
    import sys

    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import cases

    for name, (make, _, _) in cases.CASES.items():
        with pytest.raises(solver.B200SqpError) as e:
            solver.BatchedLevenbergMarquardt(make(), 4)
        assert e.value.code == abi.ERR_NO_DEVICE, name
    missing = problems.quadrotor(8)
    missing.grid = abi.GRID_MULTIPLE_SHOOTING  # a shooting quadrotor is a valid structure but not in the kernel tables
    with pytest.raises(solver.B200SqpError) as e:
        solver.BatchedLevenbergMarquardt(missing, 4)
    assert e.value.code == abi.ERR_UNSUPPORTED
