"""C-ABI library: loads without a GPU, exports every symbol include/qmb200.h declares, host-side loaders agree with the
oracle's independent parse of the same files; compute entry points fail loudly without a device (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from helpers import ROOT
from oracle import abi_fill, config, gait as G, scenarios


def test_exports_every_declared_symbol(built):
    import qm_door_b200 as q
    L = q.lib()
    hdr = open(os.path.join(ROOT, "include", "qmb200.h")).read()
    names = set(re.findall(r"\b(qmb200_[a-z_0-9]+)\s*\(", hdr))
    assert len(names) >= 18
    for n in sorted(names):
        assert hasattr(L, n), "missing export: " + n
    assert L.qmb200_version() == 100


def test_struct_layout_matches_header(built):
    """ctypes mirrors and the C structs agree on size (guards against silent ABI drift)."""
    import subprocess, tempfile
    from qm_door_b200 import _abi
    src = '#include <stdio.h>\n#include "%s/include/qmb200.h"\nint main(){printf("%%zu %%zu %%zu\\n",sizeof(qmb200_model_desc),sizeof(qmb200_problem_desc),sizeof(qmb200_solver_desc));return 0;}\n' % ROOT
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "a.c"), "w").write(src)
        subprocess.check_call(["gcc", "-o", os.path.join(d, "a"), os.path.join(d, "a.c")])
        sizes = [int(v) for v in subprocess.check_output([os.path.join(d, "a")]).split()]
    assert sizes == [C.sizeof(_abi.ModelDesc), C.sizeof(_abi.ProblemDesc), C.sizeof(_abi.SolverDesc)]


def test_loaders_match_oracle_parse(descs, oracle_inputs):
    from qm_door_b200 import _abi
    model, problem, solver, x_init = descs
    m, P = oracle_inputs
    a, b = _abi.struct_to_dict(model), _abi.struct_to_dict(abi_fill.model_desc(m))
    for k in a:
        assert np.allclose(np.asarray(a[k], dtype=float), np.asarray(b[k], dtype=float), rtol=1e-14, atol=1e-16), k
    a, b = _abi.struct_to_dict(problem), _abi.struct_to_dict(abi_fill.problem_desc(m, P))
    for k in a:
        assert np.allclose(np.asarray(a[k]), np.asarray(b[k]), rtol=1e-12, atol=1e-15), k
    assert np.array_equal(x_init, P.x_init)
    assert solver.dt == 0.015 and solver.horizon == 1.0 and solver.g_max == 1e-2 and solver.g_min == 1e-6


@pytest.mark.skipif(not os.path.exists("/root/reference/qm_controllers/config/task.info"), reason="reference checkout absent")
def test_loaders_on_the_reference_files(built, oracle_inputs):
    """Drop-in check: the C++ loaders read the reference's own robot.urdf / task.info / gait.info unchanged."""
    import qm_door_b200 as q
    from qm_door_b200 import _abi
    R = "/root/reference/"
    model = q.load_model(R + "qm_description/urdf/quadruped_manipulator/robot.urdf")
    problem, solver, x_init = q.load_problem(model, R + "qm_controllers/config/task.info", R + "qm_controllers/config/reference.info")
    m, P = oracle_inputs
    assert abs(model.total_mass - 27.86796983) < 1e-8
    a, b = _abi.struct_to_dict(model), _abi.struct_to_dict(abi_fill.model_desc(m))
    for k in a:
        assert np.allclose(np.asarray(a[k], dtype=float), np.asarray(b[k], dtype=float), rtol=1e-14, atol=1e-16), k
    for name in P.gait_list:
        sw, md = q.load_gait(name, R + "qm_controllers/config/gait.info")
        assert np.array_equal(sw, P.gaits[name]["times"]) and list(md) == P.gaits[name]["modes"]


def test_missing_file_raises_like_the_reference(built):
    """QMInterface.cpp:41-62 throws std::invalid_argument on missing files -> non-zero status + message here."""
    import qm_door_b200 as q
    with pytest.raises(q.Qmb200Error, match="file not found"):
        q.load_model("/nonexistent/robot.urdf")


def test_tile_schedule_matches_oracle(built, oracle_inputs):
    import qm_door_b200 as q
    m, P = oracle_inputs
    for name in ("trot", "flying_trot", "dynamic_walk", "stance"):
        sw, md = q.load_gait(name)
        ev, ms, n = q.tile_schedule(sw, md, -0.73, 1.9, 64)
        e2, m2 = G.tile_schedule(P.gaits[name], -0.73, 1.9)
        assert n == len(e2) and np.array_equal(ev[:n], e2) and np.array_equal(ms[:n + 1], m2)


def test_workload_reference_pose(built, oracle_inputs):
    from qm_door_b200 import workload
    m, P = oracle_inputs
    _, ts = scenarios.standing_target(m, P)
    assert np.allclose(workload.nominal_ee_pose(), ts[0][30:], atol=1e-12)
    W = workload.Workload(3, horizon=0.2)
    x0, phase = scenarios.perturbed_states(m, P, 3)
    assert np.allclose(W.x0, x0) and np.allclose(W.phase, phase)


def test_compute_fails_loudly_without_gpu(descs):
    import qm_door_b200 as q
    model, problem, solver, _ = descs
    if q.lib().qmb200_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(q.Qmb200Error, match="no CUDA device"):
        q.MpcContext(model, problem, solver, 4)
