"""The C++ host-side mirror of the reference's plugin seams (include/qmb200_adapters.hpp: what QMController::setupMpc / setupWbc
build, QMController.cpp:273-307) is compiled against the C-ABI and driven the way QMController::update drives the reference
objects (tests/cpp/adapter_driver.cpp). CPU: it builds, links, loads the robot / task files and fails loudly without a CUDA
device. GPU: its results equal the ctypes path on the same scenario bit for bit (same library underneath)."""
import os
import subprocess

import numpy as np
import pytest

from helpers import ROOT


@pytest.fixture(scope="module")
def driver(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("adapters") / "adapter_driver")
    libdir = os.path.join(ROOT, "qm_door_b200")
    subprocess.run(["make", "-C", ROOT, "qm_door_b200/libqmb200.so"], check=True, stdout=subprocess.DEVNULL)
    subprocess.run(["g++", "-std=c++17", "-Wall", "-Werror", "-O1", "-o", exe, os.path.join(ROOT, "tests", "cpp", "adapter_driver.cpp"),
                    "-L" + libdir, "-lqmb200", "-Wl,-rpath," + libdir], check=True)
    return exe


def write_scenario(path, cycles, t0, cycle_dt, tq, x0, events, modes, tt, tx, rbd, period, wbc_time):
    f = lambda a: " ".join("%.17g" % v for v in np.asarray(a, dtype=float).ravel())
    with open(path, "w") as fh:
        fh.write("%d %.17g %.17g %.17g\n%s\n%d\n%s\n%s\n%d\n%s\n%s\n%s\n%.17g %.17g\n" % (
            cycles, t0, cycle_dt, tq, f(x0), len(events), f(events), " ".join(str(int(m)) for m in modes), len(tt), f(tt), f(tx),
            f(rbd), period, wbc_time))


def scenario(q):
    from qm_door_b200 import workload
    W = workload.Workload(1, horizon=1.0, dt=0.015, seed=11)        # provides a perturbed state, a trot schedule and the standing target
    Ww = workload.WbcWorkload(1, seed=11)
    ne = int(W.nevents[0])
    return dict(x0=W.x0[0], events=W.events[0, :ne], modes=W.modes[0, :ne + 1], tt=W.target_t[0], tx=W.target_x[0], rbd=Ww.rbd[0])


def read_result(path):
    out = {}
    for ln in open(path):
        p = ln.split()
        out[p[0]] = float(p[1]) if len(p) == 2 else np.array([float(v) for v in p[2:]])
    return out


def test_driver_builds_and_fails_loudly_without_gpu(driver, tmp_path):
    import qm_door_b200 as q
    if q.lib().qmb200_device_count() > 0:
        pytest.skip("a CUDA device is present: covered by the GPU test")
    s = scenario(q)
    write_scenario(tmp_path / "s.txt", 1, 0.0, 0.01, 0.0, s["x0"], s["events"], s["modes"], s["tt"], s["tx"], s["rbd"], 0.001, 11.0)
    r = subprocess.run([driver, q.DEFAULT_TASK, q.DEFAULT_URDF, q.DEFAULT_REFERENCE, str(tmp_path / "s.txt"), str(tmp_path / "r.txt")],
                       capture_output=True, text=True)
    assert r.returncode == 1 and "no CUDA device" in r.stderr, r.stderr      # files were ingested, construction of the context threw
    r = subprocess.run([driver, "/nonexistent/task.info", q.DEFAULT_URDF, q.DEFAULT_REFERENCE, "x", "y"], capture_output=True, text=True)
    assert r.returncode == 1 and "adapter_driver:" in r.stderr               # std::invalid_argument like QMInterface.cpp:41-62


@pytest.mark.gpu
def test_driver_equals_ctypes_path(driver, tmp_path):
    import qm_door_b200 as q
    s = scenario(q)
    cycles, t0, cdt, tq, period, wtime = 2, 0.0, 0.01, 0.0137, 0.001, 11.0
    write_scenario(tmp_path / "s.txt", cycles, t0, cdt, tq, s["x0"], s["events"], s["modes"], s["tt"], s["tx"], s["rbd"], period, wtime)
    r = subprocess.run([driver, q.DEFAULT_TASK, q.DEFAULT_URDF, q.DEFAULT_REFERENCE, str(tmp_path / "s.txt"), str(tmp_path / "r.txt")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    res = read_result(tmp_path / "r.txt")
    # the same calls through ctypes, with the solver settings the C++ loader produced (task.info: horizon 1.0 s, dt 0.015 s)
    model = q.load_model()
    problem, solver, _ = q.load_problem(model)
    ctx = q.MpcContext(model, problem, solver, 1)
    E, K = solver.max_events, solver.max_targets
    ev = np.full((1, E), 1e30); ev[0, :len(s["events"])] = s["events"]
    md = np.full((1, E + 1), 15, dtype=np.int32); md[0, :len(s["modes"])] = s["modes"]
    ne = np.array([len(s["events"])], dtype=np.int32)
    assert K == len(s["tt"])
    for c in range(cycles):
        out = ctx.cycle(np.array([t0 + c * cdt]), s["x0"][None], ev, md, ne, s["tt"][None], s["tx"][None])
    n = int(out["n"][0])
    assert n == int(res["nodes"]) and np.array_equal(out["t"][0, :n], res["t"])
    assert np.array_equal(out["x"][0, :n].ravel(), res["x"]) and np.array_equal(out["u"][0, :n].ravel(), res["u"])
    xd, ud, mode = ctx.evaluate_policy(np.array([tq]))
    assert np.array_equal(xd[0], res["x_des"]) and np.array_equal(ud[0], res["u_des"]) and int(mode[0]) == int(res["mode"])
    wbc = q.load_wbc(model)
    wctx = q.WbcContext(model, wbc, 1)
    cmd, _ = wctx.update(xd, ud, s["rbd"][None], mode, np.array([period]), np.array([wtime]))
    assert np.array_equal(cmd[0], res["cmd"])
    ctx.close(); wctx.close()


# ---------------------------------------------------------------------------------------------------------------------------------
# Reference-typed adapters (include/qmb200_ocs2_adapters.hpp): B200SqpMpc : ocs2::MPC_BASE and B200HierarchicalWbc : qm::WbcBase
# installed by a QMController subclass through setupMpc(ros::NodeHandle&) / setupWbc(ros::NodeHandle&, const std::string&), compiled
# against mock headers that carry the reference's signatures (tests/cpp/mock; OCS2 / Eigen / ROS are not in this image).
@pytest.fixture(scope="module")
def controller(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("controller") / "controller_b200")
    libdir = os.path.join(ROOT, "qm_door_b200")
    subprocess.run(["make", "-C", ROOT, "qm_door_b200/libqmb200.so"], check=True, stdout=subprocess.DEVNULL)
    subprocess.run(["g++", "-std=c++17", "-Wall", "-Werror", "-O1", "-I" + os.path.join(ROOT, "tests", "cpp", "mock"), "-o", exe,
                    os.path.join(ROOT, "tests", "cpp", "controller_b200.cpp"), "-L" + libdir, "-lqmb200", "-Wl,-rpath," + libdir], check=True)
    return exe


def controller_scenario(q, t0):
    from qm_door_b200 import workload
    W = workload.Workload(1, horizon=1.0, dt=0.015, seed=12, t_span=t0 + 1.0, max_events=64)
    Ww = workload.WbcWorkload(1, seed=12)
    ne = int(W.nevents[0])
    return dict(x0=W.x0[0], events=W.events[0, :ne], modes=W.modes[0, :ne + 1], tt=W.target_t[0], tx=W.target_x[0], rbd=Ww.rbd[0])


def test_controller_subclass_builds_checks_terms_and_fails_loudly_without_gpu(controller, tmp_path):
    import qm_door_b200 as q
    s = controller_scenario(q, 11.0)
    write_scenario(tmp_path / "s.txt", 2, 11.0, 0.01, 11.0137, s["x0"], s["events"], s["modes"], s["tt"], s["tx"], s["rbd"], 0.002, 0.0)
    files = [q.DEFAULT_TASK, q.DEFAULT_URDF, q.DEFAULT_REFERENCE, str(tmp_path / "s.txt"), str(tmp_path / "r.txt")]
    # a term outside the eight names of QMInterface::setupOptimalControlProblem is refused before anything touches the device
    r = subprocess.run([controller] + files + ["foreign"], capture_output=True, text=True)
    assert r.returncode == 1 and "outside the eight names" in r.stderr, r.stderr
    if q.lib().qmb200_device_count() > 0:
        pytest.skip("a CUDA device is present: the run itself is covered by the GPU test")
    r = subprocess.run([controller] + files, capture_output=True, text=True)
    assert r.returncode == 1 and "no CUDA device" in r.stderr, r.stderr     # OCP accepted, files ingested, context creation threw


@pytest.mark.gpu
def test_controller_subclass_equals_ctypes_path(controller, tmp_path):
    """init -> starting -> MPC-thread tick -> update of the (mock) reference controller on top of B200SqpMpc / B200HierarchicalWbc:
    the policy, its evaluation and the WBC command equal the ctypes path on the same inputs bit for bit."""
    import qm_door_b200 as q
    from helpers import solver_for
    t0, cdt, tq, period, cycles = 11.0, 0.01, 11.0137, 0.002, 3
    s = controller_scenario(q, t0)
    write_scenario(tmp_path / "s.txt", cycles, t0, cdt, tq, s["x0"], s["events"], s["modes"], s["tt"], s["tx"], s["rbd"], period, 0.0)
    r = subprocess.run([controller, q.DEFAULT_TASK, q.DEFAULT_URDF, q.DEFAULT_REFERENCE, str(tmp_path / "s.txt"), str(tmp_path / "r.txt")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    res = read_result(tmp_path / "r.txt")
    model = q.load_model()
    problem, solver, _ = q.load_problem(model)
    sd = solver_for(solver, 1.0, 0.015, max_events=64, max_nodes=100)
    ctx = q.MpcContext(model, problem, sd, 1)
    E = sd.max_events
    ev = np.full((1, E), 1e30); ev[0, :len(s["events"])] = s["events"]
    md = np.full((1, E + 1), 15, dtype=np.int32); md[0, :len(s["modes"])] = s["modes"]
    ne = np.array([len(s["events"])], dtype=np.int32)
    for c in range(cycles):
        out = ctx.cycle(np.array([t0 + c * cdt]), s["x0"][None], ev, md, ne, s["tt"][None], s["tx"][None])
    n = int(out["n"][0])
    assert n == int(res["nodes"]) and np.array_equal(out["t"][0, :n], res["t"])
    assert np.array_equal(out["x"][0, :n].ravel(), res["x"]) and np.array_equal(out["u"][0, :n].ravel(), res["u"])
    xd, ud, mode = ctx.evaluate_policy(np.array([tq]))
    assert np.array_equal(xd[0], res["x_des"]) and np.array_equal(ud[0], res["u_des"]) and int(mode[0]) == int(res["mode"])
    wctx = q.WbcContext(model, q.load_wbc(model), 1)
    cmd, _ = wctx.update(xd, ud, s["rbd"][None], mode, np.array([period]), np.array([tq]))      # WBC time = observation time (QMController.cpp:147)
    assert np.array_equal(cmd[0], res["cmd"]) and np.array_equal(cmd[0, 36:], res["torque"])
    ctx.close(); wctx.close()
