// Host-side use of the library the way qm_controllers would (include/qmb200_adapters.hpp): the objects QMController::setupMpc /
// setupWbc build (QMController.cpp:273-307) and the calls QMController::update makes per tick (:116-157, :178-191).
// Reads one scenario from a text file, runs it through the C++ adapters and writes the results as text; tests/test_adapters.py
// compares them with the ctypes path on the same inputs. Without a CUDA device construction throws (no CPU fallback).
#include <cstdio>
#include <fstream>
#include <iostream>
#include "../../include/qmb200_adapters.hpp"

using qmb200::vector_t;

static vector_t read_vec(std::istream& in, size_t n) {
  vector_t v(n);
  for (auto& e : v) in >> e;
  return v;
}
static void write_vec(std::ostream& out, const char* name, const double* p, size_t n) {
  out << name << " " << n;
  char buf[40];
  for (size_t i = 0; i < n; ++i) { snprintf(buf, sizeof(buf), " %.17g", p[i]); out << buf; }
  out << "\n";
}

int main(int argc, char** argv) {
  if (argc != 6) { std::cerr << "usage: adapter_driver task.info robot.urdf reference.info scenario.txt result.txt\n"; return 2; }
  try {
    qmb200::InterfaceB200 itf(argv[1], argv[2], argv[3]);
    std::ifstream in(argv[4]);
    if (!in) throw std::invalid_argument("cannot open scenario file");
    int cycles, nev, nk;
    double t0, cycle_dt, tq;
    in >> cycles >> t0 >> cycle_dt >> tq;
    const vector_t x0 = read_vec(in, 30);
    in >> nev;
    qmb200::ModeScheduleB200 sched;
    sched.eventTimes = read_vec(in, nev);
    for (int i = 0; i <= nev; ++i) { int m; in >> m; sched.modeSequence.push_back(m); }
    in >> nk;
    qmb200::TargetTrajectoriesB200 target;
    target.timeTrajectory = read_vec(in, nk);
    for (int k = 0; k < nk; ++k) target.stateTrajectory.push_back(read_vec(in, QM_NTARGET));
    const vector_t rbd = read_vec(in, 55);
    double period, wbc_time;
    in >> period >> wbc_time;
    if (!in) throw std::invalid_argument("scenario file is incomplete");

    qmb200::SqpMpcB200 mpc(itf);                       // setupMpc
    qmb200::HierarchicalWbcB200 wbc(itf);              // setupWbc
    for (int c = 0; c < cycles; ++c) mpc.advanceMpc(t0 + c * cycle_dt, x0, sched, target);
    if (!mpc.initialPolicyReceived()) throw std::runtime_error("no policy");
    vector_t xs, us;
    size_t mode = 0;
    mpc.evaluatePolicy(tq, xs, us, mode);
    const vector_t cmd = wbc.update(xs, us, rbd, mode, period, wbc_time);

    std::ofstream out(argv[5]);
    const int n = mpc.numNodes();
    out << "nodes " << n << "\nmode " << mode << "\nalpha " << mpc.stepSize() << "\n";
    write_vec(out, "t", mpc.timeTrajectory().data(), n);
    write_vec(out, "x", mpc.stateTrajectory().data(), (size_t)n * 30);
    write_vec(out, "u", mpc.inputTrajectory().data(), (size_t)n * 30);
    write_vec(out, "x_des", xs.data(), 30);
    write_vec(out, "u_des", us.data(), 30);
    write_vec(out, "cmd", cmd.data(), 54);
    return 0;
  } catch (const std::exception& e) {
    std::cerr << "adapter_driver: " << e.what() << "\n";
    return 1;
  }
}
