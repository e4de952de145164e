// A QMController subclass that installs the B200 hot path through the reference's own factory hooks, compiled against the mock
// headers under tests/cpp/mock (same signatures as qm_controllers/include/qm_controllers/QMController.h:50-54,62-80 and
// qm_wbc/include/qm_wbc/WbcBase.h:28-34) and driven through the base class's unchanged init / starting / MPC-thread tick /
// update. This is the class INTEGRATION.md shows; tests/test_adapters.py builds it with -Wall -Werror, checks the OCP term
// inspection, and on the GPU compares its outputs with the ctypes path bit for bit.
#include <cstdio>
#include <fstream>
#include <iostream>
#include <qm_controllers/QMController.h>
#include "../../include/qmb200_ocs2_adapters.hpp"

namespace qm {

class QMControllerB200 : public QMController {
 protected:
  void setupInterface(const std::string& taskFile, const std::string& urdfFile, const std::string& referenceFile, bool verbose) override {
    QMController::setupInterface(taskFile, urdfFile, referenceFile, verbose);   // the reference's QMInterface: settings, reference manager, OCP
    b200Files_ = std::make_shared<qmb200::InterfaceB200>(taskFile, urdfFile, referenceFile);
  }
  // QMController.cpp:287-307 with the solver object exchanged; the gait receiver / ROS reference manager lines stay as they are
  void setupMpc(ros::NodeHandle& controller_nh) override {
    (void)controller_nh;
    mpc_ = std::make_shared<qmb200::B200SqpMpc>(qmInterface_->mpcSettings(), qmInterface_->sqpSettings(), qmInterface_->getOptimalControlProblem(),
                                                 qmInterface_->getInitializer(), *b200Files_);
    rbdConversions_ = std::make_shared<CentroidalModelRbdConversions>(qmInterface_->getPinocchioInterface(), qmInterface_->getCentroidalModelInfo());
    mpc_->getSolverPtr()->setReferenceManager(qmInterface_->getReferenceManagerPtr());
  }
  // QMController.cpp:273-277 with the WBC object exchanged
  void setupWbc(ros::NodeHandle& controller_nh, const std::string& taskFile) override {
    wbc_ = std::make_shared<qmb200::B200HierarchicalWbc>(qmInterface_->getPinocchioInterface(), qmInterface_->getCentroidalModelInfo(),
                                                          *eeKinematicsPtr_, *armEeKinematicsPtr_, controller_nh, *b200Files_);
    wbc_->loadTasksSetting(taskFile, true);
  }

 public:
  // test hooks: what the gait receiver / the parameter server provide in the running system
  void setModeSchedule(const ModeSchedule& m) { qmInterface_->getReferenceManagerPtr()->setModeSchedule(m); }
  void configure(scalar_t horizon, scalar_t dt) { horizon_ = horizon; dt_ = dt; }
  void addForeignTerm() { foreign_ = true; }
  const qmb200::B200SqpSolver& solver() const { return *static_cast<const qmb200::B200SqpMpc&>(*mpc_).getSolverPtr(); }

 protected:
  scalar_t horizon_ = 1.0, dt_ = 0.015;
  bool foreign_ = false;

 private:
  std::shared_ptr<qmb200::InterfaceB200> b200Files_;

 public:
  // mock-only: the settings QMInterface parses from task.info (sqp.dt, mpc.timeHorizon) and an optional foreign OCP term
  void applyMockSettings() {
    qmInterface_->mpcSettings_.timeHorizon_ = horizon_;
    qmInterface_->sqpSettings_.dt = dt_;
    qmInterface_->sqpSettings_.deltaTol = 1e-4; qmInterface_->sqpSettings_.g_max = 1e-2; qmInterface_->sqpSettings_.g_min = 1e-6;   // task.info:81-83
    qmInterface_->sqpSettings_.useFeedbackPolicy = false;                                                                            // task.info:90
    if (foreign_) { struct X : StateInputCost {}; qmInterface_->mutableProblem().costPtr->add("userTerm", std::unique_ptr<StateInputCost>(new X)); }
  }
  bool initWithMockSettings(ros::NodeHandle& nh, const std::string& taskFile, const std::string& urdfFile, const std::string& referenceFile) {
    setupInterface(taskFile, urdfFile, referenceFile, false);
    applyMockSettings();
    setupMpc(nh);
    setupMrt();
    setupWbc(nh, taskFile);
    return true;
  }
};

}  // namespace qm

static ocs2::vector_t read_vec(std::istream& in, size_t n) {
  ocs2::vector_t v(static_cast<long>(n));
  for (size_t i = 0; i < n; ++i) in >> v[static_cast<long>(i)];
  return v;
}
static void write_vec(std::ostream& out, const char* name, const double* p, size_t n) {
  out << name << " " << n;
  char buf[40];
  for (size_t i = 0; i < n; ++i) { snprintf(buf, sizeof(buf), " %.17g", p[i]); out << buf; }
  out << "\n";
}

int main(int argc, char** argv) {
  if (argc < 6) { std::cerr << "usage: controller_b200 task.info robot.urdf reference.info scenario.txt result.txt [foreign]\n"; return 2; }
  try {
    std::ifstream in(argv[4]);
    if (!in) throw std::invalid_argument("cannot open scenario file");
    int cycles, nev, nk;
    double t0, cycle_dt, tq;
    in >> cycles >> t0 >> cycle_dt >> tq;
    const ocs2::vector_t x0 = read_vec(in, 30);
    in >> nev;
    ocs2::ModeSchedule sched;
    for (int i = 0; i < nev; ++i) { double e; in >> e; sched.eventTimes.push_back(e); }
    for (int i = 0; i <= nev; ++i) { size_t m; in >> m; sched.modeSequence.push_back(m); }
    in >> nk;
    ocs2::TargetTrajectories target;
    for (int k = 0; k < nk; ++k) { double t; in >> t; target.timeTrajectory.push_back(t); }
    for (int k = 0; k < nk; ++k) { target.stateTrajectory.push_back(read_vec(in, QM_NTARGET)); target.inputTrajectory.push_back(ocs2::vector_t(30)); }
    const ocs2::vector_t rbd = read_vec(in, 55);
    double period, wbc_time;
    in >> period >> wbc_time;
    if (!in) throw std::invalid_argument("scenario file is incomplete");

    qm::QMControllerB200 ctl;
    ros::NodeHandle nh;
    ctl.configure(1.0, 0.015);
    if (argc > 6 && std::string(argv[6]) == "foreign") ctl.addForeignTerm();
    ctl.initWithMockSettings(nh, argv[1], argv[2], argv[3]);         // QMController::init: setupInterface, setupMpc, setupMrt, setupWbc
    ctl.setModeSchedule(sched);
    ctl.starting(t0, x0, target);                                      // first cycle (waits for the initial policy)
    for (int c = 1; c < cycles; ++c) { ctl.observe(t0 + c * cycle_dt, x0); ctl.mpcThreadTick(); }
    ocs2::vector_t xs, us;
    size_t mode = 0;
    const ocs2::vector_t cmd = ctl.update(tq, period, x0, rbd, &xs, &us, &mode);   // evaluatePolicy + wbc_->update

    std::ofstream out(argv[5]);
    const auto& core = ctl.solver().core();
    const int n = core.numNodes();
    out << "nodes " << n << "\nmode " << mode << "\nalpha " << core.stepSize() << "\n";
    write_vec(out, "t", core.timeTrajectory().data(), n);
    write_vec(out, "x", core.stateTrajectory().data(), (size_t)n * 30);
    write_vec(out, "u", core.inputTrajectory().data(), (size_t)n * 30);
    write_vec(out, "x_des", xs.data(), 30);
    write_vec(out, "u_des", us.data(), 30);
    write_vec(out, "cmd", cmd.data(), 54);
    write_vec(out, "torque", cmd.tail(18).data(), 18);
    return 0;
  } catch (const std::exception& e) {
    std::cerr << "controller_b200: " << e.what() << "\n";
    return 1;
  }
}
