"""CPU oracle for the qm_door MPC + WBC hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product path (qm_door_b200/, the
C-ABI library) may import, call or link anything in this directory; only
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs use it, and only as the checker or the timed CPU baseline.

PARITY UNPINNED: the reference (danisotelo/qm_door @ 009a37d) ships no tests,
golden vectors or recorded outputs for this path, and its numeric engines
(OCS2, Pinocchio, CppAD, HPIPM, qpOASES) are not vendored and cannot be built
in this image (no Eigen, Boost, ROS, network).  The oracle is therefore a
restatement of the reference's problem definition (files cited per function,
relative to /root/reference) plus the published upstream algorithms, trusted
through self-consistency checks (tests/test_oracle_*.py): finite-difference /
complex-step identities, conservation identities, dense-KKT == Riccati,
KKT residuals of the hierarchical QP.

Design: NumPy FP64, vectorised over arbitrary leading batch dimensions, and
complex-safe so every derivative the reference obtains from CppAD is obtained
here by complex-step differentiation of the same value function (independent
of the analytic derivatives the CUDA path uses).
"""
