"""ctypes mirror of qm_door_b200/csrc/qm_types.h (the POD descriptors of the C-ABI)."""
import ctypes as C

import numpy as np

NX, NU, NJ, NFEET, NUT, NCV, NTARGET = 30, 30, 24, 4, 18, 12, 37


class ModelDesc(C.Structure):
    _fields_ = [
        ("nj", C.c_int32),
        ("parent", C.c_int32 * NJ),
        ("jtype", C.c_int32 * NJ),
        ("depth", C.c_int32 * NJ),
        ("submask", C.c_uint32 * NJ),
        ("pathmask", C.c_uint32 * NJ),
        ("max_depth", C.c_int32),
        ("foot_joint", C.c_int32 * NFEET),
        ("ee_joint", C.c_int32),
        ("axis", C.c_double * 3 * NJ),
        ("Rp", C.c_double * 9 * NJ),
        ("pp", C.c_double * 3 * NJ),
        ("mass", C.c_double * NJ),
        ("com", C.c_double * 3 * NJ),
        ("inertia", C.c_double * 9 * NJ),
        ("foot_off", C.c_double * 3 * NFEET),
        ("ee_off", C.c_double * 3),
        ("ee_Roff", C.c_double * 9),
        ("total_mass", C.c_double),
        ("lower", C.c_double * NJ),
        ("upper", C.c_double * NJ),
        ("effort", C.c_double * NJ),
        ("root6_standard", C.c_int32),
        ("reserved", C.c_int32),
    ]


class ProblemDesc(C.Structure):
    _fields_ = [
        ("Q", C.c_double * (NX * NX)),
        ("R", C.c_double * (NU * NU)),
        ("mu_ee_pos", C.c_double), ("mu_ee_ori", C.c_double),
        ("mu_fee_pos", C.c_double), ("mu_fee_ori", C.c_double),
        ("fric_mu", C.c_double), ("fric_bar_mu", C.c_double), ("fric_bar_delta", C.c_double),
        ("fric_reg", C.c_double), ("fric_grip", C.c_double), ("fric_hess_shift", C.c_double),
        ("pos_bar_mu", C.c_double), ("pos_bar_delta", C.c_double),
        ("vel_bar_mu", C.c_double), ("vel_bar_delta", C.c_double),
        ("arm_pos_lo", C.c_double * 6), ("arm_pos_hi", C.c_double * 6),
        ("arm_vel_lo", C.c_double * 6), ("arm_vel_hi", C.c_double * 6),
        ("box_offset", C.c_double),
        ("swing_liftoff_vel", C.c_double), ("swing_touchdown_vel", C.c_double),
        ("swing_height", C.c_double), ("swing_time_scale", C.c_double),
        ("gravity", C.c_double),
    ]


class SolverDesc(C.Structure):
    _fields_ = [
        ("dt", C.c_double), ("horizon", C.c_double),
        ("delta_tol", C.c_double), ("g_max", C.c_double), ("g_min", C.c_double),
        ("alpha_decay", C.c_double), ("alpha_min", C.c_double), ("gamma_c", C.c_double),
        ("armijo_factor", C.c_double),
        ("weak_eps", C.c_double), ("dt_min", C.c_double),
        ("max_nodes", C.c_int32), ("max_events", C.c_int32), ("max_targets", C.c_int32),
        ("sqp_iterations", C.c_int32),
        ("cost_tol", C.c_double),
    ]


class WbcDesc(C.Structure):
    _fields_ = [
        ("kp_swing", C.c_double), ("kd_swing", C.c_double),
        ("kp_base_height", C.c_double), ("kd_base_height", C.c_double),
        ("kp_base_linear", C.c_double), ("kd_base_linear", C.c_double),
        ("kp_base_angular", C.c_double), ("kd_base_angular", C.c_double),
        ("kp_arm_joint", C.c_double * 6), ("kd_arm_joint", C.c_double * 6),
        ("kp_ee_linear", C.c_double * 3), ("kd_ee_linear", C.c_double * 3),
        ("kp_ee_angular", C.c_double * 3), ("kd_ee_angular", C.c_double * 3),
        ("friction_mu", C.c_double),
        ("tau_max", C.c_double * 18),
        ("swing_weight", C.c_double), ("init_time", C.c_double), ("gravity", C.c_double),
        ("mpc_variant", C.c_int32), ("reserved", C.c_int32),
    ]


class TargetDesc(C.Structure):
    _fields_ = [
        ("com_height", C.c_double), ("feet_height", C.c_double), ("arm_dist", C.c_double), ("time_to_target", C.c_double),
        ("target_displacement_velocity", C.c_double), ("target_rotation_velocity", C.c_double),
        ("default_joint_state", C.c_double * 18),
    ]


class ActuatorDesc(C.Structure):
    _fields_ = [
        ("leg_kp", C.c_double), ("leg_kd", C.c_double), ("arm_kp", C.c_double), ("arm_kd", C.c_double),
        ("leg_enable_time", C.c_double), ("delay_ns", C.c_int64),
    ]


ACT_CAPACITY = 32


def _set(arr, values):
    a = np.ctypeslib.as_array(arr)
    a[...] = np.asarray(values).reshape(a.shape)


def struct_to_dict(s):
    out = {}
    for name, _ in s._fields_:
        v = getattr(s, name)
        out[name] = np.ctypeslib.as_array(v).copy() if hasattr(v, "_length_") else v
    return out
