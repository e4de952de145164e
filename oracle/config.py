"""Oracle-side ingestion of the reference's input files (TEST INFRASTRUCTURE).

* parse_info      : Boost.PropertyTree INFO format, as read through ocs2::loadData
                    at qm_interface/src/QMInterface.cpp:65-73,85,152-156,199-234,291,306,395-397
* load_matrix     : ocs2 loadEigenMatrix semantics ("(i,j) value" entries, optional "scaling")
* load_urdf_model : URDF -> tree of 24 one-DoF joints (3 prismatic + ZYX Euler root, 18 revolute),
                    fixed links merged into their parent body, joint order = Pinocchio's URDF
                    traversal (children sorted by link name), restating
                    centroidal_model::createPinocchioInterface called at QMInterface.cpp:408-416.
* load_problem    : every constant of SURVEY.md App. A the hot path consumes.
"""
import math
import os
import re
import xml.etree.ElementTree as ET

import numpy as np

MODE_NAMES = {  # [upstream] ocs2_legged_robot ModeNumber: LF=8 RF=4 LH=2 RH=1 stance bits
    "FLY": 0, "RH": 1, "LH": 2, "LH_RH": 3, "RF": 4, "RF_RH": 5, "RF_LH": 6, "RF_LH_RH": 7,
    "LF": 8, "LF_RH": 9, "LF_LH": 10, "LF_LH_RH": 11, "LF_RF": 12, "LF_RF_RH": 13,
    "LF_RF_LH": 14, "STANCE": 15,
}

JOINT_NAMES = ["LF_HAA", "LF_HFE", "LF_KFE", "RF_HAA", "RF_HFE", "RF_KFE",
               "LH_HAA", "LH_HFE", "LH_KFE", "RH_HAA", "RH_HFE", "RH_KFE",
               "z1_joint_1", "z1_joint_2", "z1_joint_3", "z1_joint_4", "z1_joint_5",
               "z1_joint_6"]  # ModelSettings.h:32-36 (joints NOT listed get fixed)
CONTACT_NAMES = ["LF_FOOT", "RF_FOOT", "LH_FOOT", "RH_FOOT"]  # ModelSettings.h:38


# ----------------------------------------------------------------------------- INFO
def _tokenize(text):
    toks = []
    for line in text.splitlines():
        line = line.split(";", 1)[0]
        i = 0
        n = len(line)
        while i < n:
            c = line[i]
            if c.isspace():
                i += 1
            elif c in "{}":
                toks.append(c)
                i += 1
            elif c == '"':
                j = line.index('"', i + 1)
                toks.append(line[i + 1:j])
                i = j + 1
            else:
                j = i
                while j < n and not line[j].isspace() and line[j] not in "{}":
                    j += 1
                toks.append(line[i:j])
                i = j
        toks.append("\n")
    return toks


def parse_info(path):
    """INFO -> nested dict {key: (value_str, children_dict)} flattened to dicts; leaf = str."""
    with open(path) as fh:
        toks = _tokenize(fh.read())
    pos = 0

    def block():
        nonlocal pos
        out = {}
        while pos < len(toks):
            t = toks[pos]
            if t == "\n":
                pos += 1
                continue
            if t == "}":
                pos += 1
                return out
            key = t
            pos += 1
            val = None
            if pos < len(toks) and toks[pos] not in ("\n", "{", "}"):
                val = toks[pos]
                pos += 1
            while pos < len(toks) and toks[pos] == "\n":
                pos += 1
            if pos < len(toks) and toks[pos] == "{":
                pos += 1
                out[key] = block()
            else:
                out[key] = val
        return out

    return block()


def get(tree, dotted):
    node = tree
    for k in dotted.split("."):
        node = node[k]
    return node


def load_matrix(tree, name, rows, cols):
    """ocs2::loadData::loadEigenMatrix: zero matrix, '(i,j) v' entries, times 'scaling' if present."""
    node = get(tree, name)
    m = np.zeros((rows, cols))
    scaling = float(node.get("scaling", 1.0))
    for k, v in node.items():
        mm = re.fullmatch(r"\((\d+),(\d+)\)", k)
        if mm:
            m[int(mm.group(1)), int(mm.group(2))] = float(v)
    return m * scaling


def load_list(tree, name):
    node = get(tree, name)
    items = sorted(((int(k[1:-1]), v) for k, v in node.items() if k.startswith("[")), key=lambda kv: kv[0])
    return [v for _, v in items]


# ----------------------------------------------------------------------------- URDF
def _rpy_to_R(r, p, y):
    cr, sr, cp, sp, cy, sy = math.cos(r), math.sin(r), math.cos(p), math.sin(p), math.cos(y), math.sin(y)
    Rx = np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]])
    Ry = np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]])
    Rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


def _vec(s, n=3):
    v = [float(t) for t in s.split()]
    assert len(v) == n
    return np.array(v)


def _origin(el):
    o = el.find("origin") if el is not None else None
    if o is None:
        return np.eye(3), np.zeros(3)
    return _rpy_to_R(*_vec(o.get("rpy", "0 0 0"))), _vec(o.get("xyz", "0 0 0"))


class RobotModel:
    """Tree of NJ one-DoF joints. All arrays are plain float64 numpy.

    parent[j]   : parent joint index (-1 = world)
    jtype[j]    : 0 prismatic, 1 revolute
    axis[j]     : joint axis in the joint (= child body) frame
    Rp[j], pp[j]: placement of joint frame j in parent body frame
    mass[j], com[j], inertia[j] : merged rigid body attached to joint j (body frame, inertia about com)
    """

    def __init__(self):
        self.nj = 0
        self.parent, self.jtype, self.axis, self.Rp, self.pp = [], [], [], [], []
        self.mass, self.com, self.inertia = [], [], []
        self.names = []
        self.frames = {}  # name -> (joint idx, offset[3], Roff[3,3])
        self.lower = []
        self.upper = []
        self.effort = []

    def finalize(self):
        for k in ("parent", "jtype"):
            setattr(self, k, np.array(getattr(self, k), dtype=np.int64))
        for k in ("axis", "Rp", "pp", "mass", "com", "inertia", "lower", "upper", "effort"):
            setattr(self, k, np.array(getattr(self, k), dtype=np.float64))
        self.nj = len(self.parent)
        self.total_mass = float(self.mass.sum())
        # path[i, j] = 1 if joint j is an ancestor-or-self of joint i
        self.path = np.zeros((self.nj, self.nj), dtype=bool)
        for i in range(self.nj):
            j = i
            while j >= 0:
                self.path[i, j] = True
                j = self.parent[j]
        self.foot_joint = np.array([self.frames[n][0] for n in CONTACT_NAMES])
        self.foot_off = np.array([self.frames[n][1] for n in CONTACT_NAMES])
        self.ee_joint, self.ee_off, self.ee_Roff = self.frames["z1_end_effector"]
        return self


def load_urdf_model(urdf_path, joint_names=JOINT_NAMES):
    root = ET.parse(urdf_path).getroot()
    links = {l.get("name"): l for l in root.findall("link")}
    joints_by_parent = {}
    child_links = set()
    for j in root.findall("joint"):
        joints_by_parent.setdefault(j.find("parent").get("link"), []).append(j)
        child_links.add(j.find("child").get("link"))
    root_links = [n for n in links if n not in child_links]
    assert len(root_links) == 1, root_links
    base = root_links[0]

    m = RobotModel()
    # floating base: composite translation + spherical ZYX (Euler rates as velocities)  [upstream]
    for k, (jt, ax) in enumerate([(0, (1, 0, 0)), (0, (0, 1, 0)), (0, (0, 0, 1)),
                                  (1, (0, 0, 1)), (1, (0, 1, 0)), (1, (1, 0, 0))]):
        m.parent.append(k - 1)
        m.jtype.append(jt)
        m.axis.append(np.array(ax, dtype=float))
        m.Rp.append(np.eye(3))
        m.pp.append(np.zeros(3))
        m.mass.append(0.0)
        m.com.append(np.zeros(3))
        m.inertia.append(np.zeros((3, 3)))
        m.names.append("root_%d" % k)
        m.lower.append(-np.inf)
        m.upper.append(np.inf)
        m.effort.append(0.0)

    bodies = {5: []}  # joint idx -> list of (mass, com_in_body, I_com_in_body)

    def add_link_inertia(jidx, link_name, R, p):
        link = links[link_name]
        inert = link.find("inertial")
        m.frames[link_name] = (jidx, p.copy(), R.copy())
        if inert is None:
            return
        Ro, po = _origin(inert)
        mass = float(inert.find("mass").get("value"))
        a = inert.find("inertia").attrib
        I = np.array([[float(a["ixx"]), float(a["ixy"]), float(a["ixz"])],
                      [float(a["ixy"]), float(a["iyy"]), float(a["iyz"])],
                      [float(a["ixz"]), float(a["iyz"]), float(a["izz"])]])
        Rb = R @ Ro
        bodies[jidx].append((mass, p + R @ po, Rb @ I @ Rb.T))

    def visit(link_name, jidx, R, p):
        """link rigidly attached to joint jidx's body at placement (R, p)."""
        add_link_inertia(jidx, link_name, R, p)
        for j in sorted(joints_by_parent.get(link_name, []), key=lambda e: e.find("child").get("link")):
            Rj, pj = _origin(j)
            Rc, pc = R @ Rj, p + R @ pj
            child = j.find("child").get("link")
            jt = j.get("type")
            if jt == "fixed" or j.get("name") not in joint_names:
                visit(child, jidx, Rc, pc)  # locked at neutral (q = 0)
            else:
                assert jt in ("revolute", "continuous")
                new = len(m.parent)
                m.parent.append(jidx)
                m.jtype.append(1)
                m.axis.append(_vec(j.find("axis").get("xyz")))
                m.Rp.append(Rc)
                m.pp.append(pc)
                m.names.append(j.get("name"))
                lim = j.find("limit")
                m.lower.append(float(lim.get("lower")))
                m.upper.append(float(lim.get("upper")))
                m.effort.append(float(lim.get("effort")))
                m.mass.append(0.0)
                m.com.append(np.zeros(3))
                m.inertia.append(np.zeros((3, 3)))
                bodies[new] = []
                visit(child, new, np.eye(3), np.zeros(3))

    visit(base, 5, np.eye(3), np.zeros(3))
    for jidx, parts in bodies.items():
        mt = sum(q[0] for q in parts)
        if mt == 0:
            continue
        c = sum(q[0] * q[1] for q in parts) / mt
        I = np.zeros((3, 3))
        for (mi, ci, Ii) in parts:
            d = ci - c
            I += Ii + mi * (d.dot(d) * np.eye(3) - np.outer(d, d))
        m.mass[jidx], m.com[jidx], m.inertia[jidx] = mt, c, I
    return m.finalize()


# ----------------------------------------------------------------------------- problem constants
class Problem:
    pass


def load_problem(task_path, reference_path, gait_path, model):
    """SURVEY.md App. A constants.  Citations: task.info line numbers of the reference."""
    t = parse_info(task_path)
    r = parse_info(reference_path)
    g = parse_info(gait_path)
    P = Problem()
    P.nx, P.nu = 30, 30
    P.x_init = load_matrix(t, "initialState", 30, 1)[:, 0]            # task.info:151-190
    P.Q = load_matrix(t, "Q", 30, 30)                                  # :193-234
    P.R_task = load_matrix(t, "R", 30, 30)                             # :249-288
    P.mu_ee_pos = float(get(t, "endEffector.muPosition"))              # :236-240
    P.mu_ee_ori = float(get(t, "endEffector.muOrientation"))
    P.mu_fee_pos = float(get(t, "finalEndEffector.muPosition"))        # :241-246
    P.mu_fee_ori = float(get(t, "finalEndEffector.muOrientation"))
    P.fric_mu = float(get(t, "frictionConeSoftConstraint.frictionCoefficient"))  # :291-298
    P.fric_bar_mu = float(get(t, "frictionConeSoftConstraint.mu"))
    P.fric_bar_delta = float(get(t, "frictionConeSoftConstraint.delta"))
    P.fric_reg = 25.0          # [upstream] FrictionConeConstraint::Config defaults (ctor gets only mu, QMInterface.cpp:346)
    P.fric_grip = 0.0
    P.fric_hess_shift = 1e-6
    P.pos_bar_mu = float(get(t, "jointPositionLimits.mu"))             # :300-308
    P.pos_bar_delta = float(get(t, "jointPositionLimits.delta"))
    P.vel_bar_mu = float(get(t, "jointVelocityLimits.mu"))             # :309-316
    P.vel_bar_delta = float(get(t, "jointVelocityLimits.delta"))
    P.arm_vel_lo = load_matrix(t, "jointVelocityLimits.lowerBound.arm", 6, 1)[:, 0]
    P.arm_vel_hi = load_matrix(t, "jointVelocityLimits.upperBound.arm", 6, 1)[:, 0]
    P.arm_pos_lo = model.lower[-6:].copy()                             # QMInterface.cpp:199-200
    P.arm_pos_hi = model.upper[-6:].copy()
    P.pos_err_gain = float(get(t, "model_settings.positionErrorGain"))
    P.phase_transition_stance_time = float(get(t, "model_settings.phaseTransitionStanceTime"))
    sw = get(t, "swing_trajectory_config")                              # :24-31
    P.swing = {k: float(sw[k]) for k in ("liftOffVelocity", "touchDownVelocity", "swingHeight",
                                         "touchdownAfterHorizon", "swingTimeScale")}
    s = get(t, "sqp")                                                   # :76-93
    P.sqp = dict(dt=float(s["dt"]), sqpIteration=int(s["sqpIteration"]), deltaTol=float(s["deltaTol"]),
                 g_max=float(s["g_max"]), g_min=float(s["g_min"]),
                 alpha_decay=0.5, alpha_min=1e-4, gamma_c=1e-6, armijoFactor=1e-4)  # [upstream] sqp::Settings defaults
    P.time_horizon = float(get(t, "mpc.timeHorizon"))
    P.friction_wbc = float(get(t, "frictionConeTask.frictionCoefficient"))  # :347-350
    P.com_height = float(r["comHeight"])
    P.target_displacement_velocity = float(r["targetDisplacementVelocity"])     # QmTargetTrajectoriesPublisher_node.cpp:268-272
    P.target_rotation_velocity = float(r["targetRotationVelocity"])
    P.default_joint_state = load_matrix(r, "defaultJointState", 18, 1)[:, 0]
    P.gaits = {}
    P.gait_list = load_list(g, "list")
    for name in P.gait_list:
        P.gaits[name] = dict(modes=[MODE_NAMES[mname] for mname in load_list(g, name + ".modeSequence")],
                             times=[float(v) for v in load_list(g, name + ".switchingTimes")])
    P.init_mode_schedule = dict(
        modes=[MODE_NAMES[mname] for mname in load_list(r, "initialModeSchedule.modeSequence")],
        events=[float(v) for v in load_list(r, "initialModeSchedule.eventTimes")])
    return P


# ----------------------------------------------------------------------------- fixtures (the GPU box has no /root/reference)
def dump_fixture(model, P, path):
    """Write the parsed model + problem constants as JSON (generated by tools/gen_fixtures.py in the build container)."""
    import json
    md = {k: getattr(model, k).tolist() for k in ("parent", "jtype", "axis", "Rp", "pp", "mass", "com", "inertia",
                                                   "lower", "upper", "effort")}
    md["lower"] = [(-1e30 if not np.isfinite(v) else v) for v in md["lower"]]
    md["upper"] = [(1e30 if not np.isfinite(v) else v) for v in md["upper"]]
    md["names"] = model.names
    md["frames"] = {k: [int(v[0]), np.asarray(v[1]).tolist(), np.asarray(v[2]).tolist()] for k, v in model.frames.items()}
    pd = {}
    for k, v in vars(P).items():
        pd[k] = v.tolist() if isinstance(v, np.ndarray) else v
    with open(path, "w") as fh:
        json.dump(dict(model=md, problem=pd), fh, indent=0)


def load_fixture(path):
    import json
    with open(path) as fh:
        d = json.load(fh)
    m = RobotModel()
    for k in ("parent", "jtype", "axis", "Rp", "pp", "mass", "com", "inertia", "lower", "upper", "effort"):
        setattr(m, k, d["model"][k])
    m.names = d["model"]["names"]
    m.frames = {k: (int(v[0]), np.array(v[1]), np.array(v[2])) for k, v in d["model"]["frames"].items()}
    m.finalize()
    P = Problem()
    for k, v in d["problem"].items():
        setattr(P, k, np.array(v) if isinstance(v, list) and k not in ("gait_list",) else v)
    return m, P


REFERENCE_ROOT = "/root/reference"


def load_default():
    """Model + problem from the reference checkout when present (build container), else from the committed fixture."""
    import os
    here = os.path.dirname(os.path.abspath(__file__))
    fixture = os.path.join(here, "..", "tests", "golden", "oracle_inputs.json")
    urdf = os.path.join(REFERENCE_ROOT, "qm_description/urdf/quadruped_manipulator/robot.urdf")
    if os.path.exists(urdf):
        m = load_urdf_model(urdf)
        cfg = os.path.join(REFERENCE_ROOT, "qm_controllers/config")
        P = load_problem(os.path.join(cfg, "task.info"), os.path.join(cfg, "reference.info"), os.path.join(cfg, "gait.info"), m)
        return m, P
    return load_fixture(fixture)
