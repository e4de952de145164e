"""Generate the committed fixtures from the reference checkout (run in the build container only).

  tests/golden/oracle_inputs.json   parsed robot model + problem constants for the oracle on the GPU box
  qm_door_b200/data/*.urdf|*.info   normalised input files for the product's C++ loaders: only the kinematic /
                                    inertial URDF content and the INFO keys the hot path reads, re-emitted from the parse
                                    (not copies of the reference files)
Usage: python tools/gen_fixtures.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import config  # noqa: E402

REF = "/root/reference"


def emit_urdf(path_in, path_out):
    import xml.etree.ElementTree as ET
    r = ET.parse(path_in).getroot()
    out = ['<?xml version="1.0"?>', '<robot name="%s">' % r.get("name", "robot")]
    for l in r.findall("link"):
        i = l.find("inertial")
        if i is None:
            out.append('  <link name="%s"/>' % l.get("name"))
            continue
        o = i.find("origin")
        a = i.find("inertia").attrib
        out.append('  <link name="%s">' % l.get("name"))
        out.append('    <inertial>')
        if o is not None:
            out.append('      <origin xyz="%s" rpy="%s"/>' % (o.get("xyz", "0 0 0"), o.get("rpy", "0 0 0")))
        out.append('      <mass value="%s"/>' % i.find("mass").get("value"))
        out.append('      <inertia ' + " ".join('%s="%s"' % (k, a[k]) for k in ("ixx", "ixy", "ixz", "iyy", "iyz", "izz")) + '/>')
        out.append('    </inertial>')
        out.append('  </link>')
    for j in r.findall("joint"):
        out.append('  <joint name="%s" type="%s">' % (j.get("name"), j.get("type")))
        o = j.find("origin")
        if o is not None:
            out.append('    <origin xyz="%s" rpy="%s"/>' % (o.get("xyz", "0 0 0"), o.get("rpy", "0 0 0")))
        out.append('    <parent link="%s"/>' % j.find("parent").get("link"))
        out.append('    <child link="%s"/>' % j.find("child").get("link"))
        a = j.find("axis")
        if a is not None:
            out.append('    <axis xyz="%s"/>' % a.get("xyz"))
        lim = j.find("limit")
        if lim is not None:
            out.append('    <limit ' + " ".join('%s="%s"' % (k, lim.get(k)) for k in ("lower", "upper", "effort", "velocity") if lim.get(k) is not None) + '/>')
        out.append('  </joint>')
    out.append('</robot>')
    with open(path_out, "w") as fh:
        fh.write("\n".join(out) + "\n")


def emit_info(tree, keys, path_out):
    def walk(node, ind, out):
        for k, v in node.items():
            if isinstance(v, dict):
                out.append("%s%s\n%s{" % (ind, k, ind))
                walk(v, ind + "  ", out)
                out.append("%s}" % ind)
            else:
                out.append("%s%s %s" % (ind, k, v if (v and " " not in v) else '"%s"' % v))
    out = []
    walk({k: tree[k] for k in keys if k in tree}, "", out)
    with open(path_out, "w") as fh:
        fh.write("\n".join(out) + "\n")


def main():
    urdf = os.path.join(REF, "qm_description/urdf/quadruped_manipulator/robot.urdf")
    cfg = os.path.join(REF, "qm_controllers/config")
    m = config.load_urdf_model(urdf)
    P = config.load_problem(os.path.join(cfg, "task.info"), os.path.join(cfg, "reference.info"), os.path.join(cfg, "gait.info"), m)
    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
    config.dump_fixture(m, P, os.path.join(ROOT, "tests", "golden", "oracle_inputs.json"))
    data = os.path.join(ROOT, "qm_door_b200", "data")
    os.makedirs(data, exist_ok=True)
    emit_urdf(urdf, os.path.join(data, "aliengo_z1.urdf"))
    t = config.parse_info(os.path.join(cfg, "task.info"))
    emit_info(t, ["centroidalModelType", "model_settings", "swing_trajectory_config", "sqp", "mpc", "initialState", "Q", "R",
                  "endEffector", "finalEndEffector", "frictionConeSoftConstraint", "jointPositionLimits", "jointVelocityLimits",
                  "frictionConeTask"], os.path.join(data, "aliengo_z1_task.info"))
    r = config.parse_info(os.path.join(cfg, "reference.info"))
    emit_info(r, list(r.keys()), os.path.join(data, "aliengo_z1_reference.info"))
    g = config.parse_info(os.path.join(cfg, "gait.info"))
    emit_info(g, list(g.keys()), os.path.join(data, "aliengo_z1_gait.info"))
    print("fixtures written")


if __name__ == "__main__":
    main()
