#!/usr/bin/env python
"""Regenerate quadruped_locomotion_b200/models/*.json and include/qlb_models.h
from the reference URDFs (run in the build container, where /root/reference exists).

The tables are numeric robot parameters (joint origins, link masses, centres of mass);
no reference source code is copied.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from quadruped_locomotion_b200 import legmodel  # noqa: E402

REF = os.environ.get("QLB_REFERENCE", "/root/reference")
# per-leg URDFs read by MyRobotSolver::loadLimbModelFromURDF (single_leg_test/lib/model_test_header.cpp:224-247)
LIMB_URDFS = {
    "quadruped_model": "quadruped_model/urdf/quadruped_model_%s_leg.urdf",
    "simpledog": "quadruped_model/urdf/simpledog_%s_leg.urdf",
}
URDFS = {
    "quadruped_model": "quadruped_model/urdf/quadruped_model.urdf",   # loaded by QK.cpp:21
    "simpledog": "quadruped_model/urdf/simpledog.urdf",               # the robot BASELINE names
}


def main():
    models = {}
    for name, rel in URDFS.items():
        m = legmodel.parse_urdf(os.path.join(REF, rel))
        m["limb_dynamics"] = [legmodel.parse_limb_dynamics(os.path.join(REF, LIMB_URDFS[name] % leg))
                              for leg in ("lf", "rf", "rh", "lh")]
        with open(os.path.join(legmodel.MODELS_DIR, name + ".json"), "w") as f:
            json.dump(m, f, indent=1)
        models["QLB_MODEL_" + name.upper()] = m
        print(name, "ok:", [leg["joint_names"][0] for leg in m["legs"]])
    legmodel.emit_c_header(models, os.path.join(ROOT, "include", "qlb_models.h"))


if __name__ == "__main__":
    main()
