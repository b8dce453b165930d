#!/usr/bin/env python
"""Generate tests/golden/*.npz|json.  Run in the build container (needs /root/reference for oracle/_ref).

golden_states.npz: a few hundred synthetic states (all BASELINE configs + edge cases) and the outputs of
the pipeline with the REFERENCE's own vendored QuadProg++ (oracle/_ref) as the QP solver.  The oracle's
own solver and the GPU are both tested against these numbers.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from quadruped_locomotion_b200 import legmodel, synth  # noqa: E402


def cat(parts):
    return {k: np.concatenate([p[k] for p in parts], axis=-1) for k in parts[0]}


def main():
    assert O.have_ref(), "oracle/_ref not built: /root/reference missing?"
    parts = [synth.make_states("C1"), synth.make_states("C2", 96), synth.make_states("C3", 160),
             synth.make_states("C5", 160, start=123456)]
    # edge cases: one stance leg each, no stance, heavy lateral load (friction-limited), tilted normals
    e = synth.make_states("C3", 16, start=777)
    e["mask"] = np.array([1, 2, 4, 8, 0, 0, 3, 5, 6, 9, 10, 12, 7, 11, 13, 14], dtype=np.uint8)
    parts.append(e)
    f = synth.make_states("C3", 16, start=999)
    f["mask"][:] = 0xF
    f["wrench"][0] += 400.0
    parts.append(f)
    g = synth.make_states("C5", 16, start=4242)
    nrm = np.zeros((12, 16))
    for leg in range(4):
        v = np.stack([0.15 * np.sin(np.arange(16) + leg), 0.1 * np.cos(2 * np.arange(16) - leg), np.ones(16)])
        nrm[3 * leg:3 * leg + 3] = v / np.linalg.norm(v, axis=0)
    g["normals"] = nrm
    parts.append(g)
    st = cat(parts)
    out = {}
    for model in ("quadruped_model", "simpledog"):
        M = O.model_array(legmodel.load_model(model))
        r = O.solve_wrench_batch(M, st["q"], st["quat"], st["wrench"], st["mask"], mu=st["mu"], normals=st["normals"],
                                 solver=O.SOLVER_REF)
        gi = O.solve_wrench_batch(M, st["q"], st["quat"], st["wrench"], st["mask"], mu=st["mu"], normals=st["normals"],
                                  solver=O.SOLVER_GI)
        assert np.array_equal(r["grf"], gi["grf"]), "oracle port differs from the reference solver"
        # active bits come from the port's working set (the reference solver does not export it)
        out[model + "_grf"] = r["grf"]; out[model + "_tau"] = r["tau"]; out[model + "_net"] = r["netwrench"]
        out[model + "_flags"] = gi["flags"]
    np.savez_compressed(os.path.join(HERE, "golden_states.npz"), **st, **out)
    print("wrote golden_states.npz with", st["q"].shape[1], "states")

    # known-answer vectors of SURVEY.md Appendix C / D, re-derived here with the reference solver at full
    # precision; the survey's 9-digit values are kept beside them as the cross-check
    M = O.model_array(legmodel.load_model("quadruped_model"))
    q = [0, 0.7, -1.4, 0, -0.7, 1.4, 0, 0.7, -1.4, 0, -0.7, 1.4]
    kats = []
    for name, mask, ypr, mu, b, survey_x, survey_active in [
        ("KAT-A", 0xF, (0, 0, 0), 0.6, [30, -20, 499.8, 5, -8, 3],
         [6.634151753, -3.769007897, 131.385489914, 8.365261118, -3.769007796, 110.993807035,
          8.365261118, -6.230778897, 118.508270782, 6.634151751, -6.230778997, 138.899953661], []),
        ("KAT-B", 0b1010, (0.5, -0.15, 0.1), 0.6, [20, 10, 480, -4, 6, 2],
         [-0.936321891, 3.131035207, 240.938107573, -3.118967682, 0.027144342, 238.701032812], []),
        ("KAT-C", 0b1110, (0.3, 0.05, -0.08), 0.4, [150, -40, 499.8, 0, 0, 0],
         [57.659024940, -21.606702777, 165.466808781, 20.811735193, -3.071189428, 56.358298587,
          71.529326442, -15.317894413, 277.953062094], [3, 7]),
    ]:
        quat = synth.quat_from_ypr(*[np.array(float(v)) for v in ypr])
        a = O.assemble(M, q, quat, b, mask, mu=[mu] * 4)
        r = O.solve_qp_ref(a["G"], a["g0"], a["D"], a["d"])
        kats.append(dict(name=name, q=q, quat=[float(v) for v in quat], wrench=b, mask=mask, mu=mu,
                         x=[float(v) for v in r["x"]], survey_x=survey_x, survey_active_rows=survey_active))
    kin = [
        dict(model="quadruped_model", leg=0, q=[0.1, 0.7, -1.4], foot=[0.4269975175, 0.3508867522, -0.4553273500],
             jac=[[1.0134e-6, -0.4711427874, -0.2355721189], [0.4458273500, -1.722e-6, 0.0198079905],
                  [0.2758867522, 1.5578e-6, -0.1974270000]], gtau=[10.1182694508, 5.1856109217, -0.6334266156]),
        dict(model="quadruped_model", leg=1, q=[-0.05, -0.6, 1.3], foot=[0.4515067004, -0.2707759493, -0.5096846639],
             jac=[[-2.9554e-6, 0.4897748440, 0.2355720488], [-0.5001845738, -0.0012267490, -0.0099176856],
                  [0.1957757692, 0.0244768267, 0.1981702535]], gtau=[8.9683508561, -4.4802321819, 0.6371429926]),
        dict(model="simpledog", leg=0, q=[0.1, 0.7, -1.4], foot=[0.3997016821, 0.3122489366, -0.3703818090],
             gtau=[2.7771815140, 2.3775181213, -0.8394608620]),
        dict(model="simpledog", leg=1, q=[-0.05, -0.6, 1.3], foot=[0.4198935872, -0.2544251710, -0.4020168258],
             gtau=[2.0456820153, -1.9807382589, 0.8428315447]),
        dict(model="quadruped_model", leg=0, q=[0.0, 0.0, 0.0], foot=[0.427, 0.305, -0.6255], atol=5e-6),  # survey quotes 4 digits
        dict(model="quadruped_model", leg=1, q=[0.0, 0.0, 0.0], foot=[0.427, -0.2955, -0.6255], atol=5e-5),
    ]
    solver = dict(G=[[1, -1], [-1, 2]], g0=[-2, -6], D=[[-1, -1], [1, -2], [-2, -1]], d=[-2, -2, -3],
                  x=[0.66666666666666685, 1.3333333333333335], f=-8.2222222222222197,
                  source="qp_solver/src/main.cc:49-101 with the zero equality column dropped (SURVEY Appendix C)")
    with open(os.path.join(HERE, "kat_survey.json"), "w") as fjs:
        json.dump(dict(kats=kats, kinematics=kin, solver=solver), fjs, indent=1)
    print("wrote kat_survey.json")


if __name__ == "__main__":
    main()
