"""swing-leg joint torques (MyRobotSolver::update, single_leg_test/lib/model_test_header.cpp:412-502)
= limb inverse dynamics + Cartesian PD.  The dynamics library the reference calls is not vendored, so parity for
this row is UNPINNED against the reference; what pins the restatement instead: two independent formulations
(spatial-vector recursion vs Lagrange equations by finite differences) agree, and the CUDA kernel (a third
formulation: base-frame sums) agrees with the first to rounding."""
import numpy as np
import pytest
import torch

from quadruped_locomotion_b200 import capi, legmodel


def _motion(B, seed=3):
    rng = np.random.default_rng(seed)
    s = np.array([1, -1, 1, -1])[:, None]
    q = np.zeros((12, B))
    q[0::3] = rng.uniform(-0.25, 0.25, (4, B))
    q[1::3] = s * (0.7 + rng.uniform(-0.3, 0.3, (4, B)))
    q[2::3] = -s * (1.4 + rng.uniform(-0.4, 0.4, (4, B)))
    return q, rng.normal(0, 2.0, (12, B)), rng.normal(0, 20.0, (12, B)), rng


def test_limb_tables_merge_the_foot_link():
    m = legmodel.load_model("quadruped_model")
    for leg, limb in zip(m["legs"], m["limb_dynamics"]):
        assert limb["body_mass"][0] == pytest.approx(leg["link_mass"][0], rel=1e-5)
        assert limb["body_mass"][2] == pytest.approx(leg["link_mass"][2] + leg["link_mass"][3], rel=1e-4)
        assert limb["joint_xyz"] == leg["joint_xyz"][:3] and limb["joint_rpy"] == leg["joint_rpy"][:3]
        I = limb["body_inertia"][2]
        assert np.all(np.linalg.eigvalsh(np.array([[I[0], I[1], I[2]], [I[1], I[3], I[4]], [I[2], I[4], I[5]]])) > 0)


def test_recursion_agrees_with_the_lagrange_equations(oracle):
    m = legmodel.load_model("quadruped_model")
    q, qd, qdd, _ = _motion(6)
    for i in range(6):
        for leg in range(4):
            sl = slice(3 * leg, 3 * leg + 3)
            a = oracle.limb_inverse_dynamics(m["limb_dynamics"][leg], q[sl, i], qd[sl, i], qdd[sl, i])
            b = oracle.limb_lagrangian_torques(m["limb_dynamics"][leg], q[sl, i], qd[sl, i], qdd[sl, i])
            assert np.abs(a - b).max() <= 2e-6 * max(1.0, np.abs(a).max())
    # statics: with zero velocity and acceleration the torques are the gravity torques of the kinematics oracle
    # (same masses; the limb tables merge the foot link) for the same gravity vector
    M = oracle.model_array(m)
    for leg in range(4):
        g = (0.3, -9.0, -4.0)
        tau = oracle.limb_inverse_dynamics(m["limb_dynamics"][leg], q[3 * leg:3 * leg + 3, 0], np.zeros(3), np.zeros(3), g)
        _, _, gt = oracle.leg_kinematics(M, leg, q[3 * leg:3 * leg + 3, 0], grav=g)
        assert np.abs(tau - gt).max() <= 2e-5   # URDF masses differ in the sixth digit between the files


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["quadruped_model", "simpledog"])
def test_swing_kernel_against_the_restatement(qlb_built, oracle, name):
    B = 300
    m = legmodel.load_model(name)
    M = oracle.model_array(m)
    q, qd, qdd, rng = _motion(B)
    pt = rng.normal(0, 0.3, (12, B)); vt = rng.normal(0, 0.5, (12, B))
    dev = torch.device("cuda:0")
    s = capi.Solver(name)
    s.set_limb_dynamics(m)
    prm = s.default_swing_params()
    assert list(prm.gravity) == [0.0, -9.81, 0.0] and prm.acceleration_scale == 0.5
    for c, (kp, kd) in enumerate(((300.0, 10.0), (200.0, 12.0), (400.0, 8.0))):
        prm.kp[c] = kp; prm.kd[c] = kd
    d = [torch.from_numpy(np.ascontiguousarray(x)).to(dev) for x in (q, qd, qdd, pt, vt)]
    tau = torch.zeros((12, B), dtype=torch.float64, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    s.swing_leg_torques(d[0], d[1], d[2], d[3], d[4], prm, tau, stream=stream)
    torch.cuda.synchronize()
    ref = oracle.swing_leg_torques(m, M, q, qd, qdd, pt, vt, kp=list(prm.kp), kd=list(prm.kd))
    assert np.abs(tau.cpu().numpy() - ref).max() <= 1e-10 * max(1.0, np.abs(ref).max())
    # host entry point, a batch of one (a controller tick) and a ragged batch
    for n in (1, 37):
        th = np.zeros((12, n))
        c = lambda x: np.ascontiguousarray(x[:, :n])  # noqa: E731
        s.swing_leg_torques_host(c(q), c(qd), c(qdd), c(pt), c(vt), prm, th)
        assert np.abs(th - ref[:, :n]).max() <= 1e-10 * max(1.0, np.abs(ref).max())
    # inverse dynamics only
    s.swing_leg_torques(d[0], d[1], d[2], None, None, prm, tau, stream=stream)
    torch.cuda.synchronize()
    ref0 = oracle.swing_leg_torques(m, M, q, qd, qdd)
    assert np.abs(tau.cpu().numpy() - ref0).max() <= 1e-10 * max(1.0, np.abs(ref0).max())
    s.close()


@pytest.mark.gpu
def test_swing_needs_the_limb_tables(qlb_built):
    s = capi.Solver("quadruped_model")
    t = torch.zeros((12, 8), dtype=torch.float64, device="cuda:0")
    with pytest.raises(RuntimeError):
        s.swing_leg_torques(t, t, t, None, None, s.default_swing_params(), t)
    s.close()


def test_committed_model_header_matches_the_tables(tmp_path):
    """include/qlb_models.h is generated from models/*.json (tools/make_models.py): keep them in step."""
    import os
    models = {"QLB_MODEL_" + n.upper(): legmodel.load_model(n) for n in ("quadruped_model", "simpledog")}
    out = tmp_path / "qlb_models.h"
    legmodel.emit_c_header(models, str(out))
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    assert out.read_text() == open(os.path.join(root, "include", "qlb_models.h")).read()
    flat = legmodel.limb_dynamics_to_flat(models["QLB_MODEL_QUADRUPED_MODEL"])
    assert flat.shape == (4, 48) and flat[0, 18] == models["QLB_MODEL_QUADRUPED_MODEL"]["limb_dynamics"][0]["body_mass"][0]
