"""CPU tests: pin the oracle (golden vectors, the reference's own solver, KKT self-certification)."""
import numpy as np
import pytest

from conftest import rel_err
from quadruped_locomotion_b200 import synth

STATE_KEYS = ("q", "quat", "wrench", "mask", "mu", "normals")


def _solve(O, M, st, **kw):
    return O.solve_wrench_batch(M, st["q"], st["quat"], st["wrench"], st["mask"], mu=st["mu"], normals=st["normals"], **kw)


def test_kinematics_spot_values(oracle, models, kats):
    # two independent restatements of KDL FK / Jacobian / JntToGravity agree
    for k in kats["kinematics"]:
        foot, jac, gt = oracle.leg_kinematics(models[k["model"]], k["leg"], k["q"])
        np.testing.assert_allclose(foot, k["foot"], atol=k.get("atol", 2e-10))
        if "jac" in k:
            np.testing.assert_allclose(jac, np.array(k["jac"]), atol=2e-10)
        if "gtau" in k:
            np.testing.assert_allclose(gt, k["gtau"], atol=2e-9)


def test_jacobian_is_derivative_of_fk(oracle, models):
    rng = np.random.default_rng(1)
    for name, M in models.items():
        for leg in range(4):
            q = rng.uniform(-1.5, 1.5, 3)
            f0, J, _ = oracle.leg_kinematics(M, leg, q)
            for j in range(3):
                dq = np.zeros(3); dq[j] = 1e-6
                fp, _, _ = oracle.leg_kinematics(M, leg, q + dq)
                fm, _, _ = oracle.leg_kinematics(M, leg, q - dq)
                np.testing.assert_allclose((fp - fm) / 2e-6, J[:, j], atol=1e-8)


def test_gravity_torque_is_potential_gradient(oracle, models):
    # G(q) = d/dq of the potential energy -sum m g.c(q): checks the RNE restatement against FK of the COMs
    from quadruped_locomotion_b200 import legmodel
    g = np.array([0.3, -0.2, -9.8])
    for name, M in models.items():
        mdl = legmodel.load_model(name)
        for leg in range(4):
            q = np.array([0.2, -0.5, 0.9])
            _, _, G = oracle.leg_kinematics(M, leg, q, grav=g)

            def potential(qq):
                R = np.eye(3); p = np.zeros(3); U = 0.0
                L = mdl["legs"][leg]
                for k in range(4):
                    Rk = np.zeros(9)
                    import ctypes as C
                    oracle.lib().qo_rpy_to_rot(np.ascontiguousarray(L["joint_rpy"][k], dtype=np.float64).ctypes.data_as(C.POINTER(C.c_double)),
                                               Rk.ctypes.data_as(C.POINTER(C.c_double)))
                    p = p + R @ np.array(L["joint_xyz"][k]); R = R @ Rk.reshape(3, 3)
                    if k < 3:
                        c, s = np.cos(qq[k]), np.sin(qq[k])
                        R = R @ np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]])
                    U -= L["link_mass"][k] * g @ (p + R @ np.array(L["link_com"][k]))
                return U
            for j in range(3):
                dq = np.zeros(3); dq[j] = 1e-6
                num = (potential(q + dq) - potential(q - dq)) / 2e-6
                assert abs(num - G[j]) < 1e-6


def test_solver_known_answer(oracle, kats):
    s = kats["solver"]
    r = oracle.solve_qp_gi(s["G"], s["g0"], np.array(s["D"], float), s["d"])
    np.testing.assert_allclose(r["x"], s["x"], rtol=0, atol=1e-14)
    assert abs(r["f"] - s["f"]) < 1e-13
    assert list(r["active"]) == [True, True, False]
    ri = oracle.solve_qp_ipm(s["G"], s["g0"], np.array(s["D"], float), s["d"])
    np.testing.assert_allclose(ri["x"], s["x"], atol=1e-12)


def test_survey_known_answers(oracle, models, kats):
    M = models["quadruped_model"]
    for k in kats["kats"]:
        a = oracle.assemble(M, k["q"], k["quat"], k["wrench"], k["mask"], mu=[k["mu"]] * 4)
        r = oracle.solve_qp_gi(a["G"], a["g0"], a["D"], a["d"])
        # full-precision answer of the reference solver, and the survey's 9-digit transcription of it
        np.testing.assert_allclose(r["x"], k["x"], rtol=0, atol=1e-12)
        np.testing.assert_allclose(r["x"], k["survey_x"], rtol=0, atol=5e-8)
        assert list(np.nonzero(r["active"])[0]) == k["survey_active_rows"]


def test_port_matches_reference_solver_bit_exactly(oracle, models):
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    M = models["quadruped_model"]
    for cfg, B in (("C2", 4096), ("C3", 8192), ("C5", 8192)):
        st = synth.make_states(cfg, B)
        a = _solve(oracle, M, st, solver=oracle.SOLVER_GI)
        b = _solve(oracle, M, st, solver=oracle.SOLVER_REF)
        assert np.array_equal(a["grf"], b["grf"])
        assert np.array_equal(a["tau"], b["tau"])
        assert np.array_equal(a["flags"] & 0xFFFFFF, b["flags"] & 0xFFFFFF)


def test_golden_fixture(oracle, models, golden):
    st = {k: golden[k] for k in STATE_KEYS}
    for name in ("quadruped_model", "simpledog"):
        r = _solve(oracle, models[name], st, solver=oracle.SOLVER_GI)
        assert np.array_equal(r["grf"], golden[name + "_grf"])
        assert np.array_equal(r["tau"], golden[name + "_tau"])
        assert np.array_equal(r["flags"] & 0xFFFFFF, golden[name + "_flags"] & 0xFFFFFF)
        np.testing.assert_allclose(r["netwrench"], golden[name + "_net"], atol=1e-9)


def test_kkt_self_certification(oracle, models):
    """Every oracle solution satisfies stationarity, feasibility and complementarity to 1e-9."""
    M = models["quadruped_model"]
    st = synth.make_states("C5", 400, start=5000)
    for i in range(400):
        a = oracle.assemble(M, st["q"][:, i], st["quat"][:, i], st["wrench"][:, i], st["mask"][i], mu=st["mu"][:, i])
        r = oracle.solve_qp_gi(a["G"], a["g0"], a["D"], a["d"])
        x, u = r["x"], r["u"]
        gscale = max(1.0, np.abs(a["g0"]).max())
        assert np.abs(a["G"] @ x + a["g0"] - a["D"].T @ u).max() <= 1e-9 * gscale
        slack = a["D"] @ x - a["d"]
        assert slack.min() >= -1e-9 * max(1.0, np.abs(x).max())
        assert u.min() >= 0.0
        assert np.abs(slack * u).max() <= 1e-9 * gscale
        assert not (r["active"] & (u <= 0)).any()


def test_ipm_agrees_with_active_set(oracle, models):
    M = models["quadruped_model"]
    for cfg in ("C3", "C5"):
        st = synth.make_states(cfg, 20000)
        a = _solve(oracle, M, st, solver=oracle.SOLVER_GI)
        b = _solve(oracle, M, st, solver=oracle.SOLVER_IPM)
        assert rel_err(b["grf"], a["grf"]).max() < 1e-9
        assert ((a["flags"] ^ b["flags"]) & 0xFFFFFF).astype(bool).sum() == 0


def test_reference_double_solve_is_identity(oracle, models):
    """The reference solves twice per tick (CFD.cpp:367 then :120 with C = I, c = x1): same answer."""
    M = models["quadruped_model"]
    st = synth.make_states("C3", 2000)
    a = _solve(oracle, M, st, solver=oracle.SOLVER_GI, nsolves=1)
    b = _solve(oracle, M, st, solver=oracle.SOLVER_GI, nsolves=2)
    assert rel_err(b["grf"], a["grf"]).max() < 1e-10


def test_flags_and_edge_masks(oracle, models):
    M = models["quadruped_model"]
    st = synth.make_states("C3", 16)
    st["mask"] = np.arange(16, dtype=np.uint8)
    r = _solve(oracle, M, st, solver=oracle.SOLVER_GI)
    assert np.array_equal(r["flags"] & 0xF, np.arange(16))
    assert ((r["flags"][0] >> 24) & 7) == 1 and not r["grf"][:, 0].any() and not r["tau"][:, 0].any()
    for i in range(1, 16):
        for leg in range(4):
            stance = (i >> leg) & 1
            blk = r["grf"][3 * leg:3 * leg + 3, i]
            assert bool(np.abs(blk).sum() > 0) == bool(stance)
            if not stance:
                assert not r["tau"][3 * leg:3 * leg + 3, i].any()
                assert ((r["flags"][i] >> (4 + 5 * leg)) & 31) == 0
    bad = synth.make_states("C3", 4)
    bad["q"][5, 1] = np.nan
    bad["wrench"][2, 2] = np.inf
    r = _solve(oracle, M, bad, solver=oracle.SOLVER_GI)
    assert list((r["flags"] >> 24) & 7) == [0, 4, 4, 0]
    assert not r["grf"][:, 1].any() and not r["grf"][:, 2].any()


def test_vmc_gravity_only(oracle):
    """zero pose/twist error -> the wrench is pure gravity compensation (VMC.cpp:162-188)."""
    quat = synth.quat_from_ypr(np.array(0.4), np.array(-0.1), np.array(0.2))
    pose = np.concatenate([[0.1, -0.2, 0.45], quat])
    w = oracle.vmc_wrench(pose, np.zeros(6), pose, np.zeros(6))
    R = synth.rot_from_quat(quat)
    np.testing.assert_allclose(w[:3], R.T @ np.array([0, 0, 51 * 9.8]), atol=1e-10)
    fl = R.T @ np.array([0, 0, 6 * 9.8])
    T = sum(np.cross(np.array(r), fl) for r in ([0.42, 0.075, 0], [0.42, -0.075, 0], [-0.42, -0.075, 0], [-0.42, 0.075, 0]))
    np.testing.assert_allclose(w[3:], T, atol=1e-10)
    # a pure yaw error of +0.1 rad gives a positive yaw torque kp_yaw * 0.1
    tq = synth.quat_from_ypr(np.array(0.5), np.array(-0.1), np.array(0.2))
    w2 = oracle.vmc_wrench(pose, np.zeros(6), np.concatenate([pose[:3], tq]), np.zeros(6))
    assert abs(np.linalg.norm(w2[3:] - w[3:]) - 0.1 * 4000) < 0.1 * 10000  # rotated into the base frame
    assert np.linalg.norm(w2[3:] - w[3:]) > 100


def test_negative_friction_is_reported_infeasible(oracle, models):
    """mu < 0 on a stance leg with F_min > 0: no force satisfies mu n.f >= |t.f| and n.f >= F_min.  The reference's
    solver returns +inf (QuadProg++.cc:340-344); the oracle reports state status 5 and zero outputs."""
    from quadruped_locomotion_b200 import synth
    st = synth.make_states("C3", 64)
    st["mask"][:] = 0xF
    st["mu"][2, ::2] = -0.3
    r = oracle.solve_wrench_batch(models["quadruped_model"], st["q"], st["quat"], st["wrench"], st["mask"], mu=st["mu"])
    status = (r["flags"] >> 24) & 7
    assert (status[::2] == 5).all() and (status[1::2] == 0).all()
    assert not r["grf"][:, ::2].any() and not r["tau"][:, ::2].any()


def test_nonpositive_minimal_force_is_equivalent_to_zero(oracle, models):
    """F_min <= 0 with mu > 0: the friction rows already imply n.f >= 0, so the optimum equals the one for F_min = 0
    (what the GPU path solves after clamping F_min)."""
    from quadruped_locomotion_b200 import synth
    st = synth.make_states("C5", 256)
    p0 = oracle.default_params(); p0.fmin = 0.0
    pm = oracle.default_params(); pm.fmin = -25.0
    a = oracle.solve_wrench_batch(models["quadruped_model"], st["q"], st["quat"], st["wrench"], st["mask"], mu=st["mu"], params=p0)
    b = oracle.solve_wrench_batch(models["quadruped_model"], st["q"], st["quat"], st["wrench"], st["mask"], mu=st["mu"], params=pm)
    assert np.abs(a["grf"] - b["grf"]).max() <= 1e-8 * np.abs(a["grf"]).max()
