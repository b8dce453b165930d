"""Batched contact state machine, friction margins and the velocity-queue acceleration estimate (SURVEY 8f rows 2
and 4) against their numpy restatements of the reference."""
import numpy as np
import pytest

from quadruped_locomotion_b200 import capi, synth


def test_fsm_restatement_on_the_reference_cases(oracle):
    """CPU: the transitions spelled out in ros_balance_controller.cpp:1098-1137."""
    O = oracle
    cases = [  # desired stance, footstep, contact, phase, previous -> expected state, support
        (0, 1, 0, 0.9, O.LIMB_STANCE_NORMAL, O.LIMB_SWING_NORMAL, 0),
        (0, 1, 1, 0.6, O.LIMB_SWING_NORMAL, O.LIMB_SWING_EARLY_TOUCHDOWN, 1),
        (0, 1, 1, 0.3, O.LIMB_SWING_NORMAL, O.LIMB_SWING_BUMPED_INTO_OBSTACLE, 0),
        (0, 1, 1, 0.1, O.LIMB_SWING_NORMAL, O.LIMB_SWING_NORMAL, 0),
        (0, 0, 1, 0.9, O.LIMB_SWING_NORMAL, O.LIMB_SWING_NORMAL, 0),          # not a footstep: contact ignored
        (1, 0, 0, 0.9, O.LIMB_SWING_NORMAL, O.LIMB_STANCE_NORMAL, 1),
        (1, 1, 1, 0.3, O.LIMB_SWING_NORMAL, O.LIMB_STANCE_NORMAL, 1),
        (1, 1, 0, 0.05, O.LIMB_SWING_NORMAL, O.LIMB_SWING_LATELY_TOUCHDOWN, 0),
        (1, 1, 0, 0.3, O.LIMB_SWING_LATELY_TOUCHDOWN, O.LIMB_SWING_LATELY_TOUCHDOWN, 0),   # state kept
        (1, 1, 0, 0.3, O.LIMB_STANCE_NORMAL, O.LIMB_STANCE_NORMAL, 1),                     # state kept
        (1, 1, 0, 0.7, O.LIMB_STANCE_NORMAL, O.LIMB_STANCE_LOST_CONTACT, 0),
    ]
    for d, f, c, ph, prev, want, sup in cases:
        st, mask = O.contact_fsm(np.array([d], np.uint8), np.array([f], np.uint8), np.array([c], np.uint8),
                                 np.full((4, 1), ph), np.full((4, 1), prev, np.uint8))
        assert st[0, 0] == want and (mask[0] & 1) == sup


@pytest.fixture(scope="module")
def solver(qlb_built):
    import torch
    if not torch.cuda.is_available():
        pytest.fail("GPU test selected but no CUDA device")
    s = capi.Solver("quadruped_model")
    yield s
    s.close()


@pytest.mark.gpu
def test_fsm_kernel_matches_restatement(solver, oracle):
    import torch
    rng = np.random.default_rng(3)
    B = 20011
    desired = rng.integers(0, 16, B).astype(np.uint8); footstep = rng.integers(0, 16, B).astype(np.uint8)
    contact = rng.integers(0, 16, B).astype(np.uint8)
    phase = rng.uniform(0.0, 1.0, (4, B)); prev = rng.integers(0, 9, (4, B)).astype(np.uint8)
    phase[:, :50] = np.array([0.1, 0.2, 0.5, 0.5])[:, None]          # the thresholds themselves
    want_state, want_mask = oracle.contact_fsm(desired, footstep, contact, phase, prev)
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()  # noqa: E731
    st = d(prev); mask = torch.zeros(B, dtype=torch.uint8, device="cuda")
    solver.contact_fsm(d(desired), d(footstep), d(contact), d(phase), st, mask, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert np.array_equal(st.cpu().numpy(), want_state) and np.array_equal(mask.cpu().numpy(), want_mask)
    # footstep NULL = every leg supervised; the mask feeds the solver directly
    st2 = d(prev)
    solver.contact_fsm(d(desired), None, d(contact), d(phase), st2, mask)
    torch.cuda.synchronize()
    w2, m2 = oracle.contact_fsm(desired, None, contact, phase, prev)
    assert np.array_equal(st2.cpu().numpy(), w2) and np.array_equal(mask.cpu().numpy(), m2)


@pytest.mark.gpu
def test_friction_margins_of_a_solved_batch(solver, oracle):
    import torch
    B = 3000
    st = synth.make_states("C5", B, start=77)
    out = solver.solve_wrench_numpy(st)
    want, wantn = oracle.friction_margins(out["grf"], st["quat"], st["mask"], st["mu"], st["normals"])
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()  # noqa: E731
    margin = torch.zeros(B, dtype=torch.float64, device="cuda"); minn = torch.zeros_like(margin)
    solver.friction_margins(d(out["grf"]), d(st["quat"]), d(st["mask"]), d(st["mu"]), d(st["normals"]), margin, minn)
    torch.cuda.synchronize()
    np.testing.assert_allclose(margin.cpu().numpy(), want, atol=1e-9)
    np.testing.assert_allclose(minn.cpu().numpy(), wantn, atol=1e-8)
    # a solved state never violates its pyramid; active friction rows show up as zero margin
    assert margin.min().item() > -1e-9
    act = ((out["flags"] >> 4) & 0xFFFFF)
    fr = np.zeros(B, bool)
    for leg in range(4):
        fr |= ((act >> (5 * leg + 1)) & 0xF) != 0
    assert np.abs(want[fr]).max() < 1e-8 and (want[~fr & (st["mask"] != 0)] > 1e-9).all()


@pytest.mark.gpu
def test_swing_torques_from_velocity_queue(solver):
    import torch
    solver.set_limb_dynamics("quadruped_model")
    rng = np.random.default_rng(5)
    B = 777
    q = rng.uniform(-1.0, 1.0, (12, B)); qd_back = rng.normal(size=(12, B)); qd_front = rng.normal(size=(12, B))
    period = 0.0025
    qdd = (qd_back - qd_front) / (10.0 * period)        # model_test_header.cpp:421,428
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()  # noqa: E731
    prm = solver.default_swing_params()
    t1 = torch.zeros((12, B), dtype=torch.float64, device="cuda"); t2 = torch.zeros_like(t1)
    solver.swing_leg_torques(d(q), d(qd_back), d(qdd), None, None, prm, t1)
    solver.swing_leg_torques_from_queue(d(q), d(qd_back), d(qd_front), period, None, None, prm, t2)
    torch.cuda.synchronize()
    np.testing.assert_allclose(t2.cpu().numpy(), t1.cpu().numpy(), rtol=1e-12, atol=1e-9)
