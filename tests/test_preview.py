"""the RobotState record packer and the batched preview of a planned motion
(free_gait_msgs/RobotState.msg, ros_balance_controller.cpp:761-811; StateBatchComputer.cpp:64-77,
BatchExecutor.cpp:69-83).  CPU part: the numpy restatement on a hand-written record.  GPU part: kernels
through the C ABI against that restatement, bit-exact for the byte movement."""
import numpy as np
import pytest
import torch

from conftest import rel_err
from quadruped_locomotion_b200 import capi, synth


def _records(B, seed=11, cfg="C3"):
    st = synth.make_states(cfg, B, start=seed)
    rng = np.random.default_rng(seed)
    rec = np.zeros(B, dtype=capi.RECORD_DTYPE)
    rec["base_position"] = rng.normal(0, 0.2, (B, 3)) + np.array([0, 0, 0.45])
    rec["base_orientation_xyzw"] = np.stack([st["quat"][1], st["quat"][2], st["quat"][3], st["quat"][0]], axis=1)
    rec["base_linear_velocity"] = rng.normal(0, 0.1, (B, 3))
    rec["base_angular_velocity"] = rng.normal(0, 0.1, (B, 3))
    rec["joint_position"] = st["q"].T
    rec["surface_normal"] = st["normals"].T
    for leg in range(4):
        rec["support_leg"][:, leg] = ((st["mask"] >> leg) & 1) * (1 + leg)   # any non-zero byte means "support"
    rec["reserved"] = 0xAB
    return rec, st


def test_record_layout_and_restatement(oracle):
    assert capi.RECORD_DTYPE.itemsize == 304 and capi.RECORD_DTYPE.fields["support_leg"][1] == 296
    rec = np.zeros(2, dtype=capi.RECORD_DTYPE)
    rec["base_position"][1] = (1, 2, 3)
    rec["base_orientation_xyzw"][1] = (0.1, 0.2, 0.3, 0.9)
    rec["joint_position"][1] = np.arange(12)
    rec["support_leg"][1] = (1, 0, 0, 7)
    o = oracle.pack_robot_states(rec)
    assert o["pose"][:, 1].tolist() == [1, 2, 3, 0.9, 0.1, 0.2, 0.3]      # quaternion stored w first
    assert o["q"][:, 1].tolist() == list(range(12)) and o["mask"].tolist() == [0, 0b1001]


@pytest.mark.gpu
@pytest.mark.parametrize("B", [1, 127, 128, 129, 100003])
def test_packer_is_bit_exact(qlb_built, oracle, B):
    rec, _ = _records(B)
    ref = oracle.pack_robot_states(rec)
    dev = torch.device("cuda:0")
    s = capi.Solver("quadruped_model")
    d_rec = torch.from_numpy(rec.view(np.uint8).reshape(-1)).to(dev)
    q = torch.zeros((12, B), dtype=torch.float64, device=dev); pose = torch.zeros((7, B), dtype=torch.float64, device=dev)
    twist = torch.zeros((6, B), dtype=torch.float64, device=dev); mask = torch.zeros(B, dtype=torch.uint8, device=dev)
    nrm = torch.zeros((12, B), dtype=torch.float64, device=dev)
    s.pack_robot_states(d_rec, q, pose, twist, mask, nrm, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    for name, t in (("q", q), ("pose", pose), ("twist", twist), ("mask", mask), ("normals", nrm)):
        assert np.array_equal(t.cpu().numpy(), ref[name]), name
    s.close()


@pytest.mark.gpu
def test_preview_of_a_planned_motion(qlb_built, oracle, models):
    """Every sample of a plan in one batch: records -> packer -> feet in world frame + force distribution with
    the planned state as both feedback and target (zero tracking error: the wrench is the gravity compensation)."""
    B = 4096
    rec, st = _records(B, seed=5)
    M = models["quadruped_model"]
    dev = torch.device("cuda:0")
    s = capi.Solver("quadruped_model")
    stream = torch.cuda.current_stream().cuda_stream
    d_rec = torch.from_numpy(rec.view(np.uint8).reshape(-1)).to(dev)
    f64 = lambda *shape: torch.zeros(shape, dtype=torch.float64, device=dev)  # noqa: E731
    q, pose, twist, nrm, mask = f64(12, B), f64(7, B), f64(6, B), f64(12, B), torch.zeros(B, dtype=torch.uint8, device=dev)
    feet, grf, tau, net, wout = f64(12, B), f64(12, B), f64(12, B), f64(6, B), f64(6, B)
    flags = torch.zeros(B, dtype=torch.int32, device=dev)
    s.pack_robot_states(d_rec, q, pose, twist, mask, nrm, stream=stream)
    s.feet_in_world(q, pose, feet, stream=stream)
    s.solve_state(q, pose, twist, pose, twist, mask, None, nrm, grf, tau, flags, net, wout, stream=stream)
    torch.cuda.synchronize()
    p = oracle.pack_robot_states(rec)
    n = 256   # the per-state Python loops of the restatement are slow
    fw = oracle.feet_in_world(M, p["q"][:, :n], p["pose"][:, :n])
    assert np.abs(feet.cpu().numpy()[:, :n] - fw).max() <= 1e-12
    wref = np.stack([oracle.vmc_wrench(p["pose"][:, i], p["twist"][:, i], p["pose"][:, i], p["twist"][:, i]) for i in range(B)], axis=1)
    assert rel_err(wout.cpu().numpy(), wref).max() <= 1e-11
    ref = oracle.solve_wrench_batch(M, p["q"], p["pose"][3:], wref, p["mask"], normals=p["normals"], want_margin=True)
    assert rel_err(grf.cpu().numpy(), ref["grf"]).max() <= 1e-9
    assert rel_err(tau.cpu().numpy(), ref["tau"]).max() <= 1e-9
    fl = flags.cpu().numpy().view(np.uint32)
    mism = ((fl ^ ref["flags"]) & capi.FLAG_PARITY_MASK) != 0
    assert not (mism & (ref["margin"] > 1e-6)).any()
    s.close()


@pytest.mark.gpu
@pytest.mark.parametrize("B", [1, 300, 150001])
def test_preview_plan_one_call(qlb_built, oracle, models, B):
    """qlb_preview_plan_host: the composition of the test above as ONE call on array-of-structs records, plus friction
    margins (several pipeline chunks at the largest size)."""
    rec, st = _records(B, seed=9, cfg="C5")
    s = capi.Solver("quadruped_model")
    mu = (0.5, 0.6, 0.7, 0.8)
    out = s.preview_plan_host(rec, mu=mu)
    M = models["quadruped_model"]
    p = oracle.pack_robot_states(rec)
    n = min(B, 300)
    fw = oracle.feet_in_world(M, p["q"][:, :n], p["pose"][:, :n])
    assert np.abs(out["feet_world"][:n].T - fw).max() <= 1e-12
    wref = np.stack([oracle.vmc_wrench(p["pose"][:, i], p["twist"][:, i], p["pose"][:, i], p["twist"][:, i]) for i in range(n)], axis=1)
    assert rel_err(out["wrench"][:n].T, wref).max() <= 1e-11
    mu_a = np.tile(np.array(mu)[:, None], (1, n))
    ref = oracle.solve_wrench_batch(M, p["q"][:, :n], p["pose"][3:, :n], wref, p["mask"][:n], mu=mu_a, normals=p["normals"][:, :n])
    assert rel_err(out["grf"][:n].T, ref["grf"]).max() <= 1e-9 and rel_err(out["tau"][:n].T, ref["tau"]).max() <= 1e-9
    m, mn = oracle.friction_margins(ref["grf"], p["pose"][3:, :n], p["mask"][:n], mu_a, p["normals"][:, :n])
    np.testing.assert_allclose(out["friction_margin"][:n], m, atol=1e-8)
    np.testing.assert_allclose(out["min_normal_slack"][:n], mn, atol=1e-7)
    assert (out["reserved"] == 0).all() and (((out["flags"] >> 24) & 7) <= 1).all()
    s.close()
