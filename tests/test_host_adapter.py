"""The C++ adapter classes (quadruped_locomotion_b200/host/qlb_adapter.hpp) mirror the reference's
ContactForceDistribution / VirtualModelController / State interface over the C ABI.  CPU: they compile
and link.  GPU: one tick (batch of 1) through them reproduces the CPU oracle."""
import json
import subprocess

import numpy as np
import pytest

from quadruped_locomotion_b200 import build, synth


def test_adapter_compiles_and_links(qlb_built):
    demo = build.build_host_demo()
    out = subprocess.run(["ldd", demo], capture_output=True, text=True).stdout
    assert "libqlb.so" in out and "not found" not in out


@pytest.mark.gpu
@pytest.mark.parametrize("mask,ypr,mu,wrench", [
    (0xF, (0.0, 0.0, 0.0), 0.6, [30, -20, 499.8, 5, -8, 3]),          # KAT-A
    (0b1110, (0.3, 0.05, -0.08), 0.4, [150, -40, 499.8, 0, 0, 0]),    # KAT-C (two friction rows active)
    (0b0101, (-1.2, 0.1, 0.05), 0.7, [10, 40, 520, 8, -3, 1]),
])
def test_one_tick_matches_oracle(qlb_built, oracle, models, mask, ypr, mu, wrench):
    demo = build.build_host_demo()
    q = [0, 0.7, -1.4, 0, -0.7, 1.4, 0, 0.7, -1.4, 0, -0.7, 1.4]
    quat = synth.quat_from_ypr(*[np.array(float(v)) for v in ypr])
    pos, lv, av = [0.01, -0.02, 0.44], [0.05, 0.0, -0.01], [0.0, 0.02, 0.01]
    tquat = synth.quat_from_ypr(np.array(ypr[0] + 0.01), np.array(0.0), np.array(0.0))
    tpos, tlv, tav = [0.0, 0.0, 0.45], [0.1, 0.0, 0.0], [0.0, 0.0, 0.05]
    vals = q + list(quat) + pos + lv + av + list(tquat) + tpos + tlv + tav + [mask, mu] + wrench
    r = subprocess.run([demo], input=" ".join(repr(float(v)) for v in vals), capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    out = json.loads(r.stdout)
    M = models["quadruped_model"]

    def ref(w):
        st = dict(q=np.array(q, float)[:, None], quat=np.array(quat, float)[:, None], wrench=np.array(w, float)[:, None],
                  mask=np.array([mask], np.uint8), mu=np.full((4, 1), mu))
        return oracle.solve_wrench_batch(M, st["q"], st["quat"], st["wrench"], st["mask"], mu=st["mu"])

    a = ref(wrench)
    assert out["cfd_ok"]
    # desiredContactForce_ = -x (ContactForceDistribution.cpp:502-503)
    np.testing.assert_allclose(out["cfd_contact_force"], -a["grf"][:, 0], atol=1e-8)
    np.testing.assert_allclose(out["cfd_efforts"], a["tau"][:, 0], atol=1e-8)
    np.testing.assert_allclose(out["cfd_net"], a["netwrench"][:, 0], atol=1e-8)
    assert (out["cfd_flags"] & 0xFFFFFF) == (int(a["flags"][0]) & 0xFFFFFF)
    w = oracle.vmc_wrench(np.array(pos + list(quat)), np.array(lv + av), np.array(tpos + list(tquat)), np.array(tlv + tav))
    assert out["vmc_ok"]
    np.testing.assert_allclose(out["vmc_wrench"], w, rtol=1e-12, atol=1e-9)
    b = ref(w)
    np.testing.assert_allclose(out["vmc_efforts"], b["tau"][:, 0], rtol=1e-9, atol=1e-7)


@pytest.mark.gpu
def test_all_twelve_torques_of_a_tick(qlb_built, oracle, models):
    """Stance legs from VirtualModelController::compute, swing legs from MyRobotSolver::update, merged in State."""
    from quadruped_locomotion_b200 import legmodel
    demo = build.build_host_demo(which="swing_demo")
    rng = np.random.default_rng(9)
    q = np.array([0.05, 0.7, -1.4, -0.03, -0.75, 1.45, 0.02, 0.65, -1.3, 0.0, -0.7, 1.4])
    qd = rng.normal(0, 1.5, 12); qdd = rng.normal(0, 15.0, 12)
    mask = 0b1010   # RF and LH in stance, LF and RH swinging
    kp, kd = [300.0, 250.0, 400.0], [10.0, 12.0, 8.0]
    pt = rng.normal(0, 0.3, 12); vt = rng.normal(0, 0.4, 12)
    vals = list(q) + list(qd) + list(qdd) + [mask] + kp + kd + list(pt) + list(vt)
    r = subprocess.run([demo], input=" ".join(repr(float(v)) for v in vals), capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    out = json.loads(r.stdout)
    m = legmodel.load_model("quadruped_model")
    M = models["quadruped_model"]
    # the swing rows only use their own leg's velocities: zero the others like the adapter does leg by leg
    ref_swing = oracle.swing_leg_torques(m, M, q[:, None], qd[:, None], qdd[:, None], pt[:, None], vt[:, None], kp=kp, kd=kd)[:, 0]
    pose = np.array([0, 0, 0.45, 1, 0, 0, 0.0]); tw = np.zeros(6)
    w = oracle.vmc_wrench(pose, tw, pose, tw)
    ref = oracle.solve_wrench_batch(M, q[:, None], pose[3:, None], w[:, None], np.array([mask], np.uint8))
    want = ref["tau"][:, 0].copy()
    assert [s[0] for s in out["swing"]] == [0, 2]
    for leg, *tau in out["swing"]:
        np.testing.assert_allclose(tau, ref_swing[3 * leg:3 * leg + 3], rtol=0, atol=1e-9)
        want[3 * leg:3 * leg + 3] = ref_swing[3 * leg:3 * leg + 3]
    np.testing.assert_allclose(out["efforts"], want, rtol=0, atol=1e-8)


@pytest.mark.gpu
def test_state_batch_preview_of_a_plan(qlb_built):
    """StateBatch / StateBatchComputer mirror (free_gait_core/src/executor/StateBatchComputer.cpp:20-132): a 4 s plan
    sampled at 10 ms, previewed in one GPU call."""
    demo = build.build_host_demo(which="batch_demo")
    r = subprocess.run([demo], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    out = json.loads(r.stdout)
    assert out["samples"] == 400 and out["stances"] == 8       # the first stance + seven support changes
    np.testing.assert_allclose(out["fz"], [499.8] * 3, rtol=1e-3)   # the gravity compensation is distributed (the force
    # regulariser of the QP keeps the total a fraction of a newton below the commanded weight)
    assert 0.0 < out["min_margin"] <= 1.0
    # LF foot: under the hip at the start, carried 0.2 m forward by the base at the end
    assert abs(out["lf_last"][0] - out["lf_first"][0] - 0.05 * 3.99) < 1e-9 and abs(out["lf_first"][2]) < 0.05


@pytest.mark.gpu
def test_stats_allreduce_through_the_c_abi(qlb_built):
    """qlb_stats_allreduce over an NCCL communicator from a C++ host: one thread and one context per visible GPU
    (one rank on a single-GPU box), every rank solves its own slice of the generated batch."""
    demo = build.build_host_demo(which="nccl_demo")
    r = subprocess.run([demo, "20000"], capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stdout + r.stderr
    f = r.stdout.strip().splitlines()[-1].split()      # (NCCL may print its version line first)
    assert f[0] == "gpus" and f[-1] == "AGREE" and float(f[3]) == 20000.0 * int(f[1]) and float(f[5]) == float(f[3])


@pytest.mark.gpu
def test_one_context_shared_by_host_threads(qlb_built):
    """Four host threads call qlb_solve_wrench_host on ONE context at the same time (include/qlb.h: every entry point holds
    the context's lock while it enqueues): the bits of a serial run."""
    demo = build.build_host_demo(which="threads_demo")
    r = subprocess.run([demo, "4", "20000"], capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stdout + r.stderr
    f = r.stdout.strip().splitlines()[-1].split()
    assert f[-1] == "SAME" and int(f[3]) == 80000 and int(f[5]) == 80000
