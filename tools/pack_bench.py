#!/usr/bin/env python
"""Bandwidth of the RobotState record packer (HBM-bound byte movement): GB/s against the measured HBM peak."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from quadruped_locomotion_b200 import capi  # noqa: E402


def main():
    B = 1 << 22
    dev = torch.device("cuda:0")
    s = capi.Solver("quadruped_model")
    rec = torch.randint(0, 255, (B * capi.RECORD_DTYPE.itemsize,), dtype=torch.uint8, device=dev)
    q = torch.empty((12, B), dtype=torch.float64, device=dev); pose = torch.empty((7, B), dtype=torch.float64, device=dev)
    twist = torch.empty((6, B), dtype=torch.float64, device=dev); nrm = torch.empty((12, B), dtype=torch.float64, device=dev)
    mask = torch.empty(B, dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(3):
        s.pack_robot_states(rec, q, pose, twist, mask, nrm, stream=st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record()
    for _ in range(reps):
        s.pack_robot_states(rec, q, pose, twist, mask, nrm, stream=st)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    bytes_per = capi.RECORD_DTYPE.itemsize + 37 * 8 + 1
    peak = 6554.2
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", peak)
    except Exception:
        pass
    gbs = B * bytes_per / (ms * 1e-3) * 1e-9
    # swing-leg torques: 2^20 states x 4 legs, FP64 arithmetic (limb inverse dynamics + Cartesian PD)
    Bs = 1 << 20
    s.set_limb_dynamics("quadruped_model")
    prm = s.default_swing_params()
    for c in range(3):
        prm.kp[c] = 300.0; prm.kd[c] = 10.0
    arrs = [torch.randn((12, Bs), dtype=torch.float64, device=dev) for _ in range(5)]
    tau = torch.empty((12, Bs), dtype=torch.float64, device=dev)
    for _ in range(3):
        s.swing_leg_torques(arrs[0], arrs[1], arrs[2], arrs[3], arrs[4], prm, tau, stream=st)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        s.swing_leg_torques(arrs[0], arrs[1], arrs[2], arrs[3], arrs[4], prm, tau, stream=st)
    e1.record()
    torch.cuda.synchronize()
    ms_s = e0.elapsed_time(e1) / reps
    print(json.dumps({"kernel": "qlb_swing_kernel", "states": Bs, "ms": ms_s, "states_per_s": Bs / (ms_s * 1e-3),
                      "bytes_per_state": 6 * 12 * 8, "achieved_gbs": Bs * 6 * 12 * 8 / (ms_s * 1e-3) * 1e-9, "peak_gbs": peak}))
    print(json.dumps({"kernel": "qlb_pack_kernel", "records": B, "ms": ms, "bytes_per_record": bytes_per,
                      "achieved_gbs": gbs, "peak_gbs": peak, "frac": gbs / peak, "records_per_s": B / (ms * 1e-3)}))


if __name__ == "__main__":
    main()
