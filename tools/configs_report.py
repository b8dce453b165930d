#!/usr/bin/env python
"""One line per BASELINE.json config (C1..C5) on ONE GPU: device-resident throughput of the fused pipeline and
accuracy against the CPU oracle on a bounded sample.  C4 (FP32, 2/4/8 GPUs) and C5 (2^24 samples on 8 GPUs) are
run at their per-GPU share (2^20 and 2^21 states).  Not the bench contract - evidence for the config table in
DESIGN.md; writes JSON lines to stdout."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from quadruped_locomotion_b200 import capi, legmodel, synth  # noqa: E402

CONFIGS = [
    ("C1", "C1", 1, np.float64, "single 4-leg-stance QP (batch of 1 = one controller tick)"),
    ("C2", "C2", 65536, np.float64, "65 536 trot states (two stance legs)"),
    ("C3", "C3", 1 << 20, np.float64, "2^20 randomised states, FP64"),
    ("C4", "C3", 1 << 20, np.float32, "2^20 randomised states through the FP32 interface (per-GPU share)"),
    ("C5", "C5", 1 << 21, np.float64, "Monte-Carlo sweep, 2^21 samples (per-GPU share of 2^24 on 8 GPUs)"),
]


def main():
    dev = torch.device("cuda:0")
    sol = capi.Solver("quadruped_model")
    M = O.model_array(legmodel.load_model("quadruped_model"))
    stream = torch.cuda.current_stream().cuda_stream
    for name, gen, B, dt, what in CONFIGS:
        st = synth.make_states(gen, B)
        tdt = torch.float32 if dt == np.float32 else torch.float64
        d = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in st.items()}
        d = {k: (v.to(tdt) if v.dtype == torch.float64 else v) for k, v in d.items()}
        grf = torch.empty((12, B), dtype=tdt, device=dev); tau = torch.empty_like(grf)
        net = torch.empty((6, B), dtype=tdt, device=dev); flags = torch.empty(B, dtype=torch.int32, device=dev)

        def step():
            sol.solve_wrench(d["q"], d["quat"], d["wrench"], d["mask"], d["mu"], d["normals"], grf, tau, flags, net, stream=stream)

        for _ in range(3):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 20
        e0.record()
        for _ in range(reps):
            step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        n = min(B, 32768)
        ref = O.solve_wrench_batch(M, st["q"][:, :n], st["quat"][:, :n], st["wrench"][:, :n], st["mask"][:n], mu=st["mu"][:, :n],
                                   normals=st["normals"][:, :n], want_margin=True)
        g = grf[:, :n].double().cpu().numpy(); t = tau[:, :n].double().cpu().numpy()
        fl = flags[:n].cpu().numpy().view(np.uint32)
        eg = np.abs(g - ref["grf"]).max(0) / np.maximum(1.0, np.abs(ref["grf"]).max(0))
        et = np.abs(t - ref["tau"]).max(0) / np.maximum(1.0, np.abs(ref["tau"]).max(0))
        mism = ((fl ^ ref["flags"]) & capi.FLAG_PARITY_MASK) != 0
        thr = 1e-3 if dt == np.float32 else 1e-6
        print(json.dumps({"config": name, "what": what, "states": B, "dtype": "f32 interface, f64 core" if dt == np.float32 else "f64",
                          "ms_per_call": ms, "qp_per_s": B / (ms * 1e-3), "sample_checked": n,
                          "max_rel_force_err": float(eg.max()), "max_rel_torque_err": float(et.max()),
                          "flag_mismatches": int(mism.sum()), "flag_mismatches_above_margin": int((mism & (ref["margin"] > thr)).sum()),
                          "status_ok": int((((fl >> 24) & 7) == 0).sum())}), flush=True)


if __name__ == "__main__":
    main()
