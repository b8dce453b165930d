#!/usr/bin/env python
"""Device-resident throughput of the entry points next to the headline path (one JSON line each, for profiles/):
state mode (VMC prologue + solve, FP64 and FP32 arrays), the batch statistics kernel, the state generator, friction
margins, feet in world, the contact state machine."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from quadruped_locomotion_b200 import capi, synth  # noqa: E402


def timed(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    B = 1 << 20
    dev = torch.device("cuda:0")
    s = capi.Solver("quadruped_model")
    st = synth.make_states("C3", B)
    rng = np.random.default_rng(5)
    pose = np.concatenate([rng.normal(0, 0.05, (3, B)) + np.array([[0], [0], [0.45]]), st["quat"]])
    twist = rng.normal(0, 0.05, (6, B))
    tpose = pose + np.concatenate([rng.normal(0, 0.004, (3, B)), np.zeros((4, B))])
    ttwist = rng.normal(0, 0.05, (6, B))
    stream = torch.cuda.current_stream().cuda_stream
    for dt, name in ((torch.float64, "f64"), (torch.float32, "f32")):
        up = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev).to(dt)  # noqa: E731
        q, p, tw, tp, tt, mu, nr = (up(a) for a in (st["q"], pose, twist, tpose, ttwist, st["mu"], st["normals"]))
        mask = torch.from_numpy(st["mask"]).to(dev)
        grf = torch.empty((12, B), dtype=dt, device=dev); tau = torch.empty_like(grf)
        net = torch.empty((6, B), dtype=dt, device=dev); w = torch.empty((6, B), dtype=dt, device=dev)
        flags = torch.empty(B, dtype=torch.int32, device=dev)
        ms = timed(lambda: s.solve_state(q, p, tw, tp, tt, mask, mu, nr, grf, tau, flags, net, w, stream=stream))
        fl = flags.cpu().numpy().view(np.uint32)
        ok = int(((fl >> 24) & 7 == 0).sum())
        print(json.dumps({"entry": "qlb_solve_state" + ("_f32" if name == "f32" else ""), "states": B, "ms": ms, "states_per_s": B / ms * 1e3, "ok": ok,
                          "mean_rounds": float((fl >> 27).mean()), "hard_fraction": float(((fl >> 27) > 0).mean()),
                          "workload": "C3 joint angles / attitudes, base position and twist noise 0.05, targets 4 mm away"}), flush=True)
        ms3 = timed(lambda: s.solve_state(q, p, tw, p, tw, mask, mu, nr, grf, tau, flags, net, w, stream=stream))
        fl3 = flags.cpu().numpy().view(np.uint32)
        print(json.dumps({"entry": "qlb_solve_state" + ("_f32" if name == "f32" else ""), "states": B, "ms": ms3, "states_per_s": B / ms3 * 1e3,
                          "ok": int(((fl3 >> 24) & 7 == 0).sum()), "mean_rounds": float((fl3 >> 27).mean()), "hard_fraction": float(((fl3 >> 27) > 0).mean()),
                          "workload": "the same states with target = feedback (what a plan preview asks: gravity compensation only)"}), flush=True)
        ms = timed(lambda: s.solve_state(q, p, tw, tp, tt, mask, mu, nr, grf, tau, flags, net, w, stream=stream), reps=2, warm=1)
        if name == "f64":   # the same call in wrench mode on the wrench the prologue produced: what the prologue costs
            quat0 = up(st["quat"])
            ms2 = timed(lambda: s.solve_wrench(q, quat0, w, mask, mu, nr, grf, tau, flags, net, stream=stream))
            fl2 = flags.cpu().numpy().view(np.uint32)
            print(json.dumps({"entry": "qlb_solve_wrench on the wrenches of that run", "states": B, "ms": ms2, "states_per_s": B / ms2 * 1e3,
                              "mean_rounds": float((fl2 >> 27).mean())}), flush=True)
        if name == "f64":
            wr = up(st["wrench"])
            ms = timed(lambda: s.batch_stats(flags, wr, net, stream=stream))
            print(json.dumps({"entry": "qlb_batch_stats (kernel + 248-byte copy + sync)", "states": B, "ms": ms, "states_per_s": B / ms * 1e3, "bytes_per_state": 100,
                              "achieved_gbs": B * 100 / ms * 1e-6}), flush=True)
            quat = up(st["quat"])
            margin = torch.empty(B, dtype=dt, device=dev); mn = torch.empty(B, dtype=dt, device=dev)
            ms = timed(lambda: s.friction_margins(grf, quat, mask, mu, nr, margin, mn, stream=stream))
            print(json.dumps({"entry": "qlb_friction_margins", "states": B, "ms": ms, "states_per_s": B / ms * 1e3}), flush=True)
            feet = torch.empty((12, B), dtype=dt, device=dev)
            ms = timed(lambda: s.feet_in_world(q, p, feet, stream=stream))
            print(json.dumps({"entry": "qlb_feet_in_world", "states": B, "ms": ms, "states_per_s": B / ms * 1e3}), flush=True)
            Bg = 1 << 21
            gq = torch.empty((12, Bg), dtype=dt, device=dev); gquat = torch.empty((4, Bg), dtype=dt, device=dev)
            gw = torch.empty((6, Bg), dtype=dt, device=dev); gmu = torch.empty((4, Bg), dtype=dt, device=dev)
            gm = torch.empty(Bg, dtype=torch.uint8, device=dev)
            ms = timed(lambda: s.generate_states("C5", Bg, start=0, q=gq, quat=gquat, wrench=gw, mask=gm, mu=gmu, stream=stream))
            print(json.dumps({"entry": "qlb_generate_states (C5)", "states": Bg, "ms": ms, "states_per_s": Bg / ms * 1e3}), flush=True)


if __name__ == "__main__":
    main()
