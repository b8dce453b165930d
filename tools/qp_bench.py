#!/usr/bin/env python
"""Throughput of the generic small-QP entry (qlb_qp_dense, one warp per problem) on device-resident batches of the
shapes the reference's pose optimisation solves: 3 variables / 4 inequalities (PoseOptimizationQP) and 6 variables /
10 inequalities (the SQP step).  Prints one JSON line per shape."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from quadruped_locomotion_b200 import capi  # noqa: E402


def main():
    s = capi.Solver("quadruped_model")
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(0)
    for n, m, p, B in ((3, 4, 0, 1 << 18), (6, 10, 0, 1 << 18), (12, 20, 0, 1 << 16)):
        A = rng.normal(size=(B, n + 2, n))
        G = np.einsum("bki,bkj->bij", A, A) + 0.1 * np.eye(n)
        g0 = rng.normal(size=(B, n)) * 3
        x0 = rng.normal(size=(B, n))
        CI = rng.normal(size=(B, n, m))
        ci0 = -np.einsum("bnm,bn->bm", CI, x0) + rng.uniform(0.0, 1.0, size=(B, m))
        soa = lambda a, k: torch.from_numpy(np.ascontiguousarray(a.reshape(B, k).T)).to(dev)  # noqa: E731
        dG, dg, dCI, dci = soa(G, n * n), soa(g0, n), soa(CI, n * m), soa(ci0, m)
        x = torch.zeros((n, B), dtype=torch.float64, device=dev); cost = torch.zeros(B, dtype=torch.float64, device=dev)
        st = torch.zeros(B, dtype=torch.int32, device=dev); act = torch.zeros(B, dtype=torch.int32, device=dev)
        stream = torch.cuda.current_stream().cuda_stream

        def run():
            rc = s.lib.qlb_qp_dense(s._ctx, B, n, m, p, dG.data_ptr(), dg.data_ptr(), None, None, dCI.data_ptr(), dci.data_ptr(),
                                    x.data_ptr(), cost.data_ptr(), st.data_ptr(), act.data_ptr(), stream)
            assert rc == 0
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        nact = float(torch.tensor([bin(int(v)).count("1") for v in act[:4096].cpu().numpy().view(np.uint32)], dtype=torch.float64).mean())
        print(json.dumps({"kernel": "qlb_qp_dense_kernel", "n": n, "m": m, "p": p, "problems": B, "ms": ms, "qp_per_s": B / ms * 1e3,
                          "status_ok": int((st == 0).sum().item()), "mean_active_rows": nact}), flush=True)


if __name__ == "__main__":
    main()
