"""Generic small dense QP entry (qlb_qp_dense) = the backend behind qp_solver::QuadraticProblemSolver::
minimize.  Oracle: the Goldfarb-Idnani port, itself pinned to the reference's QuadProg++ (oracle/_ref)."""
import subprocess

import numpy as np
import pytest

from quadruped_locomotion_b200 import build, capi


def _random_qps(rng, B, n, m, p):
    A = rng.normal(size=(B, n + 2, n))
    G = np.einsum("bki,bkj->bij", A, A) + 0.1 * np.eye(n)
    g0 = rng.normal(size=(B, n)) * 3
    x0 = rng.normal(size=(B, n))                      # a point that satisfies everything: feasible by construction
    CI = rng.normal(size=(B, n, m))
    ci0 = -np.einsum("bnm,bn->bm", CI, x0) + rng.uniform(0.0, 1.0, size=(B, m))
    CE = rng.normal(size=(B, n, p))
    ce0 = -np.einsum("bnp,bn->bp", CE, x0)
    return G, g0, CI, ci0, CE, ce0


def test_oracle_generic_qp_matches_reference_solver(oracle):
    """CPU: the port agrees with the reference's QuadProg++ on generic QPs, equalities included."""
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(7)
    for n, m, p in ((2, 3, 0), (3, 4, 1), (6, 10, 2), (12, 20, 3)):
        G, g0, CI, ci0, CE, ce0 = _random_qps(rng, 50, n, m, p)
        for b in range(50):
            D, d = CI[b].T, -ci0[b]
            a = oracle.solve_qp_gi(G[b], g0[b], D, d, CE=CE[b] if p else None, ce0=ce0[b] if p else None)
            r = oracle.solve_qp_ref(G[b], g0[b], D, d, CE=CE[b] if p else None, ce0=ce0[b] if p else None)
            assert np.array_equal(a["x"], r["x"]) and a["f"] == r["f"]
            # KKT: feasibility of the answer
            assert (CI[b].T @ a["x"] + ci0[b]).min() > -1e-9
            if p:
                assert np.abs(CE[b].T @ a["x"] + ce0[b]).max() < 1e-9


def _host_lib():
    """g++ build of the product's QP algorithm (csrc/qlb_qp_dense.cuh, one lane): the logic is checked on the CPU."""
    import ctypes as C
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    so = os.path.join(root, "tests", "native", "libqp_dense_host.so")
    src = os.path.join(root, "tests", "native", "qp_dense_host.cc")
    hdr = os.path.join(root, "quadruped_locomotion_b200", "csrc", "qlb_qp_dense.cuh")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-x", "c++", "-I" + os.path.dirname(hdr), "-o", so, src])
    return C.CDLL(so)


def _solve_host(lib, G, g0, CI=None, ci0=None, CE=None, ce0=None):
    import ctypes as C
    B, n = g0.shape
    m = 0 if CI is None else CI.shape[2]
    p = 0 if CE is None else CE.shape[2]
    soa = lambda a, k: np.ascontiguousarray(np.asarray(a, dtype=np.float64).reshape(B, k).T)  # noqa: E731
    arrs = [soa(G, n * n), soa(g0, n), soa(CE, n * p) if p else None, soa(ce0, p) if p else None,
            soa(CI, n * m) if m else None, soa(ci0, m) if m else None]
    x = np.zeros((n, B)); cost = np.zeros(B); st = np.zeros(B, np.uint32); act = np.zeros(B, np.uint32)
    ptr = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)  # noqa: E731
    lib.qp_dense_host(C.c_ulonglong(B), n, m, p, *[ptr(a) for a in arrs], ptr(x), ptr(cost), ptr(st), ptr(act))
    return dict(x=np.ascontiguousarray(x.T), cost=cost, status=st, active=act)


@pytest.mark.parametrize("n,m,p", [(2, 3, 0), (3, 4, 1), (6, 10, 2), (12, 20, 3), (12, 24, 0), (5, 0, 0), (4, 0, 2), (6, 24, 0)])
def test_algorithm_on_the_host_matches_oracle(oracle, n, m, p):
    """CPU: the same source that runs as the warp-cooperative CUDA kernel, compiled for one lane."""
    lib = _host_lib()
    rng = np.random.default_rng(100 + n + m + p)
    B = 129
    G, g0, CI, ci0, CE, ce0 = _random_qps(rng, B, n, m, p)
    out = _solve_host(lib, G, g0, CI if m else None, ci0 if m else None, CE if p else None, ce0 if p else None)
    assert (out["status"] == 0).all()
    for b in range(B):
        a = oracle.solve_qp_gi(G[b], g0[b], CI[b].T if m else np.zeros((0, n)), -ci0[b] if m else np.zeros(0),
                               CE=CE[b] if p else None, ce0=ce0[b] if p else None)
        assert np.abs(out["x"][b] - a["x"]).max() <= 1e-10 * max(1.0, np.abs(a["x"]).max())
        assert abs(out["cost"][b] - a["f"]) <= 1e-9 * max(1.0, abs(a["f"]))
        assert int(out["active"][b]) == sum(1 << i for i in range(m) if a["active"][i])


def test_algorithm_on_the_host_failure_statuses_and_zero_equality_column(kats):
    lib = _host_lib()
    G = np.eye(2)[None].repeat(4, 0)
    g0 = np.zeros((4, 2))
    CI = np.zeros((4, 2, 2)); ci0 = np.zeros((4, 2))
    CI[:, 0, 0] = 1.0; ci0[:, 0] = -1.0
    CI[:, 0, 1] = -1.0; ci0[:, 1] = 2.0
    ci0[1, 1] = 0.5
    G[2, 1, 1] = -1.0
    g0[3, 0] = np.nan
    out = _solve_host(lib, G, g0, CI, ci0)
    assert list(out["status"]) == [0, 1, 2, 2]
    np.testing.assert_allclose(out["x"][0], [1.0, 0.0], atol=1e-15)
    assert np.isinf(out["cost"][1]) and not out["x"][1].any()
    s = kats["solver"]
    CIk = np.array(s["D"], float).T[None]
    out = _solve_host(lib, np.array(s["G"], float)[None], np.array(s["g0"], float)[None], CIk, -np.array(s["d"], float)[None],
                      CE=np.zeros((1, 2, 1)), ce0=np.zeros((1, 1)))
    np.testing.assert_allclose(out["x"][0], s["x"], atol=1e-14)
    assert abs(out["cost"][0] - s["f"]) < 1e-13 and out["active"][0] == 0b011


@pytest.fixture(scope="module")
def solver(qlb_built):
    s = capi.Solver("quadruped_model")
    yield s
    s.close()


@pytest.mark.gpu
@pytest.mark.parametrize("n,m,p", [(2, 3, 0), (3, 4, 1), (6, 10, 2), (12, 20, 3), (12, 24, 0), (5, 0, 0), (4, 0, 2)])
def test_matches_oracle(solver, oracle, n, m, p):
    rng = np.random.default_rng(100 + n + m + p)
    B = 257
    G, g0, CI, ci0, CE, ce0 = _random_qps(rng, B, n, m, p)
    out = solver.qp_dense_numpy(G, g0, CI if m else None, ci0 if m else None, CE if p else None, ce0 if p else None)
    assert (out["status"] == 0).all()
    for b in range(B):
        a = oracle.solve_qp_gi(G[b], g0[b], CI[b].T if m else np.zeros((0, n)), -ci0[b] if m else np.zeros(0),
                               CE=CE[b] if p else None, ce0=ce0[b] if p else None)
        scale = max(1.0, np.abs(a["x"]).max())
        assert np.abs(out["x"][b] - a["x"]).max() <= 1e-10 * scale
        assert abs(out["cost"][b] - a["f"]) <= 1e-9 * max(1.0, abs(a["f"]))
        bits = sum(1 << i for i in range(m) if a["active"][i])
        assert int(out["active"][b]) == bits


@pytest.mark.gpu
def test_reference_known_answers(solver, kats):
    s = kats["solver"]  # qp_solver/src/main.cc data, zero equality column dropped
    CI = np.array(s["D"], float).T[None]
    out = solver.qp_dense_numpy(np.array(s["G"], float)[None], np.array(s["g0"], float)[None], CI, -np.array(s["d"], float)[None])
    np.testing.assert_allclose(out["x"][0], s["x"], atol=1e-14)
    assert abs(out["cost"][0] - s["f"]) < 1e-13 and out["active"][0] == 0b011
    # the same data WITH the reference's zero equality column (main.cc:64-70): treated as absent
    out = solver.qp_dense_numpy(np.array(s["G"], float)[None], np.array(s["g0"], float)[None], CI, -np.array(s["d"], float)[None],
                                CE=np.zeros((1, 2, 1)), ce0=np.zeros((1, 1)))
    np.testing.assert_allclose(out["x"][0], s["x"], atol=1e-14)


@pytest.mark.gpu
def test_failure_statuses(solver):
    G = np.eye(2)[None].repeat(4, 0)
    g0 = np.zeros((4, 2))
    CI = np.zeros((4, 2, 2)); ci0 = np.zeros((4, 2))
    CI[:, 0, 0] = 1.0; ci0[:, 0] = -1.0        # x0 >= 1
    CI[:, 0, 1] = -1.0; ci0[:, 1] = 2.0        # x0 <= 2
    ci0[1, 1] = 0.5                             # problem 1: x0 >= 1 and x0 <= 0.5  -> infeasible
    G[2, 1, 1] = -1.0                           # problem 2: indefinite Hessian
    g0[3, 0] = np.nan                           # problem 3: NaN input
    out = solver.qp_dense_numpy(G, g0, CI, ci0)
    assert list(out["status"]) == [0, 1, 2, 2]
    np.testing.assert_allclose(out["x"][0], [1.0, 0.0], atol=1e-15)
    assert np.isinf(out["cost"][1]) and not out["x"][1].any()


@pytest.mark.gpu
def test_pose_optimisation_known_answer_through_the_adapter(qlb_built):
    """qp_solver/test/PoseOptimizationQpTest.cpp:21-52 expects (0, 0, 0.3) within 1e-3."""
    demo = build.build_host_demo(which="qp_demo")
    r = subprocess.run([demo], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    rows = [ln.split() for ln in r.stdout.strip().splitlines()]
    x0 = [float(v) for v in rows[0][:3]]
    np.testing.assert_allclose(x0, [0.0, 0.0, 0.3], atol=1e-12)
    assert rows[0][3] == "0" and rows[0][4] == "0"
    x1 = [float(v) for v in rows[1][:3]]   # optimum pushed onto the support-polygon edge x = 1
    np.testing.assert_allclose(x1, [1.0, 0.0, 0.3], atol=1e-12)
    assert rows[1][3] == "0" and rows[1][4] == "1"
    # the sequential-QP loop (sequencequadraticproblemsolver.cpp:18-102) on min |x - (2,1)|^2 s.t. |x|^2 <= 1
    xs = [float(v) for v in rows[2][:2]]
    np.testing.assert_allclose(xs, np.array([2.0, 1.0]) / np.sqrt(5.0), atol=1e-9)
    assert 2 <= int(rows[2][2]) <= 20 and rows[2][3] == "0"


def test_qp_adapter_compiles(qlb_built):
    build.build_host_demo(which="qp_demo")
