#!/usr/bin/env python
"""Design prototype (CPU, numpy): how many equality-constrained rounds do the active-set strategies need?
Not product code, not a test: it informed the round logic of the fused kernel."""
import sys, os, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O
from quadruped_locomotion_b200 import legmodel, synth

def eqp(G, g0, D, d, W):
    n = g0.size
    W = sorted(W)
    k = len(W)
    if k == 0:
        return np.linalg.solve(G, -g0), {}
    K = np.zeros((n + k, n + k)); K[:n, :n] = G; K[:n, n:] = -D[W].T; K[n:, :n] = D[W]
    rhs = np.concatenate([-g0, d[W]])
    s = np.linalg.solve(K, rhs)
    return s[:n], dict(zip(W, s[n:]))

def pair_of(r, ns):
    # rows: 0..ns-1 normal; ns+4k+{0,1}: tangent 1 pair, {2,3}: tangent 2 pair. returns the opposite row or None
    if r < ns: return None
    k, j = divmod(r - ns, 4)
    return ns + 4 * k + (j ^ 1)

def pdas(G, g0, D, d, ns, max_rounds, tol=1e-9):
    """heuristic rounds as in the round-1 kernel. returns (x, W, rounds) or None if not verified"""
    x, _ = eqp(G, g0, D, d, [])
    s = D @ x - d
    sc = max(1.0, np.abs(x).max())
    if (s >= -tol * sc).all():
        return x, set(), 0
    W = set()
    viol = {r for r in range(len(d)) if s[r] < -tol * sc}
    for r in sorted(viol):
        p = pair_of(r, ns)
        if p is not None and p in viol and (s[p] < s[r] or (s[p] == s[r] and p < r)): continue
        W.add(r)
    for rnd in range(1, max_rounds + 1):
        x, u = eqp(G, g0, D, d, W)
        s = D @ x - d
        sc = max(1.0, np.abs(x).max())
        drop = {r for r in W if u[r] < -1e-12 * sc}
        add = {r for r in range(len(d)) if r not in W and s[r] < -tol * sc}
        if not drop and not add:
            return x, W, rnd
        if rnd <= 2:
            for r in drop: W.discard(r)
            for r in sorted(add):
                p = pair_of(r, ns)
                if p in drop: continue
                if p is not None and p in W: continue
                if p is not None and p in add and (s[p] < s[r] or (s[p] == s[r] and p < r)): continue
                W.add(r)
        else:
            if drop:
                r = min(drop, key=lambda r: u[r]); W.discard(r)
            else:
                r = min(add, key=lambda r: s[r]); p = pair_of(r, ns)
                if p is not None and p in W: W.discard(p)
                W.add(r)
    return None

def gi(G, g0, D, d, ns, tol=1e-9, max_it=60, select="most"):
    """Goldfarb-Idnani written as a sequence of EQP solves + interpolation. returns (x, W, rounds)"""
    m = len(d)
    x, _ = eqp(G, g0, D, d, [])
    W = []; u = {}
    rounds = 0
    p = None; up = 0.0
    while rounds < max_it:
        if p is None:
            s = D @ x - d
            sc = max(1.0, np.abs(x).max())
            cand = [r for r in range(m) if r not in W and s[r] < -tol * sc and (pair_of(r, ns) not in W)]
            if not cand:
                # a violated row whose pair is active cannot be represented: must not happen when normal row exists
                allv = [r for r in range(m) if r not in W and s[r] < -tol * sc]
                if allv: return None
                return x, set(W), rounds
            p = min(cand, key=lambda r: s[r]) if select == "most" else cand[0]
            up = 0.0
        xn, un = eqp(G, g0, D, d, W + [p])
        rounds += 1
        # interpolate multipliers of W from u to un
        t = 1.0; k = None
        for r in W:
            if un[r] < 0 and u[r] - un[r] > 0:
                tr = u[r] / (u[r] - un[r])
                if tr < t: t, k = tr, r
        if k is None:
            x = xn; u = dict(un); W = W + [p]; p = None
        else:
            x = x + t * (xn - x)
            for r in W: u[r] = u[r] + t * (un[r] - u[r])
            up = up + t * (un[p] - up)
            W.remove(k); del u[k]
    return None

def main():
    cfg = sys.argv[1] if len(sys.argv) > 1 else "C3"
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
    st = synth.make_states(cfg, B)
    M = O.model_array(legmodel.load_model("quadruped_model"))
    ref = O.solve_wrench_batch(M, st["q"], st["quat"], st["wrench"], st["mask"], mu=st["mu"], normals=st["normals"], want_margin=True)
    hp = np.zeros(64, int); hg = np.zeros(64, int); hgs = np.zeros(64, int); fail_p = 0; fail_g = 0; nhard = 0
    hybrid = np.zeros(64, int)
    bad = 0
    t0 = time.time()
    for i in range(B):
        qp = O.assemble(M, st["q"][:, i], st["quat"][:, i], st["wrench"][:, i], st["mask"][i], mu=st["mu"][:, i], normals=st["normals"][:, i])
        if qp["ns"] == 0: continue
        G, g0, D, d, ns = qp["G"], qp["g0"], qp["D"], qp["d"], qp["ns"]
        rp = pdas(G, g0, D, d, ns, 12)
        if rp is not None and rp[2] == 0:
            continue
        nhard += 1
        if rp is None: fail_p += 1
        else: hp[rp[2]] += 1
        rg = gi(G, g0, D, d, ns)
        if rg is None: fail_g += 1
        else:
            hg[rg[1] and rg[2]] += 1
            # check against oracle forces
            xr = np.concatenate([ref["grf"][3 * l:3 * l + 3, i] for l in qp["legs"]])
            e = np.abs(rg[0] - xr).max() / max(1, np.abs(xr).max())
            if e > 1e-8: bad += 1
        # hybrid: pdas up to 3 rounds then GI from scratch (+1 round for the restart)
        r3 = pdas(G, g0, D, d, ns, 3)
        if r3 is not None: hybrid[r3[2]] += 1
        elif rg is not None: hybrid[3 + 1 + rg[2]] += 1
    print(cfg, "B", B, "hard", nhard, "time %.1fs" % (time.time() - t0))
    print("pdas rounds hist (1..):", hp[:16], "fail(>12)", fail_p)
    print("gi rounds hist:", hg[:24], "fail", fail_g, "mismatch vs oracle", bad)
    print("hybrid rounds hist:", hybrid[:28], "mean over hard %.3f" % ((hybrid * np.arange(64)).sum() / max(1, hybrid.sum())))
    print("mean pdas rounds %.3f ; mean gi rounds %.3f" % ((hp * np.arange(64)).sum() / max(1, hp.sum()), (hg * np.arange(64)).sum() / max(1, hg.sum())))

if __name__ == "__main__" and (len(sys.argv) < 2 or sys.argv[1] != "dbas"):
    main()


def dbas(G, g0, D, d, ns, tol=1e-9, max_it=60, utol=1e-12):
    """dual block active set: block adds of every violated row, ratio test on the multipliers (see DESIGN)."""
    m = len(d)
    x, _ = eqp(G, g0, D, d, [])
    F = []; u = {}
    rounds = 0
    full = True
    while rounds < max_it:
        if full:
            s = D @ x - d
            sc = max(1.0, np.abs(x).max())
            viol = [r for r in range(m) if r not in F and s[r] < -tol * sc]
            if not viol:
                return x, set(F), rounds
            added = 0
            for r in sorted(viol):
                p = pair_of(r, ns)
                if p is not None and p in F: continue
                if p is not None and p in viol and (s[p] < s[r] or (s[p] == s[r] and p < r)): continue
                F.append(r); u[r] = 0.0; added += 1
            if added == 0: return None
        xn, un = eqp(G, g0, D, d, F)
        rounds += 1
        gs = max(1.0, max(abs(v) for v in un.values()))
        t = 1.0; blk = []
        for r in F:
            if un[r] < -utol * gs:
                tr = u[r] / (u[r] - un[r]) if u[r] > 0 else 0.0
                if tr < t - 1e-15: t, blk = tr, [r]
                elif tr <= t + 1e-15 and tr < 1.0: blk.append(r)
        if not blk:
            x = xn; u = dict(un); full = True
        else:
            x = x + t * (xn - x)
            for r in F: u[r] = u[r] + t * (un[r] - u[r])
            for r in blk:
                F.remove(r); del u[r]
            full = False
    return None


def main2():
    cfg = sys.argv[2] if len(sys.argv) > 2 else "C3"
    B = int(sys.argv[3]) if len(sys.argv) > 3 else 6000
    st = synth.make_states(cfg, B)
    M = O.model_array(legmodel.load_model("quadruped_model"))
    ref = O.solve_wrench_batch(M, st["q"], st["quat"], st["wrench"], st["mask"], mu=st["mu"], normals=st["normals"], want_margin=True)
    h = np.zeros(64, int); fail = 0; bad = 0; nhard = 0; flagbad = 0
    for i in range(B):
        qp = O.assemble(M, st["q"][:, i], st["quat"][:, i], st["wrench"][:, i], st["mask"][i], mu=st["mu"][:, i], normals=st["normals"][:, i])
        if qp["ns"] == 0: continue
        G, g0, D, d, ns = qp["G"], qp["g0"], qp["D"], qp["d"], qp["ns"]
        r = dbas(G, g0, D, d, ns)
        if r is None: fail += 1; continue
        if r[2] == 0: continue
        nhard += 1
        h[r[2]] += 1
        xr = np.concatenate([ref["grf"][3 * l:3 * l + 3, i] for l in qp["legs"]])
        e = np.abs(r[0] - xr).max() / max(1, np.abs(xr).max())
        if e > 1e-8: bad += 1
        bits = 0
        for rr in r[1]:
            if rr < ns: leg, row = qp["legs"][rr], 0
            else: leg, row = qp["legs"][(rr - ns) // 4], 1 + (rr - ns) % 4
            bits |= 1 << (4 + 5 * leg + row)
        if bits != (int(ref["flags"][i]) & 0xFFFFF0): flagbad += 1
    print("flag mismatches", flagbad)
    print(cfg, "dbas: hard", nhard, "rounds hist", h[:24], "fail", fail, "mismatch", bad, "mean %.3f" % ((h * np.arange(64)).sum() / max(1, h.sum())))


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "dbas":
    main2()
