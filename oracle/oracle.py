"""ctypes face of the CPU oracle.  TEST INFRASTRUCTURE ONLY.

May be imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs, and by nothing under quadruped_locomotion_b200/.
"""
from __future__ import annotations

import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import build_oracle  # noqa: E402

SOLVER_GI, SOLVER_IPM, SOLVER_REF = 0, 1, 2
MAX_N, MAX_M = 12, 24

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


class QoParams(C.Structure):
    _fields_ = [("S", C.c_double * 6), ("W", C.c_double), ("fmin", C.c_double), ("gravity", C.c_double)]


class QoVmcParams(C.Structure):
    _fields_ = [("kp_t", C.c_double * 3), ("kd_t", C.c_double * 3), ("kff_t", C.c_double * 3),
                ("kp_r", C.c_double * 3), ("kd_r", C.c_double * 3), ("kff_r", C.c_double * 3),
                ("torso_mass", C.c_double), ("leg_mass", C.c_double * 4),
                ("leg_base_position", (C.c_double * 3) * 4), ("com", C.c_double * 3),
                ("gravity_pct", C.c_double), ("gravity", C.c_double)]


class QoQp(C.Structure):
    _fields_ = [("ns", C.c_int), ("n", C.c_int), ("m", C.c_int), ("leg_of_slot", C.c_int * 4),
                ("G", C.c_double * (MAX_N * MAX_N)), ("g0", C.c_double * MAX_N),
                ("D", C.c_double * (MAX_M * MAX_N)), ("d", C.c_double * MAX_M),
                ("A", C.c_double * (6 * MAX_N)), ("b", C.c_double * 6),
                ("foot", C.c_double * 12), ("jac", C.c_double * 36), ("gtau", C.c_double * 12),
                ("bad_input", C.c_int)]


_lib = None
_ref = None
EXT_FN = C.CFUNCTYPE(C.c_double, C.c_int, C.c_int, _dp, _dp, _dp, _dp, _dp)


def _as(a):
    return a.ctypes.data_as(_dp)


def lib():
    global _lib, _ref
    if _lib is None:
        so, ref = build_oracle.build()
        _lib = C.CDLL(so)
        _lib.qo_goldfarb_idnani.restype = C.c_double
        _lib.qo_goldfarb_idnani.argtypes = [C.c_int, C.c_int, C.c_int, _dp, _dp, _dp, _dp, _dp, _dp, _dp, _ip, _dp, _ip]
        _lib.qo_ipm.restype = C.c_int
        _lib.qo_ipm.argtypes = [C.c_int, C.c_int, _dp, _dp, _dp, _dp, C.c_double, C.c_int, _dp, _dp, _ip, _dp, _ip]
        _lib.qo_solve_wrench_batch.restype = C.c_int
        _lib.qo_leg_kinematics.restype = None
        _lib.qo_assemble.restype = None
        _lib.qo_vmc_wrench.restype = None
        if ref:
            _ref = C.CDLL(ref)
            _ref.qref_solve_quadprog.restype = C.c_double
            _ref.qref_solve_quadprog.argtypes = [C.c_int, C.c_int, _dp, _dp, _dp, _dp, _dp]
            _ref.qref_solve_quadprog_eq.restype = C.c_double
            _ref.qref_solve_quadprog_eq.argtypes = [C.c_int, C.c_int, C.c_int, _dp, _dp, _dp, _dp, _dp, _dp, _dp]
    return _lib


def have_ref() -> bool:
    lib()
    return _ref is not None


def default_params() -> QoParams:
    p = QoParams()
    lib().qo_default_params(C.byref(p))
    return p


def default_vmc_params() -> QoVmcParams:
    p = QoVmcParams()
    lib().qo_default_vmc_params(C.byref(p))
    return p


def model_array(model: dict) -> np.ndarray:
    """dict from legmodel.load_model -> (4,40) float64 in qo_leg_model layout."""
    out = np.zeros((4, 40))
    for i, leg in enumerate(model["legs"]):
        out[i, 0:12] = np.asarray(leg["joint_xyz"], dtype=np.float64).ravel()
        out[i, 12:24] = np.asarray(leg["joint_rpy"], dtype=np.float64).ravel()
        out[i, 24:28] = np.asarray(leg["link_mass"], dtype=np.float64)
        out[i, 28:40] = np.asarray(leg["link_com"], dtype=np.float64).ravel()
    return np.ascontiguousarray(out)


def leg_kinematics(model_arr, leg: int, q, grav=(0.0, 0.0, -9.8)):
    q = np.ascontiguousarray(q, dtype=np.float64)
    g = np.ascontiguousarray(grav, dtype=np.float64)
    foot, jac, gt = np.zeros(3), np.zeros(9), np.zeros(3)
    lib().qo_leg_kinematics(_as(model_arr[leg]), _as(q), _as(g), _as(foot), _as(jac), _as(gt))
    return foot, jac.reshape(3, 3), gt


def solve_qp_gi(G, g0, D, d, CE=None, ce0=None):
    """Goldfarb-Idnani port.  Returns dict(x, f, active, u, iterations)."""
    G = np.ascontiguousarray(G, dtype=np.float64); g0 = np.ascontiguousarray(g0, dtype=np.float64)
    n = g0.size
    D = np.ascontiguousarray(D, dtype=np.float64).reshape(-1, n) if np.size(D) else np.zeros((0, n))
    d = np.ascontiguousarray(d, dtype=np.float64)
    m = d.size
    p = 0
    cep = cp = None
    if CE is not None and np.size(CE):
        CE = np.ascontiguousarray(CE, dtype=np.float64); ce0 = np.ascontiguousarray(ce0, dtype=np.float64)
        p = ce0.size
        cep, cp = _as(CE), _as(ce0)
    x = np.zeros(n); u = np.zeros(max(m, 1)); act = np.zeros(max(m, 1), dtype=np.int32); it = C.c_int(0)
    f = lib().qo_goldfarb_idnani(n, m, p, _as(G), _as(g0), cep, cp, _as(D), _as(d), _as(x),
                                 act.ctypes.data_as(_ip), _as(u), C.byref(it))
    return dict(x=x, f=f, active=act[:m].astype(bool), u=u[:m], iterations=it.value)


def solve_qp_ipm(G, g0, D, d, tol=1e-9, max_iter=40, x0=None):
    G = np.ascontiguousarray(G, dtype=np.float64); g0 = np.ascontiguousarray(g0, dtype=np.float64)
    n = g0.size
    D = np.ascontiguousarray(D, dtype=np.float64).reshape(-1, n); d = np.ascontiguousarray(d, dtype=np.float64)
    m = d.size
    x = np.zeros(n); u = np.zeros(m); act = np.zeros(m, dtype=np.int32); it = C.c_int(0)
    x0p = _as(np.ascontiguousarray(x0, dtype=np.float64)) if x0 is not None else None
    st = lib().qo_ipm(n, m, _as(G), _as(g0), _as(D), _as(d), tol, max_iter, x0p, _as(x),
                      act.ctypes.data_as(_ip), _as(u), C.byref(it))
    return dict(x=x, status=st, active=act.astype(bool), u=u, iterations=it.value)


def solve_qp_ref(G, g0, D, d, CE=None, ce0=None):
    """The reference's own QuadProg++ (oracle/_ref).  Returns dict(x, f)."""
    lib()
    if _ref is None:
        raise RuntimeError("oracle/_ref/libquadprog_ref.so not built (needs /root/reference)")
    G = np.ascontiguousarray(G, dtype=np.float64); g0 = np.ascontiguousarray(g0, dtype=np.float64)
    n = g0.size
    D = np.ascontiguousarray(D, dtype=np.float64).reshape(-1, n) if np.size(D) else np.zeros((0, n))
    d = np.ascontiguousarray(d, dtype=np.float64)
    x = np.zeros(n)
    if CE is not None and np.size(CE):
        CE = np.ascontiguousarray(CE, dtype=np.float64); ce0 = np.ascontiguousarray(ce0, dtype=np.float64)
        f = _ref.qref_solve_quadprog_eq(n, d.size, ce0.size, _as(G), _as(g0), _as(CE), _as(ce0), _as(D), _as(d), _as(x))
    else:
        f = _ref.qref_solve_quadprog(n, d.size, _as(G), _as(g0), _as(D), _as(d), _as(x))
    return dict(x=x, f=f)


def assemble(model_arr, q, quat, wrench, mask, mu=None, normals=None, params=None, mu_default=0.6):
    """One state -> dict with the packed QP (G, g0, D, d, A, b) and kinematics."""
    prm = params or default_params()
    q = np.ascontiguousarray(q, dtype=np.float64); quat = np.ascontiguousarray(quat, dtype=np.float64)
    wrench = np.ascontiguousarray(wrench, dtype=np.float64)
    mup = _as(np.ascontiguousarray(mu, dtype=np.float64)) if mu is not None else None
    nrp = _as(np.ascontiguousarray(normals, dtype=np.float64)) if normals is not None else None
    qp = QoQp()
    lib().qo_assemble(_as(model_arr), C.byref(prm), _as(q), _as(quat), _as(wrench), C.c_uint(int(mask)),
                      mup, C.c_double(mu_default), nrp, C.byref(qp))
    n, m = qp.n, qp.m
    return dict(ns=qp.ns, n=n, m=m, legs=list(qp.leg_of_slot)[:qp.ns],
                G=np.array(qp.G[:n * n]).reshape(n, n), g0=np.array(qp.g0[:n]),
                D=np.array(qp.D[:m * n]).reshape(m, n), d=np.array(qp.d[:m]),
                A=np.array(qp.A[:6 * n]).reshape(6, n), b=np.array(qp.b[:]),
                foot=np.array(qp.foot[:]).reshape(4, 3), jac=np.array(qp.jac[:]).reshape(4, 3, 3),
                gtau=np.array(qp.gtau[:]).reshape(4, 3), bad_input=qp.bad_input)


def solve_wrench_batch(model_arr, q, quat, wrench, mask, mu=None, normals=None, params=None,
                       mu_default=0.6, solver=SOLVER_GI, nsolves=1, threads=0, want_margin=False):
    """Whole pipeline over a batch (SoA arrays [C,B]).  Returns dict(grf, tau, flags, netwrench[, margin])."""
    prm = params or default_params()
    q = np.ascontiguousarray(q, dtype=np.float64); B = q.shape[1]
    quat = np.ascontiguousarray(quat, dtype=np.float64); wrench = np.ascontiguousarray(wrench, dtype=np.float64)
    mask = np.ascontiguousarray(mask, dtype=np.uint8)
    mu_a = np.ascontiguousarray(mu, dtype=np.float64) if mu is not None else None
    nr_a = np.ascontiguousarray(normals, dtype=np.float64) if normals is not None else None
    grf = np.zeros((12, B)); tau = np.zeros((12, B)); nw = np.zeros((6, B)); flags = np.zeros(B, dtype=np.uint32)
    margin = np.zeros(B) if want_margin else None
    ext = None
    if solver == SOLVER_REF:
        lib()
        if _ref is None:
            raise RuntimeError("oracle/_ref not built")
        ext = C.cast(_ref.qref_solve_quadprog, C.c_void_p)
    rc = lib().qo_solve_wrench_batch(
        _as(model_arr), C.byref(prm), C.c_long(B), _as(q), _as(quat), _as(wrench),
        mask.ctypes.data_as(C.POINTER(C.c_uint8)), _as(mu_a) if mu_a is not None else None,
        C.c_double(mu_default), _as(nr_a) if nr_a is not None else None, C.c_int(solver), ext,
        C.c_int(nsolves), C.c_int(threads), _as(grf), _as(tau), flags.ctypes.data_as(C.POINTER(C.c_uint32)),
        _as(nw), _as(margin) if margin is not None else None)
    if rc != 0:
        raise RuntimeError(f"qo_solve_wrench_batch failed: {rc}")
    out = dict(grf=grf, tau=tau, flags=flags, netwrench=nw)
    if want_margin:
        out["margin"] = margin
    return out


def vmc_wrench(pose, twist, tpose, ttwist, params=None):
    p = params or default_vmc_params()
    a = [np.ascontiguousarray(v, dtype=np.float64) for v in (pose, twist, tpose, ttwist)]
    w = np.zeros(6)
    lib().qo_vmc_wrench(C.byref(p), _as(a[0]), _as(a[1]), _as(a[2]), _as(a[3]), _as(w))
    return w


# ---------------------------------------------------------------- SURVEY 8f rows 1 and 2 (numpy restatements)
def pack_robot_states(records: np.ndarray) -> dict:
    """What RosBalanceController::baseCommandCallback extracts from a free_gait_msgs/RobotState
    (ros_balance_controller.cpp:761-811): desired base pose (position, then the quaternion as kindr's
    RotationQuaternion(w, x, y, z)), twist, the twelve joint positions in LF, RF, RH, LH order, support flags
    and surface normals.  records: structured array with the fields of qlb_robot_state_record."""
    B = records.shape[0]
    o = records["base_orientation_xyzw"]
    pose = np.concatenate([records["base_position"].T, o[:, 3:4].T, o[:, 0:3].T])
    twist = np.concatenate([records["base_linear_velocity"].T, records["base_angular_velocity"].T])
    mask = np.zeros(B, dtype=np.uint8)
    for leg in range(4):
        mask |= ((records["support_leg"][:, leg] != 0).astype(np.uint8) << leg).astype(np.uint8)
    return dict(q=np.ascontiguousarray(records["joint_position"].T), pose=np.ascontiguousarray(pose),
                twist=np.ascontiguousarray(twist), mask=mask, normals=np.ascontiguousarray(records["surface_normal"].T))


def feet_in_world(model_arr, q, pose) -> np.ndarray:
    """position + R_bw FK(q) per leg (StateBatchComputer.cpp:64-77: getPositionWorldToFootInWorldFrame for
    every state of the batch).  q[12,B], pose[7,B] (position, quat wxyz) -> [12,B]."""
    q = np.asarray(q, dtype=np.float64); pose = np.asarray(pose, dtype=np.float64)
    B = q.shape[1]
    out = np.zeros((12, B))
    for i in range(B):
        w, x, y, z = pose[3:, i]
        R = np.array([[w * w + x * x - y * y - z * z, 2 * (x * y - w * z), 2 * (x * z + w * y)],
                      [2 * (x * y + w * z), w * w - x * x + y * y - z * z, 2 * (y * z - w * x)],
                      [2 * (x * z - w * y), 2 * (y * z + w * x), w * w - x * x - y * y + z * z]])
        for leg in range(4):
            foot, _, _ = leg_kinematics(model_arr, leg, q[3 * leg:3 * leg + 3, i])
            out[3 * leg:3 * leg + 3, i] = pose[:3, i] + R @ foot
    return out


# ---------------------------------------------------------------- swing-leg torques (numpy restatement)
def _rpy_rot(rpy):
    """URDF <origin rpy> -> rotation (child axes in parent coordinates), through the quaternion like urdfdom."""
    hr, hp, hy = 0.5 * rpy[0], 0.5 * rpy[1], 0.5 * rpy[2]
    x = np.sin(hr) * np.cos(hp) * np.cos(hy) - np.cos(hr) * np.sin(hp) * np.sin(hy)
    y = np.cos(hr) * np.sin(hp) * np.cos(hy) + np.sin(hr) * np.cos(hp) * np.sin(hy)
    z = np.cos(hr) * np.cos(hp) * np.sin(hy) - np.sin(hr) * np.sin(hp) * np.cos(hy)
    w = np.cos(hr) * np.cos(hp) * np.cos(hy) + np.sin(hr) * np.sin(hp) * np.sin(hy)
    n = np.sqrt(x * x + y * y + z * z + w * w)
    x, y, z, w = x / n, y / n, z / n, w / n
    return np.array([[w * w + x * x - y * y - z * z, 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), w * w - x * x + y * y - z * z, 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), w * w - x * x - y * y + z * z]])


def _skew(v):
    return np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]])


def limb_inverse_dynamics(limb: dict, q, qd, qdd, gravity=(0.0, -9.81, 0.0)) -> np.ndarray:
    """Recursive Newton-Euler for one limb in SPATIAL-VECTOR form, body coordinates (the algorithm a rigid-body
    dynamics library runs for InverseDynamics(model, Q, QDot, QDDot, Tau), model_test_header.cpp:460); fixed base,
    gravity as base acceleration.  `limb` = one entry of models/<name>.json["limb_dynamics"].  Deliberately a
    different formulation from the CUDA kernel (base-frame sums)."""
    S = np.array([0, 0, 1, 0, 0, 0.0])
    v = np.zeros(6); a = np.concatenate([np.zeros(3), -np.asarray(gravity, dtype=np.float64)])
    Xup, f, Is = [], [], []

    def crm(u):   # spatial motion cross product
        return np.block([[_skew(u[:3]), np.zeros((3, 3))], [_skew(u[3:]), _skew(u[:3])]])

    for i in range(3):
        E = _rpy_rot(limb["joint_rpy"][i]).T           # parent coordinates -> frame before the joint
        r = np.asarray(limb["joint_xyz"][i], dtype=np.float64)
        XT = np.block([[E, np.zeros((3, 3))], [-E @ _skew(r), E]])
        c, s = np.cos(q[i]), np.sin(q[i])
        Ej = np.array([[c, s, 0], [-s, c, 0], [0, 0, 1.0]])
        XJ = np.block([[Ej, np.zeros((3, 3))], [np.zeros((3, 3)), Ej]])
        X = XJ @ XT
        vJ = S * qd[i]
        v = X @ v + vJ
        a = X @ a + S * qdd[i] + crm(v) @ vJ
        m = limb["body_mass"][i]; cm = np.asarray(limb["body_com"][i], dtype=np.float64)
        i6 = limb["body_inertia"][i]
        Ic = np.array([[i6[0], i6[1], i6[2]], [i6[1], i6[3], i6[4]], [i6[2], i6[4], i6[5]]])
        C = _skew(cm)
        I = np.block([[Ic + m * C @ C.T, m * C], [m * C.T, m * np.eye(3)]])
        Xup.append(X); Is.append(I)
        f.append(I @ a - crm(v).T @ (I @ v))           # v x* (I v) = -crm(v)^T (I v)
    tau = np.zeros(3)
    for i in (2, 1, 0):
        tau[i] = S @ f[i]
        if i > 0:
            f[i - 1] = f[i - 1] + Xup[i].T @ f[i]
    return tau


def limb_lagrangian_torques(limb: dict, q, qd, qdd, gravity=(0.0, -9.81, 0.0), h=1e-5) -> np.ndarray:
    """The same torques from the Lagrange equations with finite differences of the kinetic and potential
    energy - no recursion, no spatial algebra: pins limb_inverse_dynamics."""
    g = np.asarray(gravity, dtype=np.float64)

    def frames(qv):
        R = np.eye(3); p = np.zeros(3); out = []
        for i in range(3):
            p = p + R @ np.asarray(limb["joint_xyz"][i], dtype=np.float64)
            R = R @ _rpy_rot(limb["joint_rpy"][i])
            z = R[:, 2].copy()
            c, s = np.cos(qv[i]), np.sin(qv[i])
            R = R @ np.array([[c, -s, 0], [s, c, 0], [0, 0, 1.0]])
            out.append((R.copy(), p.copy(), z))
        return out

    def coms(qv):
        return [p + R @ np.asarray(limb["body_com"][i], dtype=np.float64) for i, (R, p, z) in enumerate(frames(qv))]

    def energy(qv, qdv):
        fr = frames(qv)
        T = 0.0
        cp = [coms(qv + h * e) for e in np.eye(3)]
        cm = [coms(qv - h * e) for e in np.eye(3)]
        w = np.zeros(3)
        for i, (R, p, z) in enumerate(fr):
            w = w + z * qdv[i]
            vc = sum((cp[j][i] - cm[j][i]) / (2 * h) * qdv[j] for j in range(3))
            i6 = limb["body_inertia"][i]
            Ic = np.array([[i6[0], i6[1], i6[2]], [i6[1], i6[3], i6[4]], [i6[2], i6[4], i6[5]]])
            T += 0.5 * limb["body_mass"][i] * vc @ vc + 0.5 * w @ (R @ Ic @ R.T) @ w
        U = -sum(limb["body_mass"][i] * g @ c for i, c in enumerate(coms(qv)))
        return T, U

    q = np.asarray(q, dtype=np.float64); qd = np.asarray(qd, dtype=np.float64); qdd = np.asarray(qdd, dtype=np.float64)
    hq = 1e-4
    dT_dqd = lambda qv, qdv: np.array([(energy(qv, qdv + hq * e)[0] - energy(qv, qdv - hq * e)[0]) / (2 * hq) for e in np.eye(3)])  # noqa: E731
    tau = np.zeros(3)
    # d/dt (dT/dqd) = sum_k d(dT/dqd)/dq_k qd_k + d(dT/dqd)/dqd_k qdd_k
    for k, e in enumerate(np.eye(3)):
        tau += (dT_dqd(q + hq * e, qd) - dT_dqd(q - hq * e, qd)) / (2 * hq) * qd[k]
        tau += (dT_dqd(q, qd + hq * e) - dT_dqd(q, qd - hq * e)) / (2 * hq) * qdd[k]
    for j, e in enumerate(np.eye(3)):
        Tp, Up = energy(q + hq * e, qd); Tm, Um = energy(q - hq * e, qd)
        tau[j] += -(Tp - Tm) / (2 * hq) + (Up - Um) / (2 * hq)
    return tau


def swing_leg_torques(model: dict, model_arr, q, qd, qdd, ptarget=None, vtarget=None, gravity=(0.0, -9.81, 0.0),
                      acceleration_scale=0.5, kp=(0.0, 0.0, 0.0), kd=(0.0, 0.0, 0.0)) -> np.ndarray:
    """MyRobotSolver::update for every leg of every state (model_test_header.cpp:412-502): limb inverse dynamics
    at (q, qd, acceleration_scale qdd) + J^T (kp .* (p* - p) + kd .* (v* - J qd)).  SoA [12,B] in and out."""
    q = np.asarray(q, dtype=np.float64); B = q.shape[1]
    out = np.zeros((12, B))
    for i in range(B):
        for leg in range(4):
            sl = slice(3 * leg, 3 * leg + 3)
            tau = limb_inverse_dynamics(model["limb_dynamics"][leg], q[sl, i], qd[sl, i], acceleration_scale * qdd[sl, i], gravity)
            if ptarget is not None or vtarget is not None:
                foot, J, _ = leg_kinematics(model_arr, leg, q[sl, i])
                ep = ptarget[sl, i] - foot if ptarget is not None else np.zeros(3)
                ev = vtarget[sl, i] - J @ qd[sl, i] if vtarget is not None else np.zeros(3)
                tau = tau + J.T @ (np.asarray(kp) * ep + np.asarray(kd) * ev)
            out[sl, i] = tau
    return out


# ---------------------------------------------------------------- contact state machine (numpy restatement)
LIMB_INIT, LIMB_STANCE_NORMAL, LIMB_STANCE_SLIPPING, LIMB_STANCE_LOST_CONTACT, LIMB_SWING_NORMAL, LIMB_SWING_LATE_LIFTOFF, \
    LIMB_SWING_EARLY_TOUCHDOWN, LIMB_SWING_BUMPED_INTO_OBSTACLE, LIMB_SWING_LATELY_TOUCHDOWN = range(9)


def contact_fsm(desired, footstep, contact, phase, limb_state):
    """RosBalanceController::footContactsCallback (ros_balance_controller.cpp:1086-1140), one robot after the other,
    written as the nested ifs of the reference; then the support-leg decision of update() (:242-366).
    desired / footstep / contact: uint8 masks [B]; phase [4,B]; limb_state [4,B] previous states.
    Returns (new limb_state [4,B], stance mask [B])."""
    B = desired.shape[0]
    out = np.array(limb_state, dtype=np.uint8, copy=True)
    stance = np.zeros(B, dtype=np.uint8)
    for i in range(B):
        for k in range(4):
            is_contact = bool((contact[i] >> k) & 1)
            is_footstep = bool((footstep[i] >> k) & 1) if footstep is not None else True
            if not (desired[i] >> k) & 1:                      # limbs_desired_state == SwingNormal
                out[k, i] = LIMB_SWING_NORMAL
                if is_footstep:
                    if phase[k, i] > 0.5:
                        if is_contact:
                            out[k, i] = LIMB_SWING_EARLY_TOUCHDOWN
                    elif phase[k, i] > 0.2:
                        if is_contact:
                            out[k, i] = LIMB_SWING_BUMPED_INTO_OBSTACLE
            else:                                              # limbs_desired_state == StanceNormal
                if not is_footstep:
                    out[k, i] = LIMB_STANCE_NORMAL
                else:
                    if is_contact:
                        out[k, i] = LIMB_STANCE_NORMAL
                    elif phase[k, i] < 0.1:
                        out[k, i] = LIMB_SWING_LATELY_TOUCHDOWN
                    if phase[k, i] > 0.5:
                        if not is_contact:
                            out[k, i] = LIMB_STANCE_LOST_CONTACT
            if out[k, i] in (LIMB_STANCE_NORMAL, LIMB_SWING_EARLY_TOUCHDOWN, LIMB_INIT):
                stance[i] |= 1 << k
    return out, stance


def friction_margins(grf, quat, mask, mu, normals=None, fmin=10.0):
    """Per state: min over stance legs of min(mu fn - |f.t1|, mu fn - |f.t2|) / (mu fn), and min fn - F_min, with the
    friction frame of ContactForceDistribution.cpp:286-309."""
    B = grf.shape[1]
    margin = np.zeros(B); minn = np.zeros(B)
    for i in range(B):
        w, x, y, z = quat[:, i]
        R = np.array([[w * w + x * x - y * y - z * z, 2 * (x * y - w * z), 2 * (x * z + w * y)],
                      [2 * (x * y + w * z), w * w - x * x + y * y - z * z, 2 * (y * z - w * x)],
                      [2 * (x * z - w * y), 2 * (y * z + w * x), w * w - x * x - y * y + z * z]])
        best, bestn = np.inf, np.inf
        for k in range(4):
            if not (mask[i] >> k) & 1:
                continue
            nw = normals[3 * k:3 * k + 3, i] if normals is not None else np.array([0.0, 0.0, 1.0])
            n = R.T @ nw
            t1 = np.cross(n, R.T @ np.array([0.0, 1.0, 0.0])); t1 /= np.linalg.norm(t1)
            t2 = np.cross(n, t1); t2 /= np.linalg.norm(t2)
            f = grf[3 * k:3 * k + 3, i]
            fn = f @ n
            slack = min(mu[k, i] * fn - abs(f @ t1), mu[k, i] * fn - abs(f @ t2))
            best = min(best, slack / max(mu[k, i] * fn, 1e-300)); bestn = min(bestn, fn - fmin)
        if mask[i] & 0xF:
            margin[i], minn[i] = best, bestn
    return margin, minn
