"""The oracle's rotation conventions against an independent library (scipy.spatial.transform).

KDL, kindr and RBDL are not vendored in the reference tree, so the oracle restates their published semantics ("parity
unpinned").  These tests put a third, independently written implementation beside the two
restatements: URDF fixed-axis roll-pitch-yaw and the joint chain (quadrupedkinematics.cpp:143-278,485-552), the
world->base rotation of a (w,x,y,z) quaternion (ContactForceDistribution.cpp:223,518) and the orientation error
-log(q*^-1 q) of VirtualModelController.cpp:120-124, all computed with scipy's Rotation class.  CPU only.
"""
import numpy as np
import pytest

from quadruped_locomotion_b200 import legmodel

Rot = pytest.importorskip("scipy.spatial.transform").Rotation


def _chain(leg, q):
    """Frames of the three revolute z joints and of the fixed foot joint: origin <xyz, rpy> then Rz(q)."""
    R, p = Rot.identity(), np.zeros(3)
    origins, axes, com_world = [], [], []
    for k in range(4):
        p = p + R.apply(leg["joint_xyz"][k])
        R = R * Rot.from_euler("xyz", leg["joint_rpy"][k])      # extrinsic x-y-z = URDF fixed-axis rpy
        if k < 3:
            origins.append(p.copy())
            axes.append(R.apply([0.0, 0.0, 1.0]))
            R = R * Rot.from_euler("z", q[k])
        com_world.append(p + R.apply(leg["link_com"][k]))
    return p, origins, axes, com_world


@pytest.mark.parametrize("name", ["quadruped_model", "simpledog"])
def test_fk_jacobian_gravity_against_scipy(oracle, models, name):
    mdl = legmodel.load_model(name)
    rng = np.random.default_rng(7)
    g = np.array([0.4, -0.3, -9.7])
    for leg in range(4):
        for _ in range(5):
            q = rng.uniform(-2.0, 2.0, 3)
            foot, J, G = oracle.leg_kinematics(models[name], leg, q, grav=g)
            L = mdl["legs"][leg]
            p, origins, axes, coms = _chain(L, q)
            np.testing.assert_allclose(foot, p, atol=1e-12)
            for j in range(3):
                # geometric Jacobian column: z_j x (p_foot - p_j)   (KDL ChainJntToJacSolver, position rows)
                np.testing.assert_allclose(J[:, j], np.cross(axes[j], p - origins[j]), atol=1e-12)
                # KDL JntToGravity: torque of the weights of the links behind joint j about its axis, sign as RNE
                tq = -sum(L["link_mass"][k] * g @ np.cross(axes[j], coms[k] - origins[j]) for k in range(j, 4))
                assert abs(G[j] - tq) < 1e-10


def _wxyz(r):
    x, y, z, w = r.as_quat()
    return np.array([w, x, y, z])


def test_base_rotation_and_friction_frame_against_scipy(oracle, models):
    """The assembled constraint rows carry n_base = R_wb n_world and t1 = normalize(n x R_wb e_y)
    (ContactForceDistribution.cpp:223,237,301-309): rebuild them with scipy from the same quaternion."""
    rng = np.random.default_rng(11)
    M = models["quadruped_model"]
    for _ in range(6):
        r = Rot.from_euler("ZYX", [rng.uniform(-3, 3), rng.uniform(-0.3, 0.3), rng.uniform(-0.3, 0.3)])
        quat = _wxyz(r)
        nw = np.array([0.1, -0.05, 1.0]); nw /= np.linalg.norm(nw)
        q = np.tile([0.1, 0.7, -1.4], 4)
        asm = oracle.assemble(M, q, quat, np.array([0, 0, 500.0, 0, 0, 0]), 0b1111, mu=np.full(4, 0.5),
                              normals=np.tile(nw, 4))
        D = np.asarray(asm["D"])
        n = r.inv().apply(nw)                      # world -> base
        t1 = np.cross(n, r.inv().apply([0.0, 1.0, 0.0])); t1 /= np.linalg.norm(t1)
        t2 = np.cross(n, t1); t2 /= np.linalg.norm(t2)
        ns = 4
        for k in range(4):
            cols = slice(3 * k, 3 * k + 3)
            np.testing.assert_allclose(D[k, cols], n, atol=1e-12)
            rows = D[ns + 4 * k: ns + 4 * k + 4, cols]
            np.testing.assert_allclose(rows[0], 0.5 * n + t1, atol=1e-12)
            np.testing.assert_allclose(rows[1], 0.5 * n - t1, atol=1e-12)
            np.testing.assert_allclose(rows[2], 0.5 * n + t2, atol=1e-12)
            np.testing.assert_allclose(rows[3], 0.5 * n - t2, atol=1e-12)


def test_vmc_orientation_error_and_gravity_against_scipy(oracle):
    rng = np.random.default_rng(13)
    for _ in range(8):
        r = Rot.from_rotvec(rng.normal(0, 0.4, 3))
        rd = Rot.from_rotvec(rng.normal(0, 0.4, 3))
        pose = np.concatenate([rng.normal(0, 0.1, 3), _wxyz(r)])
        tpose = np.concatenate([pose[:3], _wxyz(rd)])
        z6 = np.zeros(6)
        # (1) only the rotational proportional gain: torque = kp * (-log(q*^-1 q)) = kp * log(q^-1 q*)
        p = oracle.default_vmc_params()
        for a in range(3):
            p.kp_t[a] = p.kd_t[a] = p.kff_t[a] = p.kd_r[a] = p.kff_r[a] = 0.0
            p.kp_r[a] = 1.0
        p.gravity_pct = 0.0
        w = oracle.vmc_wrench(pose, z6, tpose, z6, params=p)
        np.testing.assert_allclose(w[3:], (r.inv() * rd).as_rotvec(), atol=1e-12)
        np.testing.assert_allclose(w[:3], 0.0, atol=1e-12)
        # (2) only gravity compensation: force = -m_total g_base, torque = sum r x (-m_leg g_base)
        p = oracle.default_vmc_params()
        for a in range(3):
            p.kp_t[a] = p.kd_t[a] = p.kff_t[a] = p.kp_r[a] = p.kd_r[a] = p.kff_r[a] = 0.0
        w = oracle.vmc_wrench(pose, z6, pose, z6, params=p)
        gb = r.inv().apply([0.0, 0.0, -p.gravity])
        mtot = p.torso_mass + sum(p.leg_mass[l] for l in range(4))
        np.testing.assert_allclose(w[:3], -mtot * gb, rtol=1e-13, atol=1e-10)
        tq = sum(np.cross(np.array([p.leg_base_position[l][a] for a in range(3)]), -p.leg_mass[l] * gb) for l in range(4))
        np.testing.assert_allclose(w[3:], tq, atol=1e-9)
