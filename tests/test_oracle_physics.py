"""Physics invariants of the CPU oracle.  The reference's step arithmetic (mujoco.mjx) is not available, so the oracle is
checked against first principles and against an independent numpy formulation of M(q) (mjcf.mass_matrix)."""
import copy
import ctypes as C

import numpy as np
import pytest

from conftest import make_handle
from open_duck_playground_b200 import mjcf

G = 9.81


def _dump(lib, h):
    stride = lib.lib.oduck_debug_stride()
    out = np.zeros((h.n, stride))
    lib.lib.oduck_debug_forward.argtypes = [C.c_void_p, C.c_void_p]
    lib.check(lib.lib.oduck_debug_forward(h.h, out.ctypes.data))
    return out


def _free_model(model, dt=None):
    m = copy.deepcopy(model)
    A = m.arrays
    A["dof_damping"][:] = 0; A["dof_frictionloss"][:] = 0; A["act_kp"][:] = 0; A["jnt_limited"][:] = 0
    if dt is not None:
        A["timestep"] = np.array(dt)
    return m


def _random_state(model, n, seed, z=2.0):
    rng = np.random.default_rng(seed)
    qpos = np.tile(model.key_qpos[: model.nq], (n, 1)).astype(np.float32)
    qpos[:, 2] = z
    quat = rng.normal(size=(n, 4)); qpos[:, 3:7] = quat / np.linalg.norm(quat, axis=1, keepdims=True)
    qpos[:, 7:] += rng.uniform(-0.3, 0.3, (n, model.nq - 7)).astype(np.float32)
    qvel = rng.uniform(-2, 2, (n, model.nv)).astype(np.float32)
    return qpos, qvel


def test_mass_matrix_matches_jacobian_formula(oracle, model_backlash, poly_table):
    n = 4
    h = make_handle(oracle, model_backlash, poly_table, n)
    qpos, qvel = _random_state(model_backlash, n, 0)
    h.set_state(qpos.ctypes.data, qvel.ctypes.data, 0)
    d = _dump(oracle, h)
    nv = model_backlash.nv
    for i in range(n):
        M = d[i, :1024].reshape(32, 32)[:nv, :nv]
        Mref = mjcf.mass_matrix(model_backlash, qpos[i].astype(np.float64))
        assert np.abs(M - Mref).max() < 1e-9
        assert np.allclose(M, M.T)


def _advance(m, qpos, qvel, eps):
    """qpos moved along qvel for a time eps (MuJoCo's mj_integratePos: world-frame base translation, base quaternion
    right-multiplied by the exponential of the BODY-frame angular velocity, hinge angles += v eps)."""
    q = qpos.astype(np.float64).copy()
    q[:3] += eps * qvel[:3]
    w = qvel[3:6] * eps
    ang = np.linalg.norm(w)
    dq = np.array([1.0, 0, 0, 0]) if ang < 1e-300 else np.concatenate([[np.cos(ang / 2)], np.sin(ang / 2) * w / ang])
    q[3:7] = mjcf.quat_mul(q[3:7], dq)
    q[7:] += eps * qvel[6:]
    return q


def test_bias_force_matches_projected_newton_euler(oracle, model_backlash, poly_table):
    """qfrc_bias (RNE: Coriolis + centrifugal + gyroscopic + gravity) against an independent formulation: Newton-Euler per body
    projected through the body Jacobians (Kane / d'Alembert, valid for MuJoCo's quasi-velocities -- world-frame base linear,
    body-frame base angular):  c = sum_b  Jp_b^T m_b (dJp_b/dt v - g)  +  Jr_b^T (I_b dJr_b/dt v + w_b x I_b w_b),
    with dJ/dt v taken as a CENTRAL DIFFERENCE of the Jacobians along the flow of v (mjcf.body_jacobians at q(+-eps)); no
    recursion over the tree, no spatial algebra -- nothing shared with the oracle's RNE but the model arrays."""
    n = 4
    m = model_backlash
    A = m.arrays
    h = make_handle(oracle, m, poly_table, n)
    qpos, qvel = _random_state(m, n, 7)
    h.set_state(qpos.ctypes.data, qvel.ctypes.data, 0)
    d = _dump(oracle, h)
    nv, eps = m.nv, 1e-6
    grav = np.array([0.0, 0.0, -G])
    for i in range(n):
        q, v = qpos[i, : m.nq].astype(np.float64), qvel[i, : nv].astype(np.float64)
        _, xmat, _, jacp, jacr = mjcf.body_jacobians(m, q)
        _, _, _, jp1, jr1 = mjcf.body_jacobians(m, _advance(m, q, v, eps))
        _, _, _, jp0, jr0 = mjcf.body_jacobians(m, _advance(m, q, v, -eps))
        c = np.zeros(nv)
        for b in range(1, m.nbody):
            Ri = xmat[b] @ mjcf.quat_to_mat(A["body_iquat"][b])
            Iw = Ri @ np.diag(A["body_inertia"][b]) @ Ri.T
            a_lin = (jp1[b] - jp0[b]) @ v / (2 * eps)                  # COM acceleration at qacc = 0
            a_ang = (jr1[b] - jr0[b]) @ v / (2 * eps)
            w = jacr[b] @ v
            c += A["body_mass"][b] * jacp[b].T @ (a_lin - grav) + jacr[b].T @ (Iw @ a_ang + np.cross(w, Iw @ w))
        bias = d[i, 1024:1024 + nv]
        # measured: max |difference| 2e-10 on forces up to 21 (N, N m), of which 0.2 - 0.65 is velocity-dependent
        assert np.abs(bias - c).max() < 1e-8, (i, np.abs(bias - c).max(), np.abs(c).max())
        # and the terms matter in this state: gravity alone is far from the answer
        c_grav = sum(-A["body_mass"][b] * jacp[b].T @ grav for b in range(1, m.nbody))
        assert np.abs(c - c_grav).max() > 0.1


def test_energy_drift_is_first_order_in_dt(oracle, model_backlash, poly_table):
    """No damping / friction / actuation / contact: E = KE + PE must be conserved up to the integrator's O(dt) error."""
    drift = []
    for dt in (0.002, 0.001, 0.0005):
        m = _free_model(model_backlash, dt)
        h = make_handle(oracle, m, poly_table, 2)
        qpos, qvel = _random_state(m, 2, 1)
        qpos[:, 3:7] = [1, 0, 0, 0]
        h.set_state(qpos.ctypes.data, qvel.ctypes.data, 0)

        def energy():
            q, v = h.buffer_numpy("QPOS"), h.buffer_numpy("QVEL")
            out = []
            for i in range(2):
                M = mjcf.mass_matrix(m, q[i, : m.nq].copy())
                xipos = mjcf.body_jacobians(m, q[i, : m.nq].copy())[2]
                out.append(0.5 * v[i, : m.nv] @ M @ v[i, : m.nv] + G * np.sum(m.body_mass[: m.nbody] * xipos[:, 2]))
            return np.array(out)
        e0 = energy()
        h.physics_substeps(0, int(round(0.1 / dt)))
        drift.append(np.abs(energy() - e0).max())
    assert drift[0] < 0.05                                           # of ~45 J
    assert 1.7 < drift[0] / drift[1] < 2.3 and 1.7 < drift[1] / drift[2] < 2.3


def test_free_fall_and_momentum(oracle, model_backlash, poly_table):
    m = _free_model(model_backlash)
    h = make_handle(oracle, m, poly_table, 3)
    qpos, qvel = _random_state(m, 3, 2)
    h.set_state(qpos.ctypes.data, qvel.ctypes.data, 0)

    def com_vel():
        q, v = h.buffer_numpy("QPOS"), h.buffer_numpy("QVEL")
        out = []
        for i in range(3):
            jacp = mjcf.body_jacobians(m, q[i, : m.nq].copy())[3]
            mass = m.body_mass[: m.nbody]
            out.append(sum(mass[b] * jacp[b] @ v[i, : m.nv] for b in range(m.nbody)) / mass.sum())
        return np.array(out)
    v0 = com_vel()
    h.physics_substeps(0, 50)
    dv = com_vel() - v0
    assert np.abs(dv[:, :2]).max() < 1e-3 and np.abs(dv[:, 2] + G * 0.1).max() < 2e-3     # only gravity acts on the COM


def test_standing_equilibrium_and_sensors(oracle, model_backlash, poly_table):
    h = make_handle(oracle, model_backlash, poly_table, 1)
    h.physics_substeps(0, 1500)                         # settle from the keyframe under the home ctrl (3 s)
    nfr, nlim = 14, 24
    f = h.buffer_numpy("EFC_FORCE")[0]
    dist = h.buffer_numpy("CONTACT_DIST")[0]
    assert (dist[:8] < 0).sum() >= 6 and np.all(dist[8:] == 1)
    # pyramid edge forces sum to the normal force: sum(f_edges) * 1 (each edge has unit normal component)
    normal = f[nfr + nlim: nfr + nlim + 32].sum()
    weight = model_backlash.body_mass[: model_backlash.nbody].sum() * G
    assert abs(normal - weight) / weight < 0.02
    sd = h.buffer_numpy("SENSORDATA")[0]
    assert np.abs(sd[0:3]).max() < 5e-3 and np.abs(sd[3:6]).max() < 5e-3          # gyro, local linvel ~ 0 at rest
    assert abs(np.linalg.norm(sd[6:9]) - G) < 0.05 and sd[11] > 0.99              # accelerometer reads +g, upvector z ~ 1
    q = h.buffer_numpy("QPOS")[0]
    assert 0.12 < q[2] < 0.2 and abs(np.linalg.norm(q[3:7]) - 1) < 1e-12


def _make_frame(n):
    """MJX math.make_frame: tangent from the y axis unless the normal is close to it, then from z."""
    a = n / np.linalg.norm(n)
    b0 = np.array([0.0, 1.0, 0.0]) if -0.5 < a[1] < 0.5 else np.array([0.0, 0.0, 1.0])
    b = b0 - a * (a @ b0)
    b /= np.linalg.norm(b)
    return np.stack([a, b, np.cross(a, b)])


def test_newton_gradient_and_hessian_from_independent_contact_jacobians(oracle, model_backlash, poly_table):
    """The Newton solver's gradient, search direction and Hessian (constraint.py / solver.py: grad = M qacc - qfrc_smooth - J^T f,
    H = M + J_active^T D J_active, search = -H^-1 grad) rebuilt outside the oracle: constraint rows J from mjcf.body_jacobians (point Jacobian of the foot
    body at the contact position = COM Jacobian + rotational Jacobian x lever arm; the oracle builds them from cdof about the
    subtree COM), pyramid edges n +- mu t from the contact normal, unit rows for dry friction and joint limits; only the
    per-row regularisation D and reference acceleration aref are taken from the oracle's dump.  A settled stance with both feet
    down, one knee pushed to its limit so that a limit row is live."""
    m = model_backlash
    A = m.arrays
    h = make_handle(oracle, m, poly_table, 1)
    h.physics_substeps(0, 400)
    q = h.buffer_numpy("QPOS").copy(); v = h.buffer_numpy("QVEL").copy()
    lim_j = next(j for j in range(m.njnt) if A["jnt_limited"][j] and A["jnt_type"][j] != mjcf.JNT_FREE)
    q[0, A["jnt_qposadr"][lim_j]] = A["jnt_range"][lim_j][1] + 0.02            # 0.02 rad beyond the upper limit
    qf, vf = q.astype(np.float32), (v + 0.05).astype(np.float32)             # some velocity everywhere: the friction rows leave zero
    h.set_state(qf.ctypes.data, vf.ctypes.data, 0)
    warm = h.buffer_numpy("QACC_WARM")[0].copy()
    d = _dump(oracle, h)[0]
    nv = m.nv
    M = d[:1024].reshape(32, 32)[:nv, :nv]
    qfrc_smooth, qacc_smooth = d[1056:1056 + nv], d[1088:1088 + nv]
    start = warm[:nv] if d[2536] < d[2537] else qacc_smooth                  # solver.py: the cheaper of warm start and qacc_smooth
    _, xmat, xipos, jacp, jacr = mjcf.body_jacobians(m, qf[0, : m.nq].astype(np.float64))
    rows = []                                                                # (J, D, aref, kind, floss)
    for dof in range(nv):
        if A["dof_frictionloss"][dof] > 0:
            e = np.zeros(nv); e[dof] = 1
            rows.append((e, d[1184 + dof], d[1264 + dof], "friction", float(A["dof_frictionloss"][dof])))
    n_lim = 0
    for j in range(m.njnt):
        if not A["jnt_limited"][j] or A["jnt_type"][j] == mjcf.JNT_FREE:
            continue
        dof, x = A["jnt_dofadr"][j], float(qf[0, A["jnt_qposadr"][j]])
        lo, hi = x - A["jnt_range"][j][0], A["jnt_range"][j][1] - x
        if min(lo, hi) < 0:
            e = np.zeros(nv); e[dof] = 1.0 if lo < hi else -1.0
            rows.append((e, d[1216 + dof], d[1296 + dof], "limit", 0.0)); n_lim += 1
    assert n_lim >= 1
    n_con = 0
    for c in range(8):
        if not d[1120 + c] < 0:
            continue
        n_con += 1
        b = int(A["foot_body"][c // 4])
        pos, frame = d[1136 + 3 * c: 1139 + 3 * c], _make_frame(d[2560 + 3 * c: 2563 + 3 * c])
        lever = pos - xipos[b]
        jp = jacp[b] + np.stack([np.cross(jacr[b][:, k], lever) for k in range(nv)], axis=1)      # floor = world: body1 adds nothing
        jf = frame @ jp
        mu = float(A["floor_friction"])
        for e4 in range(4):
            rows.append((jf[0] + (1 if e4 % 2 == 0 else -1) * mu * jf[1 + e4 // 2], d[1248 + c], d[1328 + 4 * c + e4], "contact", 0.0))
    assert n_con >= 6
    def assemble(qacc):
        H, f_total, n_active = M.copy(), np.zeros(nv), 0
        for J, D, aref, kind, floss in rows:
            x = J @ qacc - aref
            if kind == "friction":
                rf = floss / D
                if x <= -rf:
                    f = floss
                elif x >= rf:
                    f = -floss
                else:
                    f = -D * x
                    H += D * np.outer(J, J)
            else:
                f = -D * x if x < 0 else 0.0
                if x < 0:
                    H += D * np.outer(J, J)
                    n_active += 1
            f_total += J * f
        return H, M @ qacc - qfrc_smooth - f_total, n_active

    H0, grad0, n_active = assemble(start)
    assert n_active >= 4
    g_or, search_or = d[1408:1408 + nv], d[1376:1376 + nv]
    assert np.abs(grad0 - g_or).max() < 1e-9 * max(1.0, np.abs(g_or).max()), (np.abs(grad0 - g_or).max(), np.abs(g_or).max())
    # the search direction the solver takes from there is the Newton step of the independently assembled system
    assert np.abs(np.linalg.solve(H0, -grad0) - search_or).max() < 1e-7 * max(1.0, np.abs(search_or).max())
    # the dumped Hessian is the one at the END of the iteration (active set of the new qacc)
    H1, _, _ = assemble(d[1736:1736 + nv])
    H_or = d[4096:4096 + 1024].reshape(32, 32)[:nv, :nv]
    assert np.abs(H_or - M).max() > 1.0                                        # the constraint part dominates: not a test of M again
    assert np.abs(H1 - H_or).max() < 1e-9 * np.abs(H_or).max(), np.abs(H1 - H_or).max() / np.abs(H_or).max()


def test_joint_limit_pushes_back(oracle, model_backlash, poly_table):
    m = copy.deepcopy(model_backlash)
    m.arrays["act_kp"][:] = 0
    h = make_handle(oracle, m, poly_table, 1)
    qpos = m.key_qpos[: m.nq].astype(np.float32)[None].copy()
    qpos[0, 2] = 1.0
    hi = m.jnt_range[1, 1]
    qpos[0, 7] = hi + 0.05                                   # left_hip_yaw past its upper limit
    qvel = np.zeros((1, m.nv), np.float32)
    h.set_state(qpos.ctypes.data, qvel.ctypes.data, 0)
    h.forward()
    f = h.buffer_numpy("EFC_FORCE")[0]
    assert f[14] > 0                                          # first limit row active, force positive (one-sided)
    assert h.buffer_numpy("QACC")[0, 6] < 0                   # accelerates back inside the range
    h.physics_substeps(0, 200)
    assert h.buffer_numpy("QPOS")[0, 7] < hi + 0.01


def test_frictionloss_holds_a_slow_joint(oracle, model_backlash, poly_table):
    """Dry friction (0.068 N m) cancels small torques exactly: a head joint with a tiny velocity stops instead of coasting."""
    m = copy.deepcopy(model_backlash)
    A = m.arrays
    A["act_kp"][:] = 0; A["dof_damping"][:] = 0; A["gravity"][:] = 0
    h = make_handle(oracle, m, poly_table, 1)
    qpos = m.key_qpos[: m.nq].astype(np.float32)[None].copy(); qpos[0, 2] = 1.0
    qvel = np.zeros((1, m.nv), np.float32)
    d = int(m.jnt_dofadr[m.joint_id("head_yaw")])
    qvel[0, d] = 0.01
    h.set_state(qpos.ctypes.data, qvel.ctypes.data, 0)
    h.physics_substeps(0, 100)
    assert abs(h.buffer_numpy("QVEL")[0, d]) < 1e-5


def test_nan_state_terminates(oracle, model_backlash, poly_table):
    h = make_handle(oracle, model_backlash, poly_table, 2)
    keys = np.array([[0, 1], [0, 2]], np.uint32)
    h.reset(keys.ctypes.data)
    qpos = h.buffer_numpy("QPOS")[:, :31].astype(np.float32).copy()
    qpos[1, 9] = np.nan
    h.set_state(qpos.ctypes.data, 0, 0)
    act = np.zeros((2, 14), np.float32)
    h.step(act.ctypes.data)
    assert h.buffer_numpy("DONE").tolist() == [0.0, 1.0]
    assert not np.isnan(h.buffer_numpy("QPOS")[1]).any()      # auto-reset restored the first state
    assert abs(h.buffer_numpy("REWARD")[1] - 20.0 * 0.02) < 1e-12   # every term is nan_to_num-ed: only "alive" survives


def _ff_poses(model, n, seed):
    """Random in-range joint poses with both hip rolls driven inward, so that a fair share of them has foot-foot contact."""
    rng = np.random.default_rng(seed)
    q = np.tile(model.key_qpos[: model.nq], (n, 1)).astype(np.float32)
    q[:, 2] = 0.6
    for j in range(1, model.njnt):
        lo, hi = model.jnt_range[j]
        q[:, model.jnt_qposadr[j]] = rng.uniform(lo, hi, n)
    q[:, model.jnt_qposadr[model.joint_id("left_hip_roll")]] = rng.uniform(0.3, 0.436, n)
    q[:, model.jnt_qposadr[model.joint_id("right_hip_roll")]] = -rng.uniform(0.3, 0.436, n)
    for name in ("left_hip_pitch", "right_hip_pitch", "left_knee", "right_knee", "left_ankle", "right_ankle"):
        q[:, model.jnt_qposadr[model.joint_id(name)]] = model.key_qpos[model.jnt_qposadr[model.joint_id(name)]] + rng.normal(0, 0.1, n)
    return q


def test_foot_foot_contacts_are_sane_and_repulsive(oracle, model_backlash, poly_table):
    m = copy.deepcopy(model_backlash)
    m.arrays["act_kp"][:] = 0; m.arrays["gravity"][:] = 0
    n = 256
    h = make_handle(oracle, m, poly_table, n)
    q = _ff_poses(m, n, 0)
    v = np.zeros((n, m.nv), np.float32)
    h.set_state(q.ctypes.data, v.ctypes.data, v.ctypes.data)
    d = _dump(oracle, h)
    dist, pos, nrm = d[:, 1128:1132], d[:, 1136 + 24:1136 + 36].reshape(n, 4, 3), d[:, 2560 + 24:2560 + 27]
    hit = (dist < 0).any(axis=1)
    assert 10 < hit.sum() < n                                     # the rare path is exercised, and not always
    for i in np.nonzero(hit)[0]:
        xpos, xmat, _, _ = mjcf.world_kinematics(m, q[i].astype(np.float64))
        cl = xpos[7] + xmat[7] @ m.foot_center[0]; cr = xpos[16] + xmat[16] @ m.foot_center[1]
        assert abs(np.linalg.norm(nrm[i]) - 1) < 1e-9
        assert np.dot(nrm[i], cr - cl) > 0                        # normal points from the left foot (geom1) to the right foot (geom2)
        for c in range(4):
            if dist[i, c] < 0:
                assert np.linalg.norm(pos[i, c] - cl) < float(m.foot_radius) + 0.02 and np.linalg.norm(pos[i, c] - cr) < float(m.foot_radius) + 0.02
                assert dist[i, c] > -0.05
    # contact forces are repulsive: after 20 ms of free motion the deepest penetration has shrunk in (almost) every env
    h.physics_substeps(0, 10)
    d2 = _dump(oracle, h)
    before, after = dist[hit].min(axis=1), np.minimum(d2[hit, 1128:1132].min(axis=1), 0)
    assert (after > before).mean() > 0.9
    assert np.all(np.isfinite(h.buffer_numpy("QPOS")))


def test_no_foot_foot_contact_in_nominal_poses(oracle, model_backlash, poly_table):
    h = make_handle(oracle, model_backlash, poly_table, 64)
    keys = np.stack([np.zeros(64, np.uint32), np.arange(64, dtype=np.uint32)], 1)
    h.reset(keys.ctypes.data)
    assert np.all(h.buffer_numpy("CONTACT_DIST")[:, 8:12] == 1)
