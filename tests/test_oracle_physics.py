"""Physics invariants of the CPU oracle (the rigid-body part that no reference vector can pin):
  * all 18 accelerations of the ABA forward dynamics against an INDEPENDENT dense formulation (projected Newton-Euler /
    Kane: mass matrix and velocity-product terms from plain forward kinematics of the 13 bodies, Jacobians by unit
    velocities, their time derivative by central differences; 18 x 18 numpy solve) — `test_forward_dynamics_matches_dense`;
  * power balance dE/dt = tau . qd and angular-momentum rate = gravity torque about the origin (both see the joint torques /
    the internal forces, unlike the linear-momentum rate);
  * conservation laws in free flight, static equilibrium on the ground, and the go2 model constants against the numbers the
    reference documents."""
import numpy as np
import pytest

from spi_active_b200 import go2_model as gm
from spi_active_b200 import recorders


def _rand_state(rng, model, height=1.0):
    s = np.zeros(37)
    s[0:3] = [rng.uniform(-1, 1), rng.uniform(-1, 1), height]
    q = rng.standard_normal(4); s[3:7] = q / np.linalg.norm(q)
    s[7:13] = rng.uniform(-1, 1, 6)
    s[13:25] = np.asarray(model.q_default) + rng.uniform(-0.3, 0.3, 12)
    s[25:37] = rng.uniform(-3, 3, 12)
    return s


# ---- an independent forward-dynamics: numerical Lagrangian-free check through momentum --------------------
def _bodies_world(model, s):
    """World pose / velocity of the 13 bodies' centres of mass by plain forward kinematics (numpy, no spatial
    algebra) -> list of (mass, com_world, v_com_world, R_world, w_world, Ic_body)."""
    def quat_R(q):
        x, y, z, w = q
        return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                         [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                         [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])
    def axis_R(ax, a):
        c, sn = np.cos(a), np.sin(a)
        return np.array([[1, 0, 0], [0, c, -sn], [0, sn, c]]) if ax == 0 else np.array([[c, 0, sn], [0, 1, 0], [-sn, 0, c]])
    out = []
    Rb, pb, vb, wb = quat_R(s[3:7]), s[0:3], s[7:10], s[10:13]
    base = model.lumped_base()
    out.append((base.mass, pb + Rb @ np.asarray(base.com), vb + np.cross(wb, Rb @ np.asarray(base.com)), Rb, wb, base.matrix()))
    legs = model.lumped_leg_bodies()
    for leg in range(4):
        R, p, v, w = Rb, pb, vb, wb
        for j in range(3):
            i = 3 * leg + j
            r = np.asarray(model.joint_origin[i])
            p_new = p + R @ r
            v_new = v + np.cross(w, R @ r)
            axis = np.zeros(3); axis[model.joint_axis[i]] = 1.0
            w_new = w + (R @ axis) * s[25 + i]
            R_new = R @ axis_R(model.joint_axis[i], s[13 + i])
            b = legs[i]
            c = R_new @ np.asarray(b.com)
            out.append((b.mass, p_new + c, v_new + np.cross(w_new, c), R_new, w_new, b.matrix()))
            R, p, v, w = R_new, p_new, v_new, w_new
    return out


def _momentum(model, s):
    P = np.zeros(3); L = np.zeros(3); E = 0.0
    for m, c, v, R, w, Ic in _bodies_world(model, s):
        P += m * v
        Iw = R @ Ic @ R.T
        L += np.cross(c, m * v) + Iw @ w
        E += 0.5 * m * v @ v + 0.5 * w @ Iw @ w + m * 9.81 * c[2]
    return P, L, E


def test_model_constants_match_documented_numbers(nominal_model):
    """total mass 15.019 kg; base + heads 6.923 kg; 19 Isaac bodies."""
    assert abs(nominal_model.total_mass() - 15.019) < 1e-3
    assert abs(nominal_model.lumped_base().mass - 6.923) < 1e-9
    assert len(gm.BODY_NAMES) == 19 and len(gm.DOF_NAMES) == 12
    assert [gm.BODY_NAMES.index(f"{l}_foot") for l in gm.LEGS] == [4, 8, 14, 18]


def test_free_flight_conserves_momentum(oracle_lib, blob, nominal_model):
    """No contact, zero torque: linear momentum changes by m g t, angular momentum about the origin by the
    gravity torque; checked over 40 steps of 2.5 ms to integrator accuracy."""
    rng = np.random.default_rng(7)
    s0 = _rand_state(rng, nominal_model, height=5.0)
    P0, L0, E0 = _momentum(nominal_model, s0)
    n = 40
    s1 = oracle_lib.sim_step(blob, s0[None], np.zeros((1, 12)), n)[0]
    P1, L1, E1 = _momentum(nominal_model, s1)
    T = n * nominal_model.dt
    M = nominal_model.total_mass()
    np.testing.assert_allclose(P1 - P0, [0, 0, -9.81 * M * T], atol=2e-3)
    # horizontal angular momentum about the (moving) centre of mass is untouched by gravity: compare in the com frame
    def L_com(s):
        bodies = _bodies_world(nominal_model, s)
        m = sum(b[0] for b in bodies); c = sum(b[0] * b[1] for b in bodies) / m; v = sum(b[0] * b[2] for b in bodies) / m
        L = np.zeros(3)
        for mb, cb, vb, R, w, Ic in bodies:
            L += np.cross(cb - c, mb * (vb - v)) + (R @ Ic @ R.T) @ w
        return L
    np.testing.assert_allclose(L_com(s1), L_com(s0), atol=5e-3)
    assert abs(E1 - E0) / abs(E0) < 2e-3   # semi-implicit Euler: bounded energy error, no drift blow-up


def test_forward_dynamics_matches_momentum_rate(oracle_lib, blob, nominal_model):
    """d/dt of the total momentum computed from the oracle's accelerations by finite differences must equal the
    external force (gravity only, no contact): an ABA-independent check of the base + joint accelerations."""
    rng = np.random.default_rng(11)
    for _ in range(5):
        s = _rand_state(rng, nominal_model, height=5.0)
        tau = rng.uniform(-10, 10, 12)
        acc, qdd, _ = oracle_lib.forward_dynamics(blob, s, tau, with_contact=False, with_gravity=True)
        # advance velocities only by eps using the oracle accelerations (world frame)
        from math import isfinite
        def quat_R(q):
            x, y, z, w = q
            return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                             [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                             [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])
        R = quat_R(s[3:7])
        wb, vb = R.T @ s[10:13], R.T @ s[7:10]
        lin_world = R @ (acc[3:6] + np.cross(wb, vb))
        ang_world = R @ acc[0:3]
        eps = 1e-6
        # first-order propagate the full state by eps with the oracle's accelerations
        s2 = s.copy()
        s2[7:10] += eps * lin_world; s2[10:13] += eps * ang_world; s2[25:37] += eps * qdd
        s2[0:3] += eps * s[7:10]; s2[13:25] += eps * s[25:37]
        w = s[10:13]; x, y, z, ww = s[3:7]
        dq = 0.5 * np.array([w[0] * ww + w[1] * z - w[2] * y, w[1] * ww + w[2] * x - w[0] * z,
                             w[2] * ww + w[0] * y - w[1] * x, -(w[0] * x + w[1] * y + w[2] * z)])
        s2[3:7] = s[3:7] + eps * dq; s2[3:7] /= np.linalg.norm(s2[3:7])
        P0, L0, _ = _momentum(nominal_model, s)
        P1, L1, _ = _momentum(nominal_model, s2)
        M = nominal_model.total_mass()
        np.testing.assert_allclose((P1 - P0) / eps, [0, 0, -9.81 * M], atol=2e-3)
        assert all(isfinite(v) for v in qdd)


def _quat_R(q):
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def _world_accel(s, acc):
    """oracle base acceleration (spatial, base coordinates: [ang3, lin3]) -> classical world-frame (lin, ang) of the origin."""
    R = _quat_R(s[3:7])
    wb, vb = R.T @ s[10:13], R.T @ s[7:10]
    return R @ (acc[3:6] + np.cross(wb, vb)), R @ acc[0:3]


def _advance_config(s, eps):
    """q(t + eps) along the current generalized velocity, velocities untouched (exact quaternion exponential)."""
    s2 = s.copy()
    s2[0:3] += eps * s[7:10]
    s2[13:25] += eps * s[25:37]
    w = s[10:13]
    th = np.linalg.norm(w) * eps
    ax = w / max(np.linalg.norm(w), 1e-300)
    dq = np.array([*(ax * np.sin(0.5 * th)), np.cos(0.5 * th)])          # world-frame rotation: q2 = dq * q
    x1, y1, z1, w1 = dq; x2, y2, z2, w2 = s[3:7]
    s2[3:7] = [w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2, w1 * y2 - x1 * z2 + y1 * w2 + z1 * x2,
               w1 * z2 + x1 * y2 - y1 * x2 + z1 * w2, w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2]
    return s2


def _body_twists(model, s, nu=None):
    """[(v_com, w)] of the 13 bodies for generalized velocity nu = (v_base_world, w_base_world, qd) at configuration s."""
    t = s.copy()
    if nu is not None:
        t[7:10], t[10:13], t[25:37] = nu[0:3], nu[3:6], nu[6:18]
    return [(b[2], b[4]) for b in _bodies_world(model, t)]


def dense_forward_dynamics(model, s, tau, g=9.81):
    """nu_dot[18] by projected Newton-Euler:  sum_b J_b^T [m (a_b - g); I_b alpha_b + w_b x I_b w_b] = S^T tau  with
    a_b = Jv_b nu_dot + d/dt(Jv_b) nu etc.  Nothing here shares code or structure with the ABA."""
    nu = np.concatenate([s[7:10], s[10:13], s[25:37]])
    bodies = _bodies_world(model, s)
    nb = len(bodies)
    Jv = np.zeros((nb, 3, 18)); Jw = np.zeros((nb, 3, 18))
    for i in range(18):
        e = np.zeros(18); e[i] = 1.0
        for b, (v, w) in enumerate(_body_twists(model, s, e)):
            Jv[b, :, i], Jw[b, :, i] = v, w
    eps = 1e-6
    tp, tm = _body_twists(model, _advance_config(s, eps), nu), _body_twists(model, _advance_config(s, -eps), nu)
    M = np.zeros((18, 18)); rhs = np.zeros(18)
    rhs[6:18] = tau
    for b, (m, c, v, R, w, Ic) in enumerate(bodies):
        Iw = R @ Ic @ R.T
        dv = (tp[b][0] - tm[b][0]) / (2 * eps); dw = (tp[b][1] - tm[b][1]) / (2 * eps)      # d/dt(J) nu
        M += m * Jv[b].T @ Jv[b] + Jw[b].T @ Iw @ Jw[b]
        rhs -= Jv[b].T @ (m * (dv + np.array([0.0, 0.0, g]))) + Jw[b].T @ (Iw @ dw + np.cross(w, Iw @ w))
    return np.linalg.solve(M, rhs), M


def test_forward_dynamics_matches_dense(oracle_lib, blob, nominal_model):
    """All 18 accelerations (base linear / angular in the world frame, 12 joints) of the oracle's ABA against the dense
    formulation above, on random states with random joint torques; also with a perturbed base-link candidate."""
    rng = np.random.default_rng(23)
    for trial in range(6):
        s = _rand_state(rng, nominal_model, height=5.0)
        tau = rng.uniform(-15, 15, 12)
        acc, qdd, _ = oracle_lib.forward_dynamics(blob, s, tau, with_contact=False, with_gravity=True)
        lin, ang = _world_accel(s, acc)
        nud, M = dense_forward_dynamics(nominal_model, s, tau)
        assert np.linalg.eigvalsh(M).min() > 0
        np.testing.assert_allclose(np.concatenate([lin, ang, qdd]), nud, rtol=2e-6, atol=2e-5)
    # a candidate with another base mass / com / inertia (INERTIA_KEEP so that the test can state the inertia directly)
    import copy
    m2 = copy.deepcopy(nominal_model)
    m2.base = gm.Inertial(9.0, [0.05, -0.02, 0.03], [0.03, 0.12, 0.09, 0.0, 0.0, 0.0])
    names = ["mass", "comx", "comy", "comz", "inertiax", "inertiay", "inertiaz", "inertiaxy", "inertiaxz", "inertiayz"]
    p = [9.0, 0.05, -0.02, 0.03, 0.03, 0.12, 0.09, 0.0, 0.0, 0.0]
    s = _rand_state(rng, nominal_model, height=5.0); tau = rng.uniform(-15, 15, 12)
    acc, qdd, _ = oracle_lib.forward_dynamics(blob, s, tau, p, [gm.PARAM_IDS[n] for n in names], gm.FLAG_INERTIA_KEEP, False, True)
    lin, ang = _world_accel(s, acc)
    nud, _ = dense_forward_dynamics(m2, s, tau)
    np.testing.assert_allclose(np.concatenate([lin, ang, qdd]), nud, rtol=2e-6, atol=2e-5)


def test_power_balance_and_angular_momentum_rate(oracle_lib, blob, nominal_model):
    """dE/dt = tau . qd (no contact: the joint torques are the only non-conservative forces) and dL/dt about the world origin =
    sum r_com x m g.  Both depend on the joint accelerations, unlike dP/dt."""
    rng = np.random.default_rng(29)
    for _ in range(5):
        s = _rand_state(rng, nominal_model, height=5.0)
        tau = rng.uniform(-10, 10, 12)
        acc, qdd, _ = oracle_lib.forward_dynamics(blob, s, tau, with_contact=False, with_gravity=True)
        lin, ang = _world_accel(s, acc)

        def at(eps):
            t = _advance_config(s, eps)
            t[7:10] += eps * lin; t[10:13] += eps * ang; t[25:37] += eps * qdd
            return _momentum(nominal_model, t)
        eps = 1e-5
        (P1, L1, E1), (P0, L0, E0) = at(eps), at(-eps)
        np.testing.assert_allclose((E1 - E0) / (2 * eps), tau @ s[25:37], rtol=1e-5, atol=1e-4)
        torque = sum(np.cross(c, [0.0, 0.0, -9.81 * m]) for m, c, *_ in _bodies_world(nominal_model, s))
        np.testing.assert_allclose((L1 - L0) / (2 * eps), torque, rtol=1e-5, atol=1e-4)
        np.testing.assert_allclose((P1 - P0) / (2 * eps), [0, 0, -9.81 * nominal_model.total_mass()], atol=1e-4)


def test_static_stand_supports_weight(oracle_lib, blob, nominal_model):
    """Stand under the PD law for 2 s: the four foot forces sum to m g and the base settles."""
    m = nominal_model
    s = recorders.initial_state(m).astype(np.float64)[None]
    ff = None
    for k in range(100):
        q, qd = s[0, 13:25], s[0, 25:37]
        for d in range(4):
            tau = np.asarray(m.kp) * (np.asarray(m.q_default) - s[0, 13:25]) - np.asarray(m.kd) * s[0, 25:37]
            tau = np.clip(tau, -np.asarray(m.torque_limit), np.asarray(m.torque_limit))
            s, ff = oracle_lib.sim_step(blob, s, tau[None], 1, return_foot_force=True)
    assert abs(ff[0, :, 2].sum() - m.total_mass() * 9.81) / (m.total_mass() * 9.81) < 0.02
    assert np.abs(s[0, 7:13]).max() < 0.05 and 0.2 < s[0, 2] < 0.34


def test_cost_landscape_has_minimum_at_true_mass(oracle_lib, blob):
    """Sim-to-sim identifiability (the README experiment, README.md:153-161): data recorded at 6.921 kg ->
    the weighted landscape argmin lies within one grid cell of the truth."""
    import synth
    S, ds = synth.dataset("sine", 5)
    init, act, tgt, gains, mask, denom = synth.pack_numpy(ds)
    scales = np.linspace(0.5, 2.0, 20)
    params = (scales * 6.921).astype(np.float32)[:, None]
    cost, status = oracle_lib.eval_candidates(blob, params, [0], init, act, tgt, gains, mask, cost_denominator=denom)
    total = cost @ np.array([10.0, 5.0, 1.0])
    assert status.sum() == 0
    assert abs(scales[int(np.argmin(total))] - 1.0) <= (scales[1] - scales[0]) + 1e-9


def test_parameter_semantics(oracle_lib, blob, nominal_model):
    """mass_scale == mass / nominal; inertia scales with the mass unless INERTIA_KEEP; STRICT_INERTIAY drops
    inertiay (isaacgym_active_sysid.py:86 typo)."""
    rng = np.random.default_rng(5)
    s = _rand_state(rng, nominal_model, height=5.0)
    tau = rng.uniform(-5, 5, 12)
    fd = lambda p, ids, flags=0: np.concatenate(oracle_lib.forward_dynamics(blob, s, tau, p, ids, flags, False, True)[:2])
    np.testing.assert_allclose(fd([1.5], [gm.PARAM_IDS["mass_scale"]]), fd([1.5 * 6.921], [gm.PARAM_IDS["mass"]]), rtol=1e-6, atol=1e-6)
    a = fd([10.0], [0]); b = fd([10.0], [0], gm.FLAG_INERTIA_KEEP)
    assert np.abs(a - b).max() > 1e-3
    iy = [gm.PARAM_IDS["inertiay"]]
    assert np.abs(fd([0.2], iy) - fd([0.098077], iy)).max() > 1e-3
    np.testing.assert_allclose(fd([0.2], iy, gm.FLAG_STRICT_INERTIAY), fd([0.098077], iy), rtol=1e-6, atol=1e-6)
