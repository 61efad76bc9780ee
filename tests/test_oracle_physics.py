"""First-principles checks that pin the float64 physics oracle (oracle/walker_physics.c).  MuJoCo itself cannot run
here, so the restatement is pinned to mechanics instead: the mass matrix to the kinetic-energy
definition, the bias force to Lagrange's equations, the integrator to energy conservation, the constraint solve to its
KKT conditions and to weight = normal force at rest."""
import copy

import numpy as np
import pytest

from drloco_b200 import model as M
from oracle.physics import OraclePhysics, lib, INTEGRATOR_EULER


@pytest.fixture(scope="module", params=["StraightMimicWalker", "MimicWalker165cm65kg"])
def phys(request):
    m = M.get_model(request.param)
    return m, OraclePhysics(m)


def _rand_state(m, rng, z=3.0):
    q = m.qpos0 + 0.3 * rng.standard_normal(m.nv)
    q[2] = z
    return q, rng.standard_normal(m.nv)


def test_mass_matrix_is_kinetic_energy(phys):
    """v' M v / 2 == sum_b (m |v_com|^2 + w' I w) / 2 with body velocities from finite differences of the kinematics."""
    m, P = phys
    rng = np.random.default_rng(0)
    q, v = _rand_state(m, rng)
    Mq = P.mass_matrix(q)
    np.testing.assert_allclose(Mq, M.mass_matrix(m, q), rtol=0, atol=1e-12)       # independent numpy formulation
    eps = 1e-6
    c1, R1 = P.body_com(q + eps * v)
    c0, R0 = P.body_com(q - eps * v)
    vcom = (c1 - c0) / (2 * eps)
    T = 0.0
    for b in range(m.nb):
        Rdot = (R1[b] - R0[b]) / (2 * eps)
        Rm = 0.5 * (R1[b] + R0[b])
        Wx = Rdot @ Rm.T                                                      # [w]x
        w = np.array([Wx[2, 1], Wx[0, 2], Wx[1, 0]])
        Iw = Rm @ np.diag(m.body_inertia[b]) @ Rm.T
        T += 0.5 * m.body_mass[b] * vcom[b] @ vcom[b] + 0.5 * w @ Iw @ w
    T += 0.5 * np.sum(m.dof_armature * v * v)
    assert abs(0.5 * v @ Mq @ v - T) < 1e-6 * T


def test_bias_force_satisfies_lagrange(phys):
    """c = Mdot v - 1/2 d(v'Mv)/dq + dU/dq."""
    m, P = phys
    rng = np.random.default_rng(1)
    q, v = _rand_state(m, rng)
    eps = 1e-6

    def U(qq):
        com, _ = P.body_com(qq)
        return float((m.body_mass * 9.81 * com[:, 2]).sum())
    c = P.bias(q, v)
    cl = ((P.mass_matrix(q + eps * v) - P.mass_matrix(q - eps * v)) / (2 * eps)) @ v
    for j in range(m.nv):
        e = np.zeros(m.nv)
        e[j] = eps
        cl[j] -= 0.5 * (v @ P.mass_matrix(q + e) @ v - v @ P.mass_matrix(q - e) @ v) / (2 * eps)
        cl[j] += (U(q + e) - U(q - e)) / (2 * eps)
    assert np.abs(c - cl).max() < 1e-6 * max(1.0, np.abs(c).max())


def test_free_fall_and_energy_conservation(phys):
    m, _ = phys
    m2 = copy.deepcopy(m)
    m2.dof_damping[:] = 0
    m2.dof_limited[:] = 0
    P = OraclePhysics(m2)
    q = m.qpos0.copy()
    q[2] = 5.0
    a, d = P.forward(q, np.zeros(m.nv), np.zeros(m.nu))
    up = 2                                                                 # z slide
    assert d.ncon == 0 and d.nefc == 0
    assert abs(a[up] + 9.81) < 1e-12 and np.abs(np.delete(a, up)).max() < 1e-9
    rng = np.random.default_rng(2)
    q, v = _rand_state(m, rng, z=6.0)
    _, d0 = P.forward(q, v, np.zeros(m.nu))
    qq, vv = q.copy(), v.copy()
    assert not P.step(qq, vv, np.zeros(m.nu), 200)                         # 0.2 s of RK4 at 1 ms
    _, d1 = P.forward(qq, vv, np.zeros(m.nu))
    E0, E1 = d0.energy_kin + d0.energy_pot, d1.energy_kin + d1.energy_pot
    assert abs(E1 - E0) < 1e-8 * abs(E0)                                   # O(dt^4) drift


def test_resting_contact_carries_the_weight():
    m = M.get_model("StraightMimicWalker")
    P = OraclePhysics(m)
    q, v = m.qpos0.copy(), np.zeros(m.nv)
    for _ in range(200):                                                   # 1 s standing on locked-out knees
        assert not P.step(q, v, np.zeros(m.nu), 5)
    a, d = P.forward(q, v, np.zeros(m.nu), warm=P.qacc_warm)
    assert d.ncon == 8 and d.nefc >= 32                                    # 2 feet x 4 corners x 4 pyramid rows
    assert abs(d.normal_force - m.total_mass * 9.81) < 0.05               # 789.7 N
    assert d.kkt_residual < 1e-8
    assert q[2] > 1.07 and np.abs(v).max() < 1e-2


def test_constraint_solution_is_independent_of_solver_and_warmstart():
    m = M.get_model("StraightMimicWalker")
    P = OraclePhysics(m)
    rng = np.random.default_rng(3)
    q = m.qpos0.copy()
    q[2] -= 0.003
    q[6:] += 0.05 * rng.standard_normal(8)
    q[8] = -0.02                                                           # knee past its lower limit
    v = 0.5 * rng.standard_normal(m.nv)
    ctrl = 100 * rng.standard_normal(m.nu)
    a0, d0 = P.forward(q, v, ctrl)
    assert d0.ncon > 0 and d0.nlimit >= 1 and d0.kkt_residual < 1e-8
    a1, _ = P.forward(q, v, ctrl, warm=100 * rng.standard_normal(m.nv))
    np.testing.assert_allclose(a1, a0, rtol=1e-7, atol=1e-7)
    lib().orc_set_solver(1, 50)                                            # active-set iteration (what the GPU runs)
    try:
        a2, _ = P.forward(q, v, ctrl)
    finally:
        lib().orc_set_solver(0, 50)
    np.testing.assert_allclose(a2, a0, rtol=1e-7, atol=1e-7)


def test_euler_and_rk4_agree_to_first_order():
    m = M.get_model("StraightMimicWalker")
    rng = np.random.default_rng(4)
    q0, v0 = _rand_state(m, rng, z=4.0)
    outs = []
    for integ in (0, INTEGRATOR_EULER):
        P = OraclePhysics(m, integ)
        q, v = q0.copy(), v0.copy()
        P.step(q, v, np.zeros(m.nu), 5)
        outs.append((q, v))
    assert np.abs(outs[0][0] - outs[1][0]).max() < 5e-3 and np.abs(outs[0][1] - outs[1][1]).max() < 0.5


def test_blowup_is_reported():
    m = M.get_model("StraightMimicWalker")
    P = OraclePhysics(m)
    q, v = m.qpos0.copy(), np.zeros(m.nv)
    v[0] = 1e11
    assert P.step(q, v, np.zeros(m.nu), 1)


def test_tree_solve_prototype_matches_dense_solve():
    """tools/proto_tree_solve.py: H = M + sum S_b^T W_b S_b + diag solved by the articulated-body
    recursion equals the dense solve; its numpy kinematics reproduce the oracle's mass matrix."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    p = subprocess.run([sys.executable, os.path.join(root, "tools", "proto_tree_solve.py")], capture_output=True,
                       text=True, timeout=300)
    assert p.returncode == 0 and p.stdout.strip().endswith("OK"), p.stdout + p.stderr
