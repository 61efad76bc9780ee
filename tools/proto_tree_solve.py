#!/usr/bin/env python
"""CPU prototype (numpy, float64) of the O(n) tree solve proposed in DESIGN.md §8 (4) for the next kernel round.

The constraint Hessian of one dynamics evaluation is  H = M + sum_b S_b^T W_b S_b + diag(d)  with W_b the 6x6
wrench-space Hessian of the active contact rows on body b and d the active joint-limit terms.  S_b^T W_b S_b is exactly
what a spatial inertia W_b attached to body b contributes to the joint-space inertia, so H is the mass matrix of the
same kinematic tree with body inertias I_b + W_b (+ armature / limit terms on the joint diagonals) and  H x = r  can be
solved by the articulated-body recursion (inward pass: articulated inertias and bias wrenches, outward pass:
accelerations) without forming M or H and without a dense factorisation.  All quantities are expressed in one common
frame (world axes, moments about one origin), so no spatial transforms appear between bodies - the same convention the
CUDA kernel uses.

This script checks, for both walkers at random configurations:
  1. the numpy kinematics / motion vectors / spatial inertias against the oracle's mass matrix (oracle/walker_physics.c);
  2. the tree solve against a dense solve of H for random positive semi-definite W_b on the foot bodies, random limit
     terms and a random right-hand side.
It needs no GPU.  Usage: python tools/proto_tree_solve.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from drloco_b200.model import get_model  # noqa: E402
from oracle.physics import OraclePhysics  # noqa: E402


def skew(c):
    return np.array([[0, -c[2], c[1]], [c[2], 0, -c[0]], [-c[1], c[0], 0]])


def rot(axis, ang):
    k = skew(axis)
    return np.eye(3) + np.sin(ang) * k + (1 - np.cos(ang)) * (k @ k)


def kinematics(m, q):
    """motion vectors S_j = (axis; anchor x axis) about the world origin and body spatial inertias about it."""
    nb, nv = m.nb, m.nv
    R, p = [None] * nb, [None] * nb
    S = np.zeros((nv, 6))
    dofs_of = [[j for j in range(nv) if m.dof_body[j] == b] for b in range(nb)]
    for b in range(nb):
        par = int(m.body_parent[b])
        Rp, pp = (np.eye(3), np.zeros(3)) if par < 0 else (R[par], p[par])
        Rb, pb = Rp.copy(), pp + Rp @ m.body_pos[b]
        for j in dofs_of[b]:
            e = np.zeros(3)
            e[int(m.dof_axis_idx[j])] = float(m.dof_axis_sign[j])
            a = Rb @ e
            dq = q[j] - m.dof_ref[j]
            if m.dof_type[j] == 0:                       # slide
                S[j] = np.concatenate([np.zeros(3), a])
                pb = pb + a * dq
            else:                                        # hinge through the body origin
                S[j] = np.concatenate([a, np.cross(pb, a)])
                Rb = Rb @ rot(e, dq)
        R[b], p[b] = Rb, pb
    inertia = []
    for b in range(nb):
        c = p[b] + R[b] @ m.body_ipos[b]
        mass = float(m.body_mass[b])
        Ic = R[b] @ np.diag(m.body_inertia[b]) @ R[b].T
        C = skew(c)
        inertia.append(np.block([[Ic + mass * C @ C.T, mass * C], [mass * C.T, mass * np.eye(3)]]))
    supp = np.zeros((nb, nv), bool)                      # dofs that move body b
    for b in range(nb):
        x = b
        while x >= 0:
            supp[b, dofs_of[x]] = True
            x = int(m.body_parent[x])
    return S, inertia, supp, dofs_of


def dense_hessian(S, inertia, supp, diag):
    nv = S.shape[0]
    H = np.diag(np.asarray(diag, float))
    for b, Ib in enumerate(inertia):
        Sb = S * supp[b][:, None]
        H += Sb @ Ib @ Sb.T
    return H


def tree_solve(m, S, inertia, dofs_of, diag, r):
    """x with H x = r, H as in dense_hessian, by the articulated-body recursion (1-DoF joints, common frame)."""
    nv = S.shape[0]
    parent = np.full(nv, -1)
    for j in range(nv):
        b = int(m.dof_body[j])
        k = dofs_of[b].index(j)
        if k > 0:
            parent[j] = dofs_of[b][k - 1]
        else:
            x = int(m.body_parent[b])
            while x >= 0 and not dofs_of[x]:
                x = int(m.body_parent[x])
            parent[j] = dofs_of[x][-1] if x >= 0 else -1
    IA = np.zeros((nv, 6, 6))
    pA = np.zeros((nv, 6))
    for b, Ib in enumerate(inertia):
        IA[dofs_of[b][-1]] += Ib                          # a body's inertia sits on its last dof-link
    U, D, u = np.zeros((nv, 6)), np.zeros(nv), np.zeros(nv)
    for j in range(nv - 1, -1, -1):                       # inward
        U[j] = IA[j] @ S[j]
        D[j] = S[j] @ U[j] + diag[j]
        u[j] = r[j] - S[j] @ pA[j]
        if parent[j] >= 0:
            IA[parent[j]] += IA[j] - np.outer(U[j], U[j]) / D[j]
            pA[parent[j]] += pA[j] + U[j] * (u[j] / D[j])
    acc = np.zeros((nv, 6))
    x = np.zeros(nv)
    for j in range(nv):                                   # outward
        ap = acc[parent[j]] if parent[j] >= 0 else np.zeros(6)
        x[j] = (u[j] - U[j] @ ap) / D[j]
        acc[j] = ap + S[j] * x[j]
    return x


def main():
    rng = np.random.default_rng(0)
    worst_M, worst_x = 0.0, 0.0
    for env_id in ("StraightMimicWalker", "MimicWalker165cm65kg"):
        m = get_model(env_id)
        P = OraclePhysics(m)
        feet = sorted(set(int(b) for b in m.box_body))
        for trial in range(20):
            q = m.qpos0 + 0.4 * rng.standard_normal(m.nv)
            S, inertia, supp, dofs_of = kinematics(m, q)
            M = dense_hessian(S, inertia, supp, m.dof_armature)
            Mo = P.mass_matrix(q)
            worst_M = max(worst_M, float(np.abs(M - Mo).max() / np.abs(Mo).max()))
            aug = [I.copy() for I in inertia]
            for b in feet:                                # random PSD wrench-space Hessians, stiff like contacts
                A = rng.standard_normal((6, int(rng.integers(1, 9))))
                aug[b] = aug[b] + 1e3 * A @ A.T
            diag = m.dof_armature + np.where(rng.random(m.nv) < 0.2, 1e3 * rng.random(m.nv), 0.0)
            r = 100 * rng.standard_normal(m.nv)
            H = dense_hessian(S, aug, supp, diag)
            x_ref = np.linalg.solve(H, r)
            x = tree_solve(m, S, aug, dofs_of, diag, r)
            worst_x = max(worst_x, float(np.abs(x - x_ref).max() / np.abs(x_ref).max()))
        print(f"{env_id}: nv={m.nv} feet bodies {feet}")
    print(f"mass matrix vs oracle: max rel err {worst_M:.2e};  tree solve vs dense solve: max rel err {worst_x:.2e}")
    ok = worst_M < 1e-10 and worst_x < 1e-8
    print("OK" if ok else "MISMATCH")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
