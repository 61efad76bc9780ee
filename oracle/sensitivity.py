"""oracle/sensitivity.py — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Contact-sensitivity probe for the fp32-vs-float64 parity tests.

The soft-constraint forces of the walker's physics (oracle/walker_physics.c, restating MuJoCo's documented pipeline) are
discontinuous where a constraint row is instantiated: the damping part ``-b * v`` of the reference acceleration switches
on at full size the moment a foot corner reaches the ground (or a joint its limit).  Two trajectories that agree to
float32 resolution therefore separate visibly iff such an event falls between them - one sees the contact one RK4 stage
or substep before the other.  Whether that can happen is a property of the float64 trajectory alone, and the oracle
decides it: ``SensitivityProbe`` wraps the oracle physics of one environment and, at every control step, re-runs the step
from copies of the state perturbed at float32 resolution; the environment is *flagged* as soon as a perturbed copy
instantiates a different constraint set in any dynamics evaluation (substep x RK4 stage) of that step.

The parity tests then require EVERY non-flagged environment to stay within the stated tolerance and report the flagged
fraction (tests/test_gpu_parity.py, __graft_entry__.smoke()).
"""
from __future__ import annotations

import numpy as np

from oracle.physics import OraclePhysics

# Perturbation sizes: what an fp32 trajectory accumulates within a few control steps (measured, DESIGN.md section 5:
# |dq| 3e-7 -> 2e-6, |dv| 5e-5 -> 1.7e-4 absolute over 10 control steps).  ~16 ulp(f32) of q, ~1e-5 relative in v.
EPS_Q = 2e-6
EPS_V = 2e-5
N_DRAWS = 6


class SensitivityProbe:
    """physics object for OracleMimicEnv (same interface as OraclePhysics) that also tracks contact sensitivity."""

    def __init__(self, model, integrator, seed=0, eps_q=EPS_Q, eps_v=EPS_V, n_draws=N_DRAWS):
        self._p = OraclePhysics(model, integrator)
        self._probe = OraclePhysics(model, integrator)
        self.qacc_warm = self._p.qacc_warm
        self.rng = np.random.default_rng(seed)
        self.eps_q, self.eps_v, self.n_draws = eps_q, eps_v, n_draws
        self.flagged = False          # sticky until clear(): once separated, the trajectories stay separated
        self.events = 0               # control steps in which a perturbed copy changed the constraint set

    def clear(self):
        self.flagged = False

    def step(self, q, v, ctrl, nsub):
        warm0 = self._p.qacc_warm.copy()
        qb, vb = q.copy(), v.copy()
        bad, base = self._p.step_trace(q, v, ctrl, nsub)           # the real step (in place)
        if bad:
            return True
        sq, sv = np.maximum(1.0, np.abs(qb)), np.maximum(1.0, np.abs(vb))
        for _ in range(self.n_draws):
            qp = qb + self.eps_q * sq * self.rng.uniform(-1, 1, qb.shape)
            vp = vb + self.eps_v * sv * self.rng.uniform(-1, 1, vb.shape)
            self._probe.qacc_warm[:] = warm0
            badp, sig = self._probe.step_trace(qp, vp, ctrl, nsub)
            if badp or not np.array_equal(sig, base):
                self.flagged = True
                self.events += 1
                break
        return False

    def site_xpos(self, q):
        return self._p.site_xpos(q)
