"""ctypes front-end of oracle/liboracle.so (see walker_physics.c for what it restates).  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from drloco_b200.cabi import DrlWalkerModel, pack_model, INTEGRATOR_RK4, INTEGRATOR_EULER  # noqa: F401

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
MAXCON = 48


class OrcDiag(C.Structure):
    _fields_ = [("ncon", C.c_int), ("nefc", C.c_int), ("nlimit", C.c_int), ("iters", C.c_int),
                ("kkt_residual", C.c_double), ("normal_force", C.c_double),
                ("con_pos", C.c_double * 3 * MAXCON), ("con_dist", C.c_double * MAXCON), ("con_body", C.c_int * MAXCON),
                ("con_force", C.c_double * 3 * MAXCON), ("energy_kin", C.c_double), ("energy_pot", C.c_double)]


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "walker_physics.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        assert _LIB.orc_diag_size() == C.sizeof(OrcDiag), "OrcDiag layout mismatch"
        assert _LIB.orc_max_con() == MAXCON
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class OraclePhysics:
    """Float64 walker physics (MuJoCo-pipeline restatement) for one environment."""

    def __init__(self, model, integrator=INTEGRATOR_RK4):
        self.model = model
        self.cm = pack_model(model)
        self.nv, self.nu = model.nv, model.nu
        self.integrator = integrator
        self.qacc_warm = np.zeros(self.nv)
        self._l = lib()

    def forward(self, q, v, ctrl, warm=None, want_diag=True):
        q, v = np.ascontiguousarray(q, np.float64), np.ascontiguousarray(v, np.float64)
        ctrl = np.ascontiguousarray(ctrl, np.float64)
        qacc = np.zeros(self.nv)
        diag = OrcDiag()
        w = None if warm is None else np.ascontiguousarray(warm, np.float64).copy()
        rc = self._l.orc_forward(C.byref(self.cm), _p(q), _p(v), _p(ctrl), None if w is None else _p(w), _p(qacc),
                                 C.byref(diag) if want_diag else None)
        if rc:
            raise FloatingPointError("oracle forward dynamics failed")
        return qacc, diag

    def step(self, q, v, ctrl, nsub):
        """in place; returns True when MuJoCo would have raised (non-finite / huge state)."""
        assert q.dtype == np.float64 and v.dtype == np.float64 and q.flags.c_contiguous and v.flags.c_contiguous
        ctrl = np.ascontiguousarray(ctrl, np.float64)
        return bool(self._l.orc_step(C.byref(self.cm), _p(q), _p(v), _p(ctrl), _p(self.qacc_warm), int(nsub),
                                     int(self.integrator)))

    def step_trace(self, q, v, ctrl, nsub):
        """like step() (in place) but also returns the constraint-set signature of every dynamics evaluation:
        uint64 [nsub, 4, 2] = (contact candidates, limit rows) per substep and RK4 stage."""
        assert q.dtype == np.float64 and v.dtype == np.float64 and q.flags.c_contiguous and v.flags.c_contiguous
        ctrl = np.ascontiguousarray(ctrl, np.float64)
        sig = np.zeros((int(nsub), 4, 2), np.uint64)
        bad = bool(self._l.orc_step_trace(C.byref(self.cm), _p(q), _p(v), _p(ctrl), _p(self.qacc_warm), int(nsub),
                                          int(self.integrator), _p(sig)))
        return bad, sig

    def site_xpos(self, q):
        out = np.zeros((len(self.model.site_body), 3))
        self._l.orc_site_xpos(C.byref(self.cm), _p(np.ascontiguousarray(q, np.float64)), _p(out))
        return out

    def mass_matrix(self, q):
        M = np.zeros((self.nv, self.nv))
        self._l.orc_mass_matrix(C.byref(self.cm), _p(np.ascontiguousarray(q, np.float64)), _p(M))
        return M

    def bias(self, q, v):
        c = np.zeros(self.nv)
        self._l.orc_bias(C.byref(self.cm), _p(np.ascontiguousarray(q, np.float64)),
                         _p(np.ascontiguousarray(v, np.float64)), _p(c))
        return c

    def body_com(self, q):
        com = np.zeros((self.model.nb, 3))
        xmat = np.zeros((self.model.nb, 9))
        self._l.orc_body_com(C.byref(self.cm), _p(np.ascontiguousarray(q, np.float64)), _p(com), _p(xmat))
        return com, xmat.reshape(-1, 3, 3)
