/* oracle/walker_physics.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU float64 restatement of the physics the reference obtains from the MuJoCo C library through
 * gym.MujocoEnv.do_simulation -> mujoco_py.MjSim.step -> mj_step (reference call sites drloco/mujoco/mimic_env.py:52,83,
 * 539,549; model files drloco/mujoco/xml/walker3d_flat_feet.xml and walker_165cm_65kg.xml).
 *
 * MuJoCo itself (and mujoco-py, gym 0.18.0) is a third-party dependency that is NOT under /root/reference and is not
 * installed in the build image, and the reference repository holds no test or golden vector for this boundary
 * (SURVEY.md §8c).  PARITY UNPINNED: this file restates MuJoCo's *documented* pipeline ("Computation" chapter: soft
 * constraints, pyramidal cones, RK4) for exactly the options those XML files select; it cannot be validated against a
 * MuJoCo binary here.  It is pinned instead by first-principles checks in tests/ (kinetic-energy definition of M,
 * Lagrange identity for the bias forces, energy conservation, free fall, weight = total normal force at rest, KKT
 * residual of the constraint solve).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this library.
 *
 * Pipeline per dynamics evaluation (mj_forward):
 *   kinematics -> mass matrix (+armature) -> bias (Coriolis/centrifugal/gravity) -> passive damping -> motor forces
 *   -> plane/box and plane/capsule collision -> joint-limit and pyramidal contact rows with reference acceleration
 *   (solref) and regulariser (solimp, invweight0) -> convex primal solve for qacc (Newton, exact line search).
 * Integration: RK4 (classic tableau, 4 evaluations per timestep) as the XML requests, or MuJoCo's semi-implicit Euler
 * with implicit joint damping.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "../include/drloco_b200.h"

#define NV DRL_MAX_DOF
#define NB DRL_MAX_BODY
#define MAXCON (DRL_MAX_SPHERE + 4 * DRL_MAX_BOX)
#define MAXROW (4 * MAXCON + DRL_MAX_DOF)
#define MJ_MINVAL 1e-15
#define MJ_MINIMP 0.0001
#define MJ_MAXIMP 0.9999

typedef struct {
  int ncon, nefc, nlimit, iters;
  double kkt_residual;       /* || M(a - a0) - J^T f ||_inf */
  double normal_force;       /* sum over contacts of the normal force */
  double con_pos[MAXCON][3];
  double con_dist[MAXCON];
  int con_body[MAXCON];
  double con_force[MAXCON][3];  /* normal, tangent1(y), tangent2(x) components of each contact force */
  double energy_kin, energy_pot;
} OrcDiag;

typedef struct {
  double xpos[NB][3], xmat[NB][9];
  double axis[NV][3], anchor[NV][3];
  double jacp[NB][3][NV], jacr[NB][3][NV]; /* COM jacobians */
  double com[NB][3];
} Kin;

static void cross3(double* r, const double* a, const double* b) {
  double x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  r[0] = x; r[1] = y; r[2] = z;
}
static double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static void matvec3(double* r, const double* R, const double* v) {
  double x = R[0] * v[0] + R[1] * v[1] + R[2] * v[2];
  double y = R[3] * v[0] + R[4] * v[1] + R[5] * v[2];
  double z = R[6] * v[0] + R[7] * v[1] + R[8] * v[2];
  r[0] = x; r[1] = y; r[2] = z;
}

/* does dof j move body b?  (dof's body is b or an ancestor of b) */
static int supports(const DrlWalkerModel* m, int j, int b) {
  int jb = m->dof_body[j];
  while (b >= 0) {
    if (b == jb) return 1;
    b = m->body_parent[b];
  }
  return 0;
}

/* mj_kinematics for slide/hinge joints with coordinate axes anchored at the body origin. */
static void kinematics(const DrlWalkerModel* m, const double* q, Kin* k) {
  for (int b = 0; b < m->nb; b++) {
    int p = m->body_parent[b];
    double pos[3], R[9];
    if (p < 0) {
      memcpy(pos, m->body_pos[b], sizeof pos);
      memset(R, 0, sizeof R); R[0] = R[4] = R[8] = 1.0;
    } else {
      matvec3(pos, k->xmat[p], m->body_pos[b]);
      for (int i = 0; i < 3; i++) pos[i] += k->xpos[p][i];
      memcpy(R, k->xmat[p], sizeof R);
    }
    for (int j = 0; j < m->nv; j++) {
      if (m->dof_body[j] != b) continue;
      int a = m->dof_axis_idx[j];
      double sg = m->dof_axis_sign[j], d = q[j] - m->dof_ref[j];
      for (int i = 0; i < 3; i++) { k->axis[j][i] = sg * R[3 * i + a]; k->anchor[j][i] = pos[i]; }
      if (m->dof_type[j] == 0) {
        for (int i = 0; i < 3; i++) pos[i] += k->axis[j][i] * d;
      } else { /* R <- R * Rot(e_a, sg*d): mixes columns a+1, a+2 */
        int c1 = (a + 1) % 3, c2 = (a + 2) % 3;
        double c = cos(sg * d), s = sin(sg * d);
        for (int i = 0; i < 3; i++) {
          double u = R[3 * i + c1], w = R[3 * i + c2];
          R[3 * i + c1] = c * u + s * w;
          R[3 * i + c2] = -s * u + c * w;
        }
      }
    }
    memcpy(k->xpos[b], pos, sizeof pos);
    memcpy(k->xmat[b], R, sizeof R);
    double rc[3];
    matvec3(rc, R, m->body_ipos[b]);
    for (int i = 0; i < 3; i++) k->com[b][i] = pos[i] + rc[i];
  }
  /* COM jacobians */
  for (int b = 0; b < m->nb; b++)
    for (int j = 0; j < m->nv; j++) {
      double jp[3] = {0, 0, 0}, jr[3] = {0, 0, 0};
      if (supports(m, j, b)) {
        if (m->dof_type[j] == 0) {
          memcpy(jp, k->axis[j], sizeof jp);
        } else {
          double r[3] = {k->com[b][0] - k->anchor[j][0], k->com[b][1] - k->anchor[j][1], k->com[b][2] - k->anchor[j][2]};
          memcpy(jr, k->axis[j], sizeof jr);
          cross3(jp, k->axis[j], r);
        }
      }
      for (int i = 0; i < 3; i++) { k->jacp[b][i][j] = jp[i]; k->jacr[b][i][j] = jr[i]; }
    }
}

/* world-frame inertia of body b about its COM */
static void world_inertia(const DrlWalkerModel* m, const Kin* k, int b, double* Iw) {
  const double* R = k->xmat[b];
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) {
      double s = 0;
      for (int a = 0; a < 3; a++) s += R[3 * r + a] * m->body_inertia[b][a] * R[3 * c + a];
      Iw[3 * r + c] = s;
    }
}

static void mass_matrix(const DrlWalkerModel* m, const Kin* k, double* M /* nv*nv */) {
  int nv = m->nv;
  memset(M, 0, sizeof(double) * nv * nv);
  for (int b = 0; b < m->nb; b++) {
    double Iw[9];
    world_inertia(m, k, b, Iw);
    for (int r = 0; r < nv; r++)
      for (int c = 0; c < nv; c++) {
        double s = 0;
        for (int i = 0; i < 3; i++) s += m->body_mass[b] * k->jacp[b][i][r] * k->jacp[b][i][c];
        for (int i = 0; i < 3; i++)
          for (int j = 0; j < 3; j++) s += k->jacr[b][i][r] * Iw[3 * i + j] * k->jacr[b][j][c];
        M[r * nv + c] += s;
      }
  }
  for (int j = 0; j < nv; j++) M[j * nv + j] += m->dof_armature[j];
}

/* Bias force c(q,v): Newton-Euler recursion in world coordinates with qacc = 0, gravity included. */
static void bias_force(const DrlWalkerModel* m, const Kin* k, const double* q, const double* v, double* c, double* e_kin) {
  double w[NB][3], al[NB][3], vo[NB][3], ao[NB][3];
  double ek = 0;
  memset(c, 0, sizeof(double) * m->nv);
  for (int b = 0; b < m->nb; b++) {
    int p = m->body_parent[b];
    double W[3] = {0, 0, 0}, A[3] = {0, 0, 0}, V[3] = {0, 0, 0}, Ac[3] = {0, 0, 0};
    if (p >= 0) {
      double r[3], t[3], t2[3];
      matvec3(r, k->xmat[p], m->body_pos[b]);
      memcpy(W, w[p], sizeof W); memcpy(A, al[p], sizeof A);
      cross3(t, W, r);
      for (int i = 0; i < 3; i++) V[i] = vo[p][i] + t[i];
      cross3(t2, W, t);
      cross3(t, A, r);
      for (int i = 0; i < 3; i++) Ac[i] = ao[p][i] + t[i] + t2[i];
    }
    for (int j = 0; j < m->nv; j++) {
      if (m->dof_body[j] != b) continue;
      const double* ax = k->axis[j];
      if (m->dof_type[j] == 0) {
        /* the origin moves by d along an axis carried by the frame rotating with (W, A):
           v += W x (ax d) + ax qd ;  a += A x (ax d) + W x (W x ax d) + 2 W x ax qd */
        double d = q[j] - m->dof_ref[j];
        double ad[3] = {ax[0] * d, ax[1] * d, ax[2] * d}, av[3] = {ax[0] * v[j], ax[1] * v[j], ax[2] * v[j]};
        double t[3], t2[3], t3[3], t4[3];
        cross3(t, W, ad);
        cross3(t2, W, t);
        cross3(t3, A, ad);
        cross3(t4, W, av);
        for (int i = 0; i < 3; i++) { V[i] += t[i] + av[i]; Ac[i] += t3[i] + t2[i] + 2.0 * t4[i]; }
      } else {
        double t[3], av[3] = {ax[0] * v[j], ax[1] * v[j], ax[2] * v[j]};
        cross3(t, W, av);
        for (int i = 0; i < 3; i++) { W[i] += av[i]; A[i] += t[i]; }
      }
    }
    memcpy(w[b], W, sizeof W); memcpy(al[b], A, sizeof A); memcpy(vo[b], V, sizeof V); memcpy(ao[b], Ac, sizeof Ac);
    /* COM kinematics */
    double rc[3] = {k->com[b][0] - k->xpos[b][0], k->com[b][1] - k->xpos[b][1], k->com[b][2] - k->xpos[b][2]};
    double t[3], t2[3], vc[3], ac[3];
    cross3(t, W, rc);
    for (int i = 0; i < 3; i++) vc[i] = V[i] + t[i];
    cross3(t2, W, t);
    cross3(t, A, rc);
    for (int i = 0; i < 3; i++) ac[i] = Ac[i] + t[i] + t2[i];
    double Iw[9], Iwv[3], Ial[3], f[3], n[3];
    world_inertia(m, k, b, Iw);
    matvec3(Iwv, Iw, W);
    matvec3(Ial, Iw, A);
    cross3(t, W, Iwv);
    for (int i = 0; i < 3; i++) { f[i] = m->body_mass[b] * ac[i]; n[i] = Ial[i] + t[i]; }
    f[2] -= m->body_mass[b] * m->gravity_z; /* m (a - g) */
    ek += 0.5 * m->body_mass[b] * dot3(vc, vc) + 0.5 * dot3(W, Iwv);
    for (int j = 0; j < m->nv; j++) {
      double s = 0;
      for (int i = 0; i < 3; i++) s += k->jacp[b][i][j] * f[i] + k->jacr[b][i][j] * n[i];
      c[j] += s;
    }
  }
  for (int j = 0; j < m->nv; j++) ek += 0.5 * m->dof_armature[j] * v[j] * v[j];
  if (e_kin) *e_kin = ek;
}

/* dense Cholesky solve of the SPD system A x = b (n <= NV); A is overwritten */
static int chol_solve(double* A, double* b, int n) {
  for (int j = 0; j < n; j++) {
    double d = A[j * n + j];
    for (int k2 = 0; k2 < j; k2++) d -= A[j * n + k2] * A[j * n + k2];
    if (!(d > 0)) return -1;
    d = sqrt(d);
    A[j * n + j] = d;
    for (int i = j + 1; i < n; i++) {
      double s = A[i * n + j];
      for (int k2 = 0; k2 < j; k2++) s -= A[i * n + k2] * A[j * n + k2];
      A[i * n + j] = s / d;
    }
  }
  for (int i = 0; i < n; i++) {
    double s = b[i];
    for (int k2 = 0; k2 < i; k2++) s -= A[i * n + k2] * b[k2];
    b[i] = s / A[i * n + i];
  }
  for (int i = n - 1; i >= 0; i--) {
    double s = b[i];
    for (int k2 = i + 1; k2 < n; k2++) s -= A[k2 * n + i] * b[k2];
    b[i] = s / A[i * n + i];
  }
  return 0;
}

/* solimp -> impedance d(r) (MuJoCo getimpedance) */
static double impedance(const double* solimp, double pos) {
  double d0 = fmin(MJ_MAXIMP, fmax(MJ_MINIMP, solimp[0]));
  double dm = fmin(MJ_MAXIMP, fmax(MJ_MINIMP, solimp[1]));
  double width = fmax(MJ_MINVAL, solimp[2]);
  double mid = fmin(MJ_MAXIMP, fmax(MJ_MINIMP, solimp[3]));
  double power = fmax(1.0, solimp[4]);
  if (d0 == dm || width <= MJ_MINVAL) return 0.5 * (d0 + dm);
  double x = fabs(pos) / width, y;
  if (x >= 1) return dm;
  if (x <= 0) return d0;
  if (power == 1) y = x;
  else if (x <= mid) y = pow(x, power) / pow(mid, power - 1);
  else y = 1 - pow(1 - x, power) / pow(1 - mid, power - 1);
  return d0 + y * (dm - d0);
}

int g_solver_mode = 0, g_as_maxit = 50;
long g_as_hist[64], g_as_fail = 0;
void orc_set_solver(int mode, int maxit) { g_solver_mode = mode; g_as_maxit = maxit; memset(g_as_hist, 0, sizeof g_as_hist); g_as_fail = 0; }
void orc_get_solver_stats(long* hist, long* fail) { memcpy(hist, g_as_hist, sizeof g_as_hist); *fail = g_as_fail; }

/* which constraints the last orc_forward call instantiated: bit s for capsule end sphere s, bit n_sphere + 8 x + i for
 * kept corner i of box x; limit rows: bit j (lower bound of dof j violated) / bit 32 + j (upper).  Read by the
 * contact-sensitivity probe of the parity tests (oracle/sensitivity.py). */
static unsigned long long g_last_con = 0, g_last_lim = 0;
void orc_get_last_sets(unsigned long long* con, unsigned long long* lim) { *con = g_last_con; *lim = g_last_lim; }

typedef struct { double a; int i; } Brk;
static int brk_cmp(const void* x, const void* y) {
  double a = ((const Brk*)x)->a, b = ((const Brk*)y)->a;
  return (a > b) - (a < b);
}

/* One forward-dynamics evaluation.  qacc_warm (in/out, nullable) is the solver's starting point; the converged qacc
 * does not depend on it.  Returns 0 or -1 on numerical failure. */
int orc_forward(const DrlWalkerModel* m, const double* q, const double* v, const double* ctrl, double* qacc_warm,
                double* qacc, OrcDiag* diag) {
  int nv = m->nv;
  Kin kin, *k = &kin;
  double M[NV * NV], c[NV], tau[NV], a0[NV], e_kin = 0;
  kinematics(m, q, k);
  mass_matrix(m, k, M);
  bias_force(m, k, q, v, c, &e_kin);
  /* passive + actuation */
  for (int j = 0; j < nv; j++) tau[j] = -m->dof_damping[j] * v[j] - c[j];
  for (int u = 0; u < m->nu; u++) {
    double cc = fmin(m->act_ctrlrange[u][1], fmax(m->act_ctrlrange[u][0], ctrl[u]));
    double f = fmin(m->act_forcerange[u][1], fmax(m->act_forcerange[u][0], cc));
    tau[m->act_dof[u]] += m->act_gear[u] * f;
  }
  { /* a0 = M^-1 tau */
    double Mc[NV * NV];
    memcpy(Mc, M, sizeof(double) * nv * nv);
    memcpy(a0, tau, sizeof(double) * nv);
    if (chol_solve(Mc, a0, nv)) return -1;
  }

  /* ---- constraint rows ---- */
  static const double dirs[4][2] = {{1, 0}, {-1, 0}, {0, 1}, {0, -1}}; /* coefficients of (mu*Jx, mu*Jy) */
  double J[MAXROW][NV];
  double D[MAXROW], aref[MAXROW];
  int row_con[MAXROW];
  int nrow = 0, ncon = 0, nlimit = 0;
  double tc = fmax(m->solref[0], 2 * m->timestep), dr = m->solref[1];
  double dmax = fmin(MJ_MAXIMP, fmax(MJ_MINIMP, m->solimp[1]));
  double Kc = 1.0 / (dmax * dmax * tc * tc * dr * dr), Bc = 2.0 / (dmax * tc);
  unsigned long long con_set = 0, lim_set = 0;
  /* joint limits (mj_instantiateLimit) */
  for (int j = 0; j < nv; j++) {
    if (!m->dof_limited[j]) continue;
    for (int side = 0; side < 2; side++) {
      double dist = side == 0 ? q[j] - m->dof_range[j][0] : m->dof_range[j][1] - q[j];
      if (!(dist < 0)) continue;
      double sg = side == 0 ? 1.0 : -1.0;
      memset(J[nrow], 0, sizeof(double) * NV);
      J[nrow][j] = sg;
      double imp = impedance(m->solimp, dist);
      double R = fmax(MJ_MINVAL, (1 - imp) / imp * m->dof_invweight0[j]);
      D[nrow] = 1.0 / R;
      aref[nrow] = -Bc * sg * v[j] - Kc * imp * dist;
      row_con[nrow] = -1;
      lim_set |= 1ull << (j + 32 * side);
      nrow++; nlimit++;
    }
  }
  /* collision: candidates in geom order is irrelevant for the solution */
  double cpos[MAXCON][3], cdist[MAXCON], cmu[MAXCON];
  int cbody[MAXCON];
  for (int s = 0; s < m->n_sphere; s++) {
    int b = m->sphere_body[s];
    double r[3];
    matvec3(r, k->xmat[b], m->sphere_pos[s]);
    double cz = k->xpos[b][2] + r[2], dist = cz - m->sphere_radius[s];
    if (dist > 0) continue;
    cpos[ncon][0] = k->xpos[b][0] + r[0]; cpos[ncon][1] = k->xpos[b][1] + r[1];
    cpos[ncon][2] = cz - (m->sphere_radius[s] + 0.5 * dist);
    cdist[ncon] = dist; cmu[ncon] = m->sphere_mu[s]; cbody[ncon] = b; ncon++;
    con_set |= 1ull << s;
  }
  for (int x = 0; x < m->n_box; x++) {
    int b = m->box_body[x], cnt = 0;
    double ctr[3];
    matvec3(ctr, k->xmat[b], m->box_center[x]);
    for (int i = 0; i < 3; i++) ctr[i] += k->xpos[b][i];
    for (int i = 0; i < 8 && cnt < 4; i++) {
      double corner[3];
      matvec3(corner, k->xmat[b], m->box_corner[x][i]);
      double ldist = corner[2];
      if (ctr[2] + ldist > 0 || ldist > 0) continue;
      double dist = ctr[2] + ldist;
      cpos[ncon][0] = ctr[0] + corner[0]; cpos[ncon][1] = ctr[1] + corner[1];
      cpos[ncon][2] = ctr[2] + corner[2] - 0.5 * dist;
      cdist[ncon] = dist; cmu[ncon] = m->box_mu[x]; cbody[ncon] = b; ncon++; cnt++;
      con_set |= 1ull << (m->n_sphere + 8 * x + i);
    }
  }
  g_last_con = con_set; g_last_lim = lim_set;
  /* pyramidal rows (condim 3): J_n +- mu J_t1, J_n +- mu J_t2; contact frame n = +z, tangents y and x */
  for (int ci = 0; ci < ncon; ci++) {
    int b = cbody[ci];
    double Jp[3][NV];
    for (int j = 0; j < nv; j++) {
      double col[3] = {0, 0, 0};
      if (supports(m, j, b)) {
        if (m->dof_type[j] == 0) memcpy(col, k->axis[j], sizeof col);
        else {
          double r[3] = {cpos[ci][0] - k->anchor[j][0], cpos[ci][1] - k->anchor[j][1], cpos[ci][2] - k->anchor[j][2]};
          cross3(col, k->axis[j], r);
        }
      }
      for (int i = 0; i < 3; i++) Jp[i][j] = col[i];
    }
    double mu = cmu[ci], imp = impedance(m->solimp, cdist[ci]);
    double tran = m->body_invweight0[b][0];
    double Rn = fmax(MJ_MINVAL, (1 - imp) / imp * (tran + mu * mu * tran));
    double Rpy = 2 * mu * mu * Rn;
    for (int r4 = 0; r4 < 4; r4++) {
      double vel = 0;
      for (int j = 0; j < nv; j++) {
        J[nrow][j] = Jp[2][j] + mu * (dirs[r4][0] * Jp[0][j] + dirs[r4][1] * Jp[1][j]);
        vel += J[nrow][j] * v[j];
      }
      D[nrow] = 1.0 / Rpy;
      aref[nrow] = -Bc * vel - Kc * imp * cdist[ci];
      row_con[nrow] = ci;
      nrow++;
    }
  }

  /* ---- primal solve: min 1/2 (a-a0)' M (a-a0) + sum_i 1/2 D_i min(0, J_i a - aref_i)^2 ---- */
  double a[NV];
  if (qacc_warm) memcpy(a, qacc_warm, sizeof(double) * nv); else memcpy(a, a0, sizeof(double) * nv);
  int iters = 0;
  double res[MAXROW];
  if (nrow == 0) {
    memcpy(a, a0, sizeof(double) * nv);
  } else if (g_solver_mode == 1) {
    /* experimental mirror of the GPU kernel's solver: primal active-set iteration with full Newton steps
       (no line search); the fixed point is the same unique minimiser.  Statistics in g_as_hist. */
    unsigned char act[MAXROW], nact[MAXROW];
    for (int i = 0; i < nrow; i++) {
      double s = -aref[i];
      for (int j = 0; j < nv; j++) s += J[i][j] * a[j];
      act[i] = s < 0;
    }
    int conv = 0;
    for (iters = 0; iters < g_as_maxit; iters++) {
      double H[NV * NV], rhs[NV];
      memcpy(H, M, sizeof(double) * nv * nv);
      memcpy(rhs, tau, sizeof(double) * nv);
      for (int i = 0; i < nrow; i++)
        if (act[i])
          for (int r = 0; r < nv; r++) {
            rhs[r] += D[i] * aref[i] * J[i][r];
            for (int cc = 0; cc < nv; cc++) H[r * nv + cc] += D[i] * J[i][r] * J[i][cc];
          }
      if (chol_solve(H, rhs, nv)) return -1;
      memcpy(a, rhs, sizeof(double) * nv);
      int same = 1;
      for (int i = 0; i < nrow; i++) {
        double s = -aref[i];
        for (int j = 0; j < nv; j++) s += J[i][j] * a[j];
        nact[i] = s < 0;
        if (nact[i] != act[i]) same = 0;
      }
      if (nrow > 0) memcpy(act, nact, (size_t)nrow);
      if (same) { conv = 1; iters++; break; }
    }
    g_as_hist[iters < 63 ? iters : 63]++;
    if (!conv) g_as_fail++;
  } else {
    for (iters = 0; iters < 200; iters++) {
      double g[NV], H[NV * NV], dlt[NV];
      memcpy(H, M, sizeof(double) * nv * nv);
      for (int r = 0; r < nv; r++) {
        double s = 0;
        for (int cc = 0; cc < nv; cc++) s += M[r * nv + cc] * (a[cc] - a0[cc]);
        g[r] = s;
      }
      for (int i = 0; i < nrow; i++) {
        double s = -aref[i];
        for (int j = 0; j < nv; j++) s += J[i][j] * a[j];
        res[i] = s;
        if (s < 0) {
          for (int r = 0; r < nv; r++) {
            g[r] += D[i] * s * J[i][r];
            for (int cc = 0; cc < nv; cc++) H[r * nv + cc] += D[i] * J[i][r] * J[i][cc];
          }
        }
      }
      for (int r = 0; r < nv; r++) dlt[r] = -g[r];
      if (chol_solve(H, dlt, nv)) return -1;
      /* exact line search on the piecewise-quadratic cost along dlt */
      double p0 = 0, p1 = 0, jv[MAXROW];
      for (int r = 0; r < nv; r++) {
        double Md = 0;
        for (int cc = 0; cc < nv; cc++) Md += M[r * nv + cc] * dlt[cc];
        p0 += (a[r] - a0[r]) * Md;
        p1 += dlt[r] * Md;
      }
      if (!(p1 > 0)) { iters++; break; }
      Brk brk[MAXROW];
      unsigned char act[MAXROW];
      int nbk = 0;
      double s0 = p0, s1 = p1; /* phi'(alpha) = s0 + alpha*s1 on the current segment */
      for (int i = 0; i < nrow; i++) {
        double s = 0;
        for (int j = 0; j < nv; j++) s += J[i][j] * dlt[j];
        jv[i] = s;
        act[i] = (res[i] < 0 || (res[i] == 0 && s < 0)) ? 1 : 0;
        if (act[i]) { s0 += D[i] * s * res[i]; s1 += D[i] * s * s; }
        if (s != 0) {
          double al = -res[i] / s;
          if (al > 0) { brk[nbk].a = al; brk[nbk].i = i; nbk++; }
        }
      }
      qsort(brk, nbk, sizeof(Brk), brk_cmp);
      double alpha = 0;
      int done = 0;
      for (int bi = 0; bi <= nbk; bi++) {
        double hi = bi < nbk ? brk[bi].a : INFINITY;
        double cand = -s0 / s1;
        if (cand <= hi) { alpha = cand; done = 1; break; }
        int i = brk[bi].i; /* a row changes status exactly once along the ray */
        double sg = act[i] ? -1.0 : 1.0;
        s0 += sg * D[i] * jv[i] * res[i];
        s1 += sg * D[i] * jv[i] * jv[i];
        act[i] ^= 1;
      }
      if (!done || !isfinite(alpha)) alpha = 1.0;
      double step2 = 0, scale2 = 1e-30;
      for (int r = 0; r < nv; r++) { a[r] += alpha * dlt[r]; step2 += alpha * alpha * dlt[r] * dlt[r]; scale2 += a[r] * a[r]; }
      if (step2 <= 1e-20 * scale2) { iters++; break; }
    }
  }
  memcpy(qacc, a, sizeof(double) * nv);
  if (qacc_warm) memcpy(qacc_warm, a, sizeof(double) * nv);

  if (diag) {
    memset(diag, 0, sizeof *diag);
    diag->ncon = ncon; diag->nefc = nrow; diag->nlimit = nlimit; diag->iters = iters;
    double r[NV];
    for (int i = 0; i < nv; i++) {
      double s = 0;
      for (int j = 0; j < nv; j++) s += M[i * nv + j] * (a[j] - a0[j]);
      r[i] = s;
    }
    for (int i = 0; i < nrow; i++) {
      double s = -aref[i];
      for (int j = 0; j < nv; j++) s += J[i][j] * a[j];
      double f = s < 0 ? -D[i] * s : 0;
      for (int j = 0; j < nv; j++) r[j] -= J[i][j] * f;
      int ci = row_con[i];
      if (ci >= 0) {
        int r4 = (i - nlimit) % 4;
        diag->con_force[ci][0] += f;
        diag->con_force[ci][2] += cmu[ci] * dirs[r4][0] * f;
        diag->con_force[ci][1] += cmu[ci] * dirs[r4][1] * f;
        diag->normal_force += f;
      }
    }
    double mx = 0;
    for (int i = 0; i < nv; i++) mx = fmax(mx, fabs(r[i]));
    diag->kkt_residual = mx;
    for (int ci = 0; ci < ncon; ci++) {
      memcpy(diag->con_pos[ci], cpos[ci], sizeof cpos[ci]);
      diag->con_dist[ci] = cdist[ci]; diag->con_body[ci] = cbody[ci];
    }
    diag->energy_kin = e_kin;
    double ep = 0;
    for (int b = 0; b < m->nb; b++) ep -= m->body_mass[b] * m->gravity_z * k->com[b][2];
    diag->energy_pot = ep;
  }
  return 0;
}

/* nsub MuJoCo steps with ctrl held (MujocoEnv.do_simulation).  integrator: DRL_INTEGRATOR_RK4 / _EULER.
 * Returns 0, or 1 if the state left the finite range MuJoCo accepts (mj_checkPos/Vel/Acc -> MujocoException). */
static int step_sig(const DrlWalkerModel* m, double* q, double* v, const double* ctrl, double* qacc_warm, int nsub,
                    int integrator, unsigned long long* sig /* nullable: [nsub][4][2] */) {
  int nv = m->nv;
  double h = m->timestep;
  if (sig) memset(sig, 0, sizeof(unsigned long long) * (size_t)nsub * 8);
  for (int s = 0; s < nsub; s++) {
    for (int j = 0; j < nv; j++)
      if (!isfinite(q[j]) || !isfinite(v[j]) || fabs(q[j]) > 1e10 || fabs(v[j]) > 1e10) return 1;
    if (integrator == DRL_INTEGRATOR_RK4) {
      static const double A[3] = {0.5, 0.5, 1.0}, Bw[4] = {1.0 / 6, 1.0 / 3, 1.0 / 3, 1.0 / 6};
      double kq[4][NV], kv[4][NV], qs[NV], vs[NV];
      memcpy(qs, q, sizeof(double) * nv); memcpy(vs, v, sizeof(double) * nv);
      for (int st = 0; st < 4; st++) {
        if (orc_forward(m, qs, vs, ctrl, qacc_warm, kv[st], NULL)) return 1;
        if (sig) { sig[(s * 4 + st) * 2] = g_last_con; sig[(s * 4 + st) * 2 + 1] = g_last_lim; }
        memcpy(kq[st], vs, sizeof(double) * nv);
        if (st < 3)
          for (int j = 0; j < nv; j++) { qs[j] = q[j] + h * A[st] * kq[st][j]; vs[j] = v[j] + h * A[st] * kv[st][j]; }
      }
      for (int j = 0; j < nv; j++) {
        double dq = 0, dv = 0;
        for (int st = 0; st < 4; st++) { dq += Bw[st] * kq[st][j]; dv += Bw[st] * kv[st][j]; }
        q[j] += h * dq; v[j] += h * dv;
      }
    } else {
      /* mj_Euler: implicit in joint damping: (M + h B) a' = M a + ... == tau_total; here M a = tau + J'f */
      double a[NV];
      if (orc_forward(m, q, v, ctrl, qacc_warm, a, NULL)) return 1;
      if (sig) { sig[(s * 4) * 2] = g_last_con; sig[(s * 4) * 2 + 1] = g_last_lim; }
      Kin* k = (Kin*)malloc(sizeof(Kin));
      double M[NV * NV], rhs[NV];
      kinematics(m, q, k);
      mass_matrix(m, k, M);
      for (int r = 0; r < nv; r++) {
        double sacc = 0;
        for (int c2 = 0; c2 < nv; c2++) sacc += M[r * nv + c2] * a[c2];
        rhs[r] = sacc;
      }
      for (int j = 0; j < nv; j++) M[j * nv + j] += h * m->dof_damping[j];
      int bad = chol_solve(M, rhs, nv);
      free(k);
      if (bad) return 1;
      for (int j = 0; j < nv; j++) { v[j] += h * rhs[j]; q[j] += h * v[j]; }
    }
  }
  for (int j = 0; j < nv; j++)
    if (!isfinite(q[j]) || !isfinite(v[j]) || fabs(q[j]) > 1e10 || fabs(v[j]) > 1e10) return 1;
  return 0;
}

int orc_step(const DrlWalkerModel* m, double* q, double* v, const double* ctrl, double* qacc_warm, int nsub,
             int integrator) {
  return step_sig(m, q, v, ctrl, qacc_warm, nsub, integrator, NULL);
}

/* orc_step that also reports, for every dynamics evaluation (substep x RK4 stage), which contact candidates and limit
 * rows were instantiated: the constraint-set signature the contact-sensitivity probe compares. */
int orc_step_trace(const DrlWalkerModel* m, double* q, double* v, const double* ctrl, double* qacc_warm, int nsub,
                   int integrator, unsigned long long* sig) {
  return step_sig(m, q, v, ctrl, qacc_warm, nsub, integrator, sig);
}

/* site world positions (sim.data.site_xpos after set_state / sim.forward, mimic_env.py:549) */
void orc_site_xpos(const DrlWalkerModel* m, const double* q, double* out /* n_site*3 */) {
  Kin* k = (Kin*)malloc(sizeof(Kin));
  kinematics(m, q, k);
  for (int s = 0; s < m->n_site; s++) {
    int b = m->site_body[s];
    double r[3];
    matvec3(r, k->xmat[b], m->site_pos[s]);
    for (int i = 0; i < 3; i++) out[3 * s + i] = k->xpos[b][i] + r[i];
  }
  free(k);
}

/* exposed pieces for first-principles tests */
void orc_mass_matrix(const DrlWalkerModel* m, const double* q, double* M) {
  Kin* k = (Kin*)malloc(sizeof(Kin));
  kinematics(m, q, k);
  mass_matrix(m, k, M);
  free(k);
}
void orc_bias(const DrlWalkerModel* m, const double* q, const double* v, double* c) {
  Kin* k = (Kin*)malloc(sizeof(Kin));
  kinematics(m, q, k);
  bias_force(m, k, q, v, c, NULL);
  free(k);
}
void orc_body_com(const DrlWalkerModel* m, const double* q, double* com /* nb*3 */, double* xmat /* nb*9 */) {
  Kin* k = (Kin*)malloc(sizeof(Kin));
  kinematics(m, q, k);
  for (int b = 0; b < m->nb; b++) {
    memcpy(com + 3 * b, k->com[b], sizeof(double) * 3);
    memcpy(xmat + 9 * b, k->xmat[b], sizeof(double) * 9);
  }
  free(k);
}
int orc_diag_size(void) { return (int)sizeof(OrcDiag); }
int orc_max_con(void) { return MAXCON; }
