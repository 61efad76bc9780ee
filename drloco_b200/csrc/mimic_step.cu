// mimic_step.cu — the hot path of drloco_b200: one fused sm_100a kernel per VecEnv.step().
//
// Replaces, for N environments at once, the reference's per-process Python + MuJoCo stack:
//   MimicEnv.step                drloco/mujoco/mimic_env.py:60-126
//   MujocoEnv.do_simulation      frame_skip x mj_step with RK4 (xml:11), restated in oracle/walker_physics.c
//   StraightWalkingTrajectories.next / BaseReferenceTrajectories.next   straight_walk_trajecs.py:141-159, base:95-103
//   get_imitation_reward / _get_ET_reward / _get_obs / mirror_*          mimic_env.py:142-168,403-489,592-649
//   reset_model (RSI + ground-contact shift)                             mimic_env.py:526-572
//   Monitor.step statistics                                              monitor_wrapper.py:88-166
//   DummyVecEnv/SubprocVecEnv auto-reset with terminal_observation       (SB3 1.0)
//
// Mapping: one environment per group of G lanes (G = 16 for nv <= 16: two environments per warp; G = 32 otherwise),
// lane j owns dof j: its q/v/RK4 accumulators, its motion vector S_j and column j of the constraint Hessian live in
// registers for the whole launch; per-env tree quantities (body frames, spatial inertias, velocities, contact
// Hessians) live in shared memory; all frame_skip x 4 dynamics evaluations run inside the launch, HBM is touched once
// on entry and once on exit.  Control flow is kept warp-uniform (the two environments of a warp run in lockstep, with
// predicated effects), so every shuffle / ballot / barrier uses the full-warp mask and compiles to a bare instruction.
// Spatial quantities are expressed in world orientation about O = the root body origin (keeps fp32 cancellation
// independent of how far the walker has travelled).
//
// Constraint solve per evaluation (same minimiser as MuJoCo's Newton solver, see oracle/walker_physics.c):
//   H = M + sum_b S_b^T W_b S_b + diag(limits),   H qacc = tau - c + sum_b S_b^T u_b + limits
// with W_b the 6x6 wrench-space Hessian of the active pyramid rows of all contacts on body b; primal active-set
// iteration with full Newton steps, warm-started from the previous evaluation; LDL^T in registers via shuffles.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/drloco_b200.h"
#include "dev_model.h"

namespace drl {

constexpr int kNPass = 2;          // contact candidates per lane
constexpr int kMaxSolverIter = 10;
constexpr float kMinVal = 1e-15f;
constexpr unsigned kFull = 0xFFFFFFFFu;

template <int G>
struct EnvSmem {
  float v[G];               // qvel at the current stage
  float acc[G];             // qacc iterate
  float tau[G];             // actuator force per dof
  float cssn[G][2];         // cos, sin of hinge angles (slides: -, displacement)
  float axw[G][4];          // joint axes in world orientation
  float S[G][12];           // motion vectors (omega, v_O); row stride 12 floats: lanes reading different rows hit different banks
  float Fd[G][12];          // bias-acceleration terms during RNE, then Ic * S (same stride)
  float bodyR[kMaxBody][12];  // rotation (row major) + position relative to O
  float Ib[kMaxBody][12];   // spatial inertia about O: m, h[3], Ixx Ixy Ixz Iyy Iyz Izz
  float Ic[kMaxBody][12];   // composite
  union {
    struct {
      float V[kMaxBody][8];     // spatial velocity
      float A[kMaxBody][8];     // body force (n, f)
      float T[kMaxBody][8];     // S_b * qacc
      float W[kMaxBody][24];    // contact Hessian, 21 unique entries
      float U[kMaxBody][8];     // contact rhs wrench
    };
    // lower triangle of the mass matrix with an odd row stride (transposed without bank conflicts).  Lives between
    // the last use of V / A (bias force) and the first use of W / U / T (constraint solve).
    float Mt[kMaxBody * 56];
  };
  float obsbuf[kMaxObs];
  float Mc[(G == 16 ? 14 : 19) * G];   // mass-matrix column of each lane: Mc[r * G + l] (kept out of registers)
};

__device__ __forceinline__ void cross3(float& rx, float& ry, float& rz, float ax, float ay, float az, float bx,
                                       float by, float bz) {
  rx = ay * bz - az * by;
  ry = az * bx - ax * bz;
  rz = ax * by - ay * bx;
}

struct Vec6 {
  float w0, w1, w2, v0, v1, v2;
};

__device__ __forceinline__ Vec6 ld6(const float* p) {
  float4 a = *reinterpret_cast<const float4*>(p);
  float2 b = *reinterpret_cast<const float2*>(p + 4);
  return Vec6{a.x, a.y, a.z, a.w, b.x, b.y};
}
__device__ __forceinline__ void st6(float* p, const Vec6& x) {
  *reinterpret_cast<float4*>(p) = make_float4(x.w0, x.w1, x.w2, x.v0);
  *reinterpret_cast<float2*>(p + 4) = make_float2(x.v1, x.v2);
}
__device__ __forceinline__ float dot6(const Vec6& a, const Vec6& b) {
  return a.w0 * b.w0 + a.w1 * b.w1 + a.w2 * b.w2 + a.v0 * b.v0 + a.v1 * b.v1 + a.v2 * b.v2;
}
__device__ __forceinline__ void axpy6(Vec6& y, float a, const Vec6& x) {
  y.w0 += a * x.w0; y.w1 += a * x.w1; y.w2 += a * x.w2;
  y.v0 += a * x.v0; y.v1 += a * x.v1; y.v2 += a * x.v2;
}

// spatial inertia (m, h, I about O) times twist -> momentum/wrench (angular, linear)
__device__ __forceinline__ Vec6 inertia_mul(const float* I, const Vec6& t) {
  float4 a = *reinterpret_cast<const float4*>(I);       // m hx hy hz
  float4 b = *reinterpret_cast<const float4*>(I + 4);   // Ixx Ixy Ixz Iyy
  float2 c = *reinterpret_cast<const float2*>(I + 8);   // Iyz Izz
  float m = a.x, hx = a.y, hy = a.z, hz = a.w;
  Vec6 r;
  float cx, cy, cz;
  cross3(cx, cy, cz, hx, hy, hz, t.v0, t.v1, t.v2);      // angular: I w + h x v
  r.w0 = b.x * t.w0 + b.y * t.w1 + b.z * t.w2 + cx;
  r.w1 = b.y * t.w0 + b.w * t.w1 + c.x * t.w2 + cy;
  r.w2 = b.z * t.w0 + c.x * t.w1 + c.y * t.w2 + cz;
  cross3(cx, cy, cz, t.w0, t.w1, t.w2, hx, hy, hz);      // linear: m v + w x h
  r.v0 = m * t.v0 + cx;
  r.v1 = m * t.v1 + cy;
  r.v2 = m * t.v2 + cz;
  return r;
}

// reciprocal: hardware approximation + one Newton step (relative error ~1e-7, no IEEE-division slow path)
__device__ __forceinline__ float fast_rcp(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return fmaf(r, fmaf(-x, r, 1.f), r);
}

// general solimp power (MuJoCo default is 2, handled inline by impedance())
__device__ __noinline__ float impedance_pow(float x, float mid, float power) {
  return (x <= mid) ? powf(x, power) / powf(mid, power - 1.f)
                    : 1.f - powf(1.f - x, power) / powf(1.f - mid, power - 1.f);
}

__device__ __forceinline__ float impedance(const DevModel& M, float dist) {
  float x = fabsf(dist) * M.imp_inv_width;
  if (x >= 1.f) return M.imp_dmax;
  if (x <= 0.f) return M.imp_d0;
  float y;
  if (M.imp_power == 2.f) {
    y = (x <= M.imp_mid) ? x * x * M.imp_inv_mid : 1.f - (1.f - x) * (1.f - x) * M.imp_inv_1mmid;
  } else if (M.imp_power == 1.f) {
    y = x;
  } else {
    y = impedance_pow(x, M.imp_mid, M.imp_power);
  }
  return M.imp_d0 + y * (M.imp_dmax - M.imp_d0);
}

// per-lane role constants
struct LaneConst {
  int l;               // lane within the env group
  unsigned emask;      // lanes of this lane's environment within the warp
  bool isdof, isbody;
  int body, type, limited, last;
  float sign, ref, damping, armature, lo, hi, invw;
  unsigned anc, desc, subb;
};

struct Counters {
  int evals, iters, capped;
};

// active set carried from one dynamics evaluation to the next (lane <-> contact candidate is a fixed mapping)
struct ActiveSet {
  unsigned bits[kNPass];  // active pyramid rows of this lane's candidate in each pass
  unsigned prev_act;      // which of this lane's candidates were in contact at the previous evaluation
  bool lbit, prev_lim;    // joint-limit row of this lane's dof
};

// symmetric 6x6 index into 21 packed entries (i <= j)
__device__ __forceinline__ constexpr int sym6(int i, int j) {
  return (i <= j) ? (i * 6 - (i * (i - 1)) / 2 + (j - i)) : (j * 6 - (j * (j - 1)) / 2 + (i - j));
}

// pop the two lowest set bits of a mask (i1 = i0 and second = false when only one is left): the chain loops below
// consume two entries per trip so that their shared-memory loads are in flight together
__device__ __forceinline__ void pop2(unsigned& mk, int& i0, int& i1, bool& second) {
  i0 = __ffs(mk) - 1;
  mk &= mk - 1;
  second = mk != 0u;
  i1 = second ? __ffs(mk) - 1 : i0;
  mk &= mk - 1;
}

// does predicate p hold on any lane of this lane's environment?
__device__ __forceinline__ bool env_any(bool p, unsigned emask) { return (__ballot_sync(kFull, p) & emask) != 0u; }

// LDL^T solve with the symmetric matrix spread one column per lane: H[0..NV-1] = rows of this lane's column (full
// column, both triangles), H[NV] = this lane's rhs entry.  Right-looking elimination; column k is left unscaled
// (H[r][k] = l_rk d_k) so that each trailing update is one shuffle + one FMA.  Returns x for this lane's row.
template <int NV, int G>
__device__ __forceinline__ float ldl_solve_cols(float (&H)[NV + 1], int l) {
  float invd = 0.f;
#pragma unroll
  for (int k = 0; k < NV; k++) {
    const float dk = __shfl_sync(kFull, H[k], k, G);
    const float inv = fast_rcp(fmaxf(dk, 1e-30f));
    const float lck = H[k] * inv;          // lanes c > k: l_ck = H[c][k] / d_k (H is symmetric)
    const bool upd = l > k;
    if (l == k) invd = inv;
#pragma unroll
    for (int r = k + 1; r <= NV; r++) {
      const float vr = __shfl_sync(kFull, H[r], k, G);     // H[r][k];  r == NV: forward-substituted rhs z_k
      if (upd) H[r] = fmaf(-vr, lck, H[r]);
    }
  }
  // x_c = (z_c - sum_{r>c} H[r][c] x_r) / d_c
  float sacc = H[NV], x = 0.f;
#pragma unroll
  for (int k = NV - 1; k >= 0; k--) {
    const float xk = __shfl_sync(kFull, sacc * invd, k, G);
    if (l < k) sacc = fmaf(-H[k], xk, sacc);
    if (l == k) x = xk;
  }
  return x;
}

// Kinematics of the tree for the joint configuration published in E.sn / E.cs: body frames relative to O and world
// joint axes.  (mj_kinematics for hinge joints anchored at the body origin; root slides move O itself.)
template <int G>
__device__ __forceinline__ void tree_kinematics(const DevModel& M, EnvSmem<G>& E, int l) {
  const int slot = l / 3, r = l - 3 * slot;
  for (int lev = 0; lev < M.nlevel; lev++) {
    if (slot < M.level_count[lev]) {
      const int b = M.level_body[lev][slot], p = M.body_parent[b];
      float R0, R1, R2, pr;
      if (p < 0) {
        R0 = r == 0 ? 1.f : 0.f; R1 = r == 1 ? 1.f : 0.f; R2 = r == 2 ? 1.f : 0.f; pr = 0.f;
      } else {
        R0 = E.bodyR[p][3 * r]; R1 = E.bodyR[p][3 * r + 1]; R2 = E.bodyR[p][3 * r + 2];
        pr = E.bodyR[p][9 + r] + R0 * M.body_pos[b][0] + R1 * M.body_pos[b][1] + R2 * M.body_pos[b][2];
      }
      // hinges only: the root slides translate O itself and their world axes are constants (see forward_dynamics)
      const int j0 = M.body_hinge0[b], j1 = M.body_dof0[b] + M.body_ndof[b];
      for (int j = j0; j < j1; j++) {
        const int code = M.dof_code[j];                 // axis index | negative-axis flag << 2
        const int k = code & 3;
        const float ax = k == 0 ? R0 : (k == 1 ? R1 : R2);
        E.axw[j][r] = (code & 4) ? -ax : ax;
        const float2 cs = *reinterpret_cast<const float2*>(&E.cssn[j][0]);
        const float c = cs.x, sn = cs.y;
        if (k == 0) { float u = R1, w = R2; R1 = c * u + sn * w; R2 = c * w - sn * u; }
        else if (k == 1) { float u = R2, w = R0; R2 = c * u + sn * w; R0 = c * w - sn * u; }
        else { float u = R0, w = R1; R0 = c * u + sn * w; R1 = c * w - sn * u; }
      }
      E.bodyR[b][3 * r] = R0; E.bodyR[b][3 * r + 1] = R1; E.bodyR[b][3 * r + 2] = R2;
      E.bodyR[b][9 + r] = pr;
    }
    __syncwarp();
  }
}

// Column l of the joint-space inertia matrix.  M[r][c] = S_c . (Ic_{body(r)} S_r) for r = c or a descendant of c (CRBA,
// E.Fd holds Ic S).  Each lane computes the part of its column at and below the diagonal; the part above comes from
// the transposed entries through shared memory (odd row stride: the row-wise store and the column-wise load are both
// conflict-free).  E.Mt aliases V/A/T/W/U: callers guarantee those are dead; ends with a barrier.
template <int NV, int G>
__device__ __forceinline__ void mass_column(EnvSmem<G>& E, const LaneConst& L, const Vec6& S) {
  float Mcol[NV];
  constexpr int kMs = (NV % 2 == 0) ? NV + 1 : NV + 2;
  static_assert(NV * kMs <= (int)(sizeof(E.Mt) / sizeof(float)), "Mt too small");
  const int l = L.l;
  const unsigned lowmask = L.isdof ? (L.desc | (1u << l)) : 0u;
#pragma unroll
  for (int r = 0; r < NV; r++) {
    const float d = dot6(S, ld6(E.Fd[r]));
    Mcol[r] = ((lowmask >> r) & 1u) ? d : 0.f;
    if (L.isdof) E.Mt[r * kMs + l] = Mcol[r];
  }
  __syncwarp();
#pragma unroll
  for (int r = 0; r < NV; r++) {
    if (L.isdof && r < l) Mcol[r] = E.Mt[l * kMs + r];
    if (r == l) Mcol[r] += L.armature;
    E.Mc[r * G + l] = Mcol[r];
  }
  __syncwarp();
}

// One forward-dynamics evaluation (mj_forward).  q, v: this lane's coordinates; a: warm start in, qacc out.
// Must be called by all 32 lanes of the warp (warp-uniform control flow).
template <int NV, int G, bool DBG>
__device__ __forceinline__ void forward_dynamics(const DevModel& M, EnvSmem<G>& E, const LaneConst& L, float q,
                                                 float v, float tau, float& a, ActiveSet& AS, Counters& cnt,
                                                 float* dbg) {
  const int l = L.l;
  // ---- 1. publish joint trig + velocity -------------------------------------------------------------
  {
    float s = q - L.ref, c = 1.f;
    if (L.isdof && L.type == 1) sincosf(L.sign * (q - L.ref), &s, &c);
    if (L.isdof) { *reinterpret_cast<float2*>(&E.cssn[l][0]) = make_float2(c, s); E.v[l] = v; }
  }
  __syncwarp();
  float zO = M.root_z0;
  for (int j = 0; j < M.nslide; j++) zO = fmaf(M.dof_slide_z[j], E.cssn[j][1], zO);
  // ---- 2. body frames ----------------------------------------------------------------------------
  tree_kinematics<G>(M, E, l);
  // ---- 3. motion vectors, body inertias about O ------------------------------------------------------
  Vec6 S = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (L.isdof) {
    if (L.type == 1) {
      const float ax = E.axw[l][0], ay = E.axw[l][1], az = E.axw[l][2];
      const float px = E.bodyR[L.body][9], py = E.bodyR[L.body][10], pz = E.bodyR[L.body][11];
      S.w0 = ax; S.w1 = ay; S.w2 = az;
      cross3(S.v0, S.v1, S.v2, px, py, pz, ax, ay, az);     // v_O = anchor x axis
    } else {                                                // root slide: constant world axis +-e_k
      const int k = M.dof_code[l] & 3;
      const float sg = (M.dof_code[l] & 4) ? -1.f : 1.f;
      S.v0 = k == 0 ? sg : 0.f; S.v1 = k == 1 ? sg : 0.f; S.v2 = k == 2 ? sg : 0.f;
    }
    st6(E.S[l], S);
  }
  if (L.isbody) {
    const float* R = E.bodyR[l];
    const float ix = M.body_ipos[l][0], iy = M.body_ipos[l][1], iz = M.body_ipos[l][2];
    const float m = M.body_mass[l];
    const float cx = R[9] + R[0] * ix + R[1] * iy + R[2] * iz;
    const float cy = R[10] + R[3] * ix + R[4] * iy + R[5] * iz;
    const float cz = R[11] + R[6] * ix + R[7] * iy + R[8] * iz;
    const float I0 = M.body_inertia[l][0], I1 = M.body_inertia[l][1], I2 = M.body_inertia[l][2];
    float Ixx = R[0] * R[0] * I0 + R[1] * R[1] * I1 + R[2] * R[2] * I2;
    float Ixy = R[0] * R[3] * I0 + R[1] * R[4] * I1 + R[2] * R[5] * I2;
    float Ixz = R[0] * R[6] * I0 + R[1] * R[7] * I1 + R[2] * R[8] * I2;
    float Iyy = R[3] * R[3] * I0 + R[4] * R[4] * I1 + R[5] * R[5] * I2;
    float Iyz = R[3] * R[6] * I0 + R[4] * R[7] * I1 + R[5] * R[8] * I2;
    float Izz = R[6] * R[6] * I0 + R[7] * R[7] * I1 + R[8] * R[8] * I2;
    Ixx += m * (cy * cy + cz * cz); Iyy += m * (cx * cx + cz * cz); Izz += m * (cx * cx + cy * cy);
    Ixy -= m * cx * cy; Ixz -= m * cx * cz; Iyz -= m * cy * cz;
    float* I = E.Ib[l];
    *reinterpret_cast<float4*>(I) = make_float4(m, m * cx, m * cy, m * cz);
    *reinterpret_cast<float4*>(I + 4) = make_float4(Ixx, Ixy, Ixz, Iyy);
    *reinterpret_cast<float2*>(I + 8) = make_float2(Iyz, Izz);
  }
  __syncwarp();
  // ---- 4. velocities (RNE forward), composite inertias -------------------------------------------------
  if (L.isdof) {
    Vec6 Vp = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (unsigned mk = L.anc; mk;) {
      int i0, i1; bool two;
      pop2(mk, i0, i1, two);
      const Vec6 s0 = ld6(E.S[i0]), s1 = ld6(E.S[i1]);
      const float v0 = E.v[i0], v1 = two ? E.v[i1] : 0.f;
      axpy6(Vp, v0, s0);
      axpy6(Vp, v1, s1);
    }
    // cdof_dot * v = (Vp x_m S) v
    Vec6 cd;
    float tx, ty, tz;
    cross3(cd.w0, cd.w1, cd.w2, Vp.w0, Vp.w1, Vp.w2, S.w0, S.w1, S.w2);
    cross3(cd.v0, cd.v1, cd.v2, Vp.w0, Vp.w1, Vp.w2, S.v0, S.v1, S.v2);
    cross3(tx, ty, tz, Vp.v0, Vp.v1, Vp.v2, S.w0, S.w1, S.w2);
    cd.v0 += tx; cd.v1 += ty; cd.v2 += tz;
    cd.w0 *= v; cd.w1 *= v; cd.w2 *= v; cd.v0 *= v; cd.v1 *= v; cd.v2 *= v;
    st6(E.Fd[l], cd);
    if (L.last) {
      axpy6(Vp, v, S);
      st6(E.V[L.body], Vp);
    }
  }
  if (L.isbody) {
    float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
    float2 a2 = make_float2(0.f, 0.f);
    for (unsigned mk = M.body_sub[l]; mk;) {
      int b0, b1; bool two;
      pop2(mk, b0, b1, two);
      const float w = two ? 1.f : 0.f;
      const float4 x0 = *reinterpret_cast<const float4*>(E.Ib[b0]), y0 = *reinterpret_cast<const float4*>(E.Ib[b1]);
      const float4 x1 = *reinterpret_cast<const float4*>(E.Ib[b0] + 4), y1 = *reinterpret_cast<const float4*>(E.Ib[b1] + 4);
      const float2 x2 = *reinterpret_cast<const float2*>(E.Ib[b0] + 8), y2 = *reinterpret_cast<const float2*>(E.Ib[b1] + 8);
      a0.x += x0.x + w * y0.x; a0.y += x0.y + w * y0.y; a0.z += x0.z + w * y0.z; a0.w += x0.w + w * y0.w;
      a1.x += x1.x + w * y1.x; a1.y += x1.y + w * y1.y; a1.z += x1.z + w * y1.z; a1.w += x1.w + w * y1.w;
      a2.x += x2.x + w * y2.x; a2.y += x2.y + w * y2.y;
    }
    *reinterpret_cast<float4*>(E.Ic[l]) = a0;
    *reinterpret_cast<float4*>(E.Ic[l] + 4) = a1;
    *reinterpret_cast<float2*>(E.Ic[l] + 8) = a2;
  }
  __syncwarp();
  // ---- 5. body forces; contact candidates ---------------------------------------------------------------
  if (L.isbody) {
    Vec6 Ab = {0.f, 0.f, 0.f, 0.f, 0.f, -M.gravity_z};     // fictitious base acceleration = -gravity
    for (unsigned mk = M.body_supp[l]; mk;) {
      int i0, i1; bool two;
      pop2(mk, i0, i1, two);
      const Vec6 f0 = ld6(E.Fd[i0]), f1 = ld6(E.Fd[i1]);
      axpy6(Ab, 1.f, f0);
      axpy6(Ab, two ? 1.f : 0.f, f1);
    }
    const Vec6 Vb = ld6(E.V[l]);
    Vec6 f = inertia_mul(E.Ib[l], Ab);
    const Vec6 mom = inertia_mul(E.Ib[l], Vb);
    float tx, ty, tz;
    cross3(tx, ty, tz, Vb.w0, Vb.w1, Vb.w2, mom.w0, mom.w1, mom.w2);
    f.w0 += tx; f.w1 += ty; f.w2 += tz;
    cross3(tx, ty, tz, Vb.v0, Vb.v1, Vb.v2, mom.v0, mom.v1, mom.v2);
    f.w0 += tx; f.w1 += ty; f.w2 += tz;
    cross3(tx, ty, tz, Vb.w0, Vb.w1, Vb.w2, mom.v0, mom.v1, mom.v2);
    f.v0 += tx; f.v1 += ty; f.v2 += tz;
    st6(E.A[l], f);
  }
  // contacts: candidate s = pass*G + l
  bool cact[kNPass];
  float cPx[kNPass], cPy[kNPass], cPz[kNPass], cD[kNPass], cmu[kNPass], car[kNPass][4];
  int cbody[kNPass];
  unsigned conmask = 0;
  const int wl = threadIdx.x & 31;
#pragma unroll
  for (int ps = 0; ps < kNPass; ps++) {
    const int s = ps * G + l;
    const bool valid = s < M.ncand;
    const int b = valid ? M.cand_body[s] : 0;
    const float* R = E.bodyR[b];
    bool act = false;
    float Px = 0.f, Py = 0.f, Pz = 0.f, dist = 0.f;
    const bool isbox = s < M.nbox_cand;
    if (valid) {
      const float x = M.cand_pos[s][0], y = M.cand_pos[s][1], z = M.cand_pos[s][2];
      const float rx = R[0] * x + R[1] * y + R[2] * z;
      const float ry = R[3] * x + R[4] * y + R[5] * z;
      const float rz = R[6] * x + R[7] * y + R[8] * z;
      if (isbox) {
        const float ux = M.cand_aux[s][0], uy = M.cand_aux[s][1], uz = M.cand_aux[s][2];
        const float Cx = R[9] + R[0] * ux + R[1] * uy + R[2] * uz;
        const float Cy = R[10] + R[3] * ux + R[4] * uy + R[5] * uz;
        const float Cz = R[11] + R[6] * ux + R[7] * uy + R[8] * uz;
        const float cz = zO + Cz;
        act = !(cz + rz > 0.f || rz > 0.f);
        dist = cz + rz;
        Px = Cx + rx; Py = Cy + ry; Pz = Cz + rz - 0.5f * dist;
      } else {
        const float rad = M.cand_aux[s][0];
        const float cz = zO + R[11] + rz;
        dist = cz - rad;
        act = !(dist > 0.f);
        Px = R[9] + rx; Py = R[10] + ry; Pz = R[11] + rz - rad - 0.5f * dist;
      }
    }
    // plane-box keeps at most the first four penetrating corners (MuJoCo mjc_PlaneBox)
    if (ps == 0) {
      const unsigned bal = __ballot_sync(kFull, act && isbox);
      const unsigned seg = 0xFFu << (wl & ~7);
      const int rank = __popc(bal & seg & ((1u << wl) - 1u));
      if (isbox && rank >= 4) act = false;
    }
    cact[ps] = act; cPx[ps] = Px; cPy[ps] = Py; cPz[ps] = Pz; cbody[ps] = b;
    cD[ps] = 0.f; cmu[ps] = 0.f;
    car[ps][0] = car[ps][1] = car[ps][2] = car[ps][3] = 0.f;
    if (act) {
      const float mu = M.cand_mu[s];
      const float imp = impedance(M, dist);
      // D = 1 / (2 mu^2 R_n),  R_n = (1-imp)/imp * invweight * (1 + mu^2)
      const float Rn = fmaxf(kMinVal, (1.f - imp) * M.body_invw_tran[b] * (1.f + mu * mu));
      cD[ps] = imp * fast_rcp(2.f * mu * mu * Rn);
      cmu[ps] = mu;
      conmask |= 1u << b;
      // reference acceleration of the four pyramid rows: aref = -B (J v) - K imp dist   (E.V is complete since step 4)
      const Vec6 Vb = ld6(E.V[b]);
      float ux, uy, uz;
      cross3(ux, uy, uz, Vb.w0, Vb.w1, Vb.w2, Px, Py, Pz);
      ux += Vb.v0; uy += Vb.v1; uz += Vb.v2;
      const float base = -M.Kc * imp * dist;
      car[ps][0] = -M.Bc * (uz + mu * ux) + base;
      car[ps][1] = -M.Bc * (uz - mu * ux) + base;
      car[ps][2] = -M.Bc * (uz + mu * uy) + base;
      car[ps][3] = -M.Bc * (uz - mu * uy) + base;
    }
  }
  __syncwarp();   // E.A complete
  // bodies with a contact in either environment of the warp (W/U are kept valid for the union in both)
  conmask = __reduce_or_sync(kFull, conmask);
  // ---- 6. bias force, smooth rhs, Ic*S ---------------------------------------------------------------------
  float rhs0 = 0.f;
  Vec6 Fdc = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (L.isdof) {
    float cb = 0.f;
    for (unsigned mk = L.subb; mk;) {
      int b0, b1; bool two;
      pop2(mk, b0, b1, two);
      const Vec6 f0 = ld6(E.A[b0]), f1 = ld6(E.A[b1]);
      const float d0 = dot6(S, f0), d1 = dot6(S, f1);
      cb += d0 + (two ? d1 : 0.f);
    }
    rhs0 = tau - L.damping * v - cb;
    Fdc = inertia_mul(E.Ic[L.body], S);
    st6(E.Fd[l], Fdc);     // safe: the cdd values in E.Fd were consumed before the last barrier
    if (DBG) { dbg[0 * 32 + l] = cb; dbg[1 * 32 + l] = rhs0; }
  }
  __syncwarp();
  // ---- 7. mass-matrix column ---------------------------------------------------------------------------------
  mass_column<NV, G>(E, L, S);
  if (DBG) {
#pragma unroll
    for (int r = 0; r < NV; r++) dbg[(2 + r) * 32 + l] = E.Mc[r * G + l];
  }
  // ---- joint limits ---------------------------------------------------------------------------------------------
  float lsg = 0.f, lD = 0.f, laref = 0.f;
  if (L.isdof && L.limited) {
    float dist = 0.f;
    if (q < L.lo) { lsg = 1.f; dist = q - L.lo; }
    else if (q > L.hi) { lsg = -1.f; dist = L.hi - q; }
    if (lsg != 0.f) {
      const float imp = impedance(M, dist);
      lD = imp * fast_rcp(fmaxf(kMinVal, (1.f - imp) * L.invw));
      laref = -M.Bc * lsg * v - M.Kc * imp * dist;
    }
  }
  cnt.evals++;
  // ---- 8. active-set iteration ----------------------------------------------------------------------------------
  // The active set of the previous evaluation (same lane <-> same contact candidate) is the starting guess; a contact
  // or limit that was not present before starts with all of its rows active.  Without any constraint in the warp the
  // loop body runs once and is the plain solve M qacc = rhs0.
#pragma unroll
  for (int ps = 0; ps < kNPass; ps++) {
    if (!cact[ps]) AS.bits[ps] = 0u;
    else if (!((AS.prev_act >> ps) & 1u)) AS.bits[ps] = 0xFu;
  }
  if (lsg == 0.f) AS.lbit = false;
  else if (!AS.prev_lim) AS.lbit = true;
  const bool sph_any = __any_sync(kFull, cact[1]);
  const bool any_limit = __any_sync(kFull, lsg != 0.f);
  const bool constrained = (conmask != 0u) || any_limit;
  float H[NV + 1];
  for (int it = 0; it < kMaxSolverIter; it++) {
    if (constrained) cnt.iters++;
    if (conmask != 0u) {
      // Per-body accumulators W (21) / U (6).  Box bodies are written by their 8-lane segment (zeros when the segment
      // has no active corner), capsule-only bodies are zeroed here and filled below; no atomics anywhere, so the
      // result does not depend on scheduling.
      if (sph_any) {
        for (unsigned mk = conmask & ~M.box_body_mask; mk; mk &= mk - 1) {
          const int b = __ffs(mk) - 1;
          for (int i = l; i < 24; i += G) E.W[b][i] = 0.f;
          if (l < 8) E.U[b][l] = 0.f;
        }
      }
#pragma unroll
      for (int ps = 0; ps < kNPass; ps++) {
        if (ps == 1 && !sph_any) continue;
        // wrench-space Hessian of this contact's active pyramid rows: w = (P x d, d), W += D w w^T, U += D aref w
        float wv[28];
#pragma unroll
        for (int i = 0; i < 28; i++) wv[i] = 0.f;
        const unsigned bt = cact[ps] ? AS.bits[ps] : 0u;
        if (bt) {
          const float D = cD[ps], mu = cmu[ps];
          const float s0 = (bt & 1u) ? 1.f : 0.f, s1 = (bt & 2u) ? 1.f : 0.f;
          const float s2 = (bt & 4u) ? 1.f : 0.f, s3 = (bt & 8u) ? 1.f : 0.f;
          const float Qxx = D * mu * mu * (s0 + s1), Qyy = D * mu * mu * (s2 + s3), Qzz = D * (s0 + s1 + s2 + s3);
          const float Qxz = D * mu * (s0 - s1), Qyz = D * mu * (s2 - s3);
          const float Px = cPx[ps], Py = cPy[ps], Pz = cPz[ps];
          // X = [P]x Q with Q rows (Qxx,0,Qxz) (0,Qyy,Qyz) (Qxz,Qyz,Qzz);  Nn row i = P x X_i
          const float X00 = Py * Qxz, X01 = -Pz * Qyy + Py * Qyz, X02 = -Pz * Qyz + Py * Qzz;
          const float X10 = Pz * Qxx - Px * Qxz, X11 = -Px * Qyz, X12 = Pz * Qxz - Px * Qzz;
          const float X20 = -Py * Qxx, X21 = Px * Qyy, X22 = -Py * Qxz + Px * Qyz;
          float t0, t1, t2;
          cross3(wv[sym6(0, 0)], wv[sym6(0, 1)], wv[sym6(0, 2)], Px, Py, Pz, X00, X01, X02);
          cross3(t0, wv[sym6(1, 1)], wv[sym6(1, 2)], Px, Py, Pz, X10, X11, X12);
          cross3(t1, t2, wv[sym6(2, 2)], Px, Py, Pz, X20, X21, X22);
          (void)t0; (void)t1; (void)t2;
          wv[sym6(0, 3)] = X00; wv[sym6(0, 4)] = X01; wv[sym6(0, 5)] = X02;
          wv[sym6(1, 3)] = X10; wv[sym6(1, 4)] = X11; wv[sym6(1, 5)] = X12;
          wv[sym6(2, 3)] = X20; wv[sym6(2, 4)] = X21; wv[sym6(2, 5)] = X22;
          wv[sym6(3, 3)] = Qxx; wv[sym6(3, 5)] = Qxz; wv[sym6(4, 4)] = Qyy; wv[sym6(4, 5)] = Qyz;
          wv[sym6(5, 5)] = Qzz;
          const float a0 = s0 * car[ps][0], a1 = s1 * car[ps][1], a2 = s2 * car[ps][2], a3 = s3 * car[ps][3];
          const float gx = D * mu * (a0 - a1), gy = D * mu * (a2 - a3), gz = D * (a0 + a1 + a2 + a3);
          cross3(wv[21], wv[22], wv[23], Px, Py, Pz, gx, gy, gz);
          wv[24] = gx; wv[25] = gy; wv[26] = gz;
        }
        if (ps == 0) {
          // pass 0 holds the box corners: the 8 lanes of a segment belong to one box = one body -> butterfly sum,
          // then the segment stores its body's accumulators
#pragma unroll
          for (int i = 0; i < 27; i++) {
            float t = wv[i];
            t += __shfl_xor_sync(kFull, t, 1);
            t += __shfl_xor_sync(kFull, t, 2);
            t += __shfl_xor_sync(kFull, t, 4);
            wv[i] = t;
          }
          if (l < M.nbox_cand) {
            const int sl = wl & 7, b = cbody[0];
#pragma unroll
            for (int i = 0; i < 27; i++) {
              if ((i & 7) == sl) {
                if (i < 21) E.W[b][i] = wv[i]; else E.U[b][i - 21] = wv[i];
              }
            }
          }
        } else {
          // capsule end spheres: rare; added one contact at a time in lane order (deterministic)
          __syncwarp();
          for (unsigned sm = __ballot_sync(kFull, bt != 0u); sm; sm &= sm - 1) {
            if (wl == __ffs(sm) - 1) {
              float* Wb = E.W[cbody[ps]];
              float* Ub = E.U[cbody[ps]];
#pragma unroll
              for (int i = 0; i < 21; i++) Wb[i] += wv[i];
#pragma unroll
              for (int i = 0; i < 6; i++) Ub[i] += wv[21 + i];
            }
            __syncwarp();
          }
        }
      }
      __syncwarp();
    }
    // Hessian column
#pragma unroll
    for (int r = 0; r < NV; r++) H[r] = E.Mc[r * G + l];
    H[NV] = rhs0;
    if (constrained && L.isdof) {
      for (unsigned mk = conmask; mk; mk &= mk - 1) {
        const int b = __ffs(mk) - 1;
        const unsigned supp = M.body_supp[b];
        if ((supp >> l) & 1u) {
          float Wl[24];
#pragma unroll
          for (int i = 0; i < 24; i += 4) {
            const float4 t = *reinterpret_cast<const float4*>(&E.W[b][i]);
            Wl[i] = t.x; Wl[i + 1] = t.y; Wl[i + 2] = t.z; Wl[i + 3] = t.w;
          }
          const float Sv[6] = {S.w0, S.w1, S.w2, S.v0, S.v1, S.v2};
          float y[6];
#pragma unroll
          for (int i = 0; i < 6; i++) {
            float t = 0.f;
#pragma unroll
            for (int j = 0; j < 6; j++) t = fmaf(Wl[sym6(i, j)], Sv[j], t);
            y[i] = t;
          }
          const Vec6 yv = {y[0], y[1], y[2], y[3], y[4], y[5]};
          H[NV] += dot6(S, ld6(E.U[b]));
#pragma unroll
          for (int r = 0; r < NV; r++) {
            if ((supp >> r) & 1u) H[r] += dot6(ld6(E.S[r]), yv);
          }
        }
      }
      if (AS.lbit) {
#pragma unroll
        for (int r = 0; r < NV; r++)
          if (r == l) H[r] += lD;
        H[NV] += lD * lsg * laref;
      }
    }
    a = ldl_solve_cols<NV, G>(H, l);
    if (!L.isdof) a = 0.f;
    if (L.isdof) E.acc[l] = a;
    if (!constrained) break;
    __syncwarp();
    // ---- re-evaluate the rows at the new qacc: J_i a = w_i . (S_b a) ----
    if (L.isbody && ((conmask >> l) & 1u)) {
      Vec6 Tb = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      for (unsigned mk = M.body_supp[l]; mk;) {
        int i0, i1; bool two;
        pop2(mk, i0, i1, two);
        const Vec6 s0 = ld6(E.S[i0]), s1 = ld6(E.S[i1]);
        const float a0 = E.acc[i0], a1 = two ? E.acc[i1] : 0.f;
        axpy6(Tb, a0, s0);
        axpy6(Tb, a1, s1);
      }
      st6(E.T[l], Tb);
    }
    __syncwarp();
    bool changed = false;
#pragma unroll
    for (int ps = 0; ps < kNPass; ps++) {
      if (cact[ps]) {
        const Vec6 Tb = ld6(E.T[cbody[ps]]);
        float ux, uy, uz;
        cross3(ux, uy, uz, Tb.w0, Tb.w1, Tb.w2, cPx[ps], cPy[ps], cPz[ps]);
        ux += Tb.v0; uy += Tb.v1; uz += Tb.v2;
        const float mu = cmu[ps];
        const unsigned nb = ((uz + mu * ux - car[ps][0] < 0.f) ? 1u : 0u) | ((uz - mu * ux - car[ps][1] < 0.f) ? 2u : 0u) |
                            ((uz + mu * uy - car[ps][2] < 0.f) ? 4u : 0u) | ((uz - mu * uy - car[ps][3] < 0.f) ? 8u : 0u);
        changed = changed || (nb != AS.bits[ps]);
        AS.bits[ps] = nb;
      }
    }
    {
      const bool nl = (lsg != 0.f) && (lsg * a - laref < 0.f);
      changed = changed || (nl != AS.lbit);
      AS.lbit = nl;
    }
    if (!__any_sync(kFull, changed)) break;
    if (it == kMaxSolverIter - 1 && env_any(changed, L.emask)) cnt.capped++;
  }
  AS.prev_act = (cact[0] ? 1u : 0u) | (cact[1] ? 2u : 0u);
  AS.prev_lim = lsg != 0.f;
  if (DBG) {
    dbg[(2 + NV) * 32 + l] = a;
    if (l == 0) {
      dbg[(3 + NV) * 32 + 0] = zO;
      dbg[(3 + NV) * 32 + 1] = (float)__popc(conmask);
    }
    int nc = 0;
#pragma unroll
    for (int ps = 0; ps < kNPass; ps++) nc += cact[ps] ? 1 : 0;
    dbg[(4 + NV) * 32 + l] = (float)nc;
  }
}

// splitmix64 finaliser: counter-based generator for the RSI draws (stateless in (seed, env, reset#))
__device__ __forceinline__ unsigned long long mix64(unsigned long long x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

struct Cursor {
  int i_step, pos, count, ep_dur, rsi_step, n_det, resets, flags;
};

// Deterministic initialisation (evaluation, straight:237-265) leaves the reference reading the table and the length
// of mocap step 0 (`reset()` aliases `_qpos_full = data[0]`, :161-167, and nothing re-points it) while `_i_step` and
// `_step` already name step n; the first step transition repairs it.  flags bit 2: still reading step 0's table;
// bit 3: the episode began with a deterministic init, so the in-place COM-Z shift went to step 0's table.
constexpr int kFlagReadStep0 = 4, kFlagDetEpisode = 8;
__device__ __forceinline__ int data_step(const Cursor& c) { return (c.flags & kFlagReadStep0) ? 0 : c.i_step; }
__device__ __forceinline__ int shift_step(const Cursor& c) { return (c.flags & kFlagDetEpisode) ? 0 : c.rsi_step; }

// StraightWalkingTrajectories.next (straight:141-159, :322-348) / BaseReferenceTrajectories.next (base:95-103)
__device__ __forceinline__ void cursor_next(const DevModel& M, const StepArgs& A, Cursor& c, float& dist) {
  const int len = A.step_len[data_step(c)];
  c.pos += M.increment;
  if (M.cursor_mode == DRL_CURSOR_STEPWISE) {
    const int dif = c.pos - len + 1;
    if (dif > 0) {
      if (c.i_step >= M.n_steps - 1) {
        c.i_step = A.left_step[c.i_step] ? 0 : 1;
      } else {
        c.i_step += 1;
        c.count += 1;
      }
      dist = A.step_last_comx[c.rsi_step];   // Q2: `_step` stays the RSI step
      c.pos = dif;
      c.flags &= ~kFlagReadStep0;
    }
  } else {
    if (c.pos >= len - 1) c.pos = 0;
  }
}

// reference sample for this lane's dof at the cursor, with the per-episode COM-X / COM-Z adjustments
__device__ __forceinline__ void ref_lookup(const DevModel& M, const StepArgs& A, const Cursor& c, float dist,
                                           float zoff, int l, int G, bool isdof, float& rq, float& rv) {
  rq = 0.f; rv = 0.f;
  if (!isdof) return;
  const int ds = data_step(c);
  const size_t row = (size_t)(A.step_off[ds] + c.pos) * (size_t)(2 * G);
  rq = A.ref[row + l];
  rv = A.ref[row + G + l];
  if (M.cursor_mode == DRL_CURSOR_STEPWISE) {
    if (l == 0) rq += dist;
    if (l == M.com_z_col && ds == shift_step(c)) rq -= zoff;
  } else {
    if (l == M.com_z_col) rq -= zoff;
  }
}

// sums / minima over the G lanes of an environment (xor offsets below G never leave the group)
template <int G>
__device__ __forceinline__ float group_sum(float x) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) x += __shfl_xor_sync(kFull, x, o);
  return x;
}
template <int G>
__device__ __forceinline__ float group_min(float x) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) x = fminf(x, __shfl_xor_sync(kFull, x, o));
  return x;
}

// desired walking velocity vector (straight:417-422,479-480 / loco3d:51-97)
__device__ __forceinline__ void desired_velocity(const DevModel& M, const StepArgs& A, const Cursor& c, float& d0,
                                                 float& d1) {
  d1 = 0.f;
  if (A.speed_profile != nullptr) {            // mimic_env.py:406-408 (scalar profile; second component stays 0)
    d0 = A.speed_profile[c.ep_dur % A.speed_profile_len];
  } else if (M.cursor_mode == DRL_CURSOR_STEPWISE) {
    int idx = c.i_step - c.count + 1;
    d0 = A.step_vel[idx < 0 ? 0 : idx];
  } else {
    const int len = A.step_len[0];
    int end = c.pos + M.des_vel_window;
    if (end > len - 1) end = len - 1;
    const int n = end - c.pos;
    if (n <= 0) {
      d0 = d1 = nanf("");   // mean of an empty slice (Q23)
    } else {
      d0 = (float)((A.des_vel_prefix[2 * end] - A.des_vel_prefix[2 * c.pos]) / (double)n);
      d1 = (float)((A.des_vel_prefix[2 * end + 1] - A.des_vel_prefix[2 * c.pos + 1]) / (double)n);
    }
  }
}

// _get_obs (mimic_env.py:403-437) into E.obsbuf (unmirrored); returns phase / des_vel for the extras
template <int G>
__device__ __forceinline__ void build_obs(const DevModel& M, const StepArgs& A, EnvSmem<G>& E, const LaneConst& L,
                                          const Cursor& c, float q, float v, float& phase0, float& des0) {
  const int l = L.l;
  float d0, d1;
  desired_velocity(M, A, c, d0, d1);
  des0 = d0;
  const int np = M.n_phase_obs, nd = M.n_des_vel;
  phase0 = 0.f;
  if (M.phase_mode == DRL_PHASE_FROM_CURSOR) {
    phase0 = (float)c.pos / (float)A.step_len[data_step(c)];
    if (l == 0) E.obsbuf[0] = phase0;
  } else if (L.isdof) {
    for (int k = 0; k < M.n_phase_joints; k++)
      if (M.phase_joints[k] == l) {
        E.obsbuf[2 * k] = atan2f(v, -q) * 0.31830988618379067154f;
        E.obsbuf[2 * k + 1] = sqrtf(q * q + v * v) / 5.f;
      }
  }
  if (l == 0) {
    E.obsbuf[np] = d0;
    if (nd > 1) E.obsbuf[np + 1] = d1;
  }
  if (L.isdof) {
    if (l >= 1) E.obsbuf[np + nd + l - 1] = q;
    E.obsbuf[np + nd + M.nv - 1 + l] = v;
  }
  __syncwarp();
}

// obs (optionally mirrored, mimic_env.py:440-480) from E.obsbuf to global memory; `enable` predicates the stores
template <int G>
__device__ __forceinline__ void write_obs(const DevModel& M, EnvSmem<G>& E, int l, bool mirror, bool enable,
                                          float* __restrict__ dst) {
  if (enable)
    for (int k = l; k < M.obs_dim; k += G)
      dst[k] = mirror ? M.mirror_obs_sign[k] * E.obsbuf[M.mirror_obs_idx[k]] : E.obsbuf[k];
  __syncwarp();
}

// MimicEnv.reset_model (mimic_env.py:526-572): RSI, ground-contact shift, refs.next().  Warp-uniform: every lane
// computes a reset, the caller commits it only for environments that need one.
template <int G>
__device__ __noinline__ void reset_env(const DevModel& M, const StepArgs& A, EnvSmem<G>& E, const LaneConst& L,
                                       int env, Cursor& c, float& q, float& v, float& dist, float& zoff) {
  const int l = L.l;
  c.ep_dur = 0;
  if (A.eval_mode || A.speed_profile != nullptr) {   // mimic_env.py:536-537; straight:237-265 / base:69-77
    if (M.cursor_mode == DRL_CURSOR_STEPWISE) {
      c.i_step = c.n_det;
      c.pos = (3 * A.step_len[c.i_step]) / 4;
      c.n_det += 1;
      if (c.n_det >= M.eval_n_times) c.n_det = 0;
      c.flags |= kFlagReadStep0 | kFlagDetEpisode;
    } else {
      c.i_step = 0; c.pos = 0;
    }
  } else if (A.inj_istep != nullptr && A.inj_pos != nullptr && A.inj_pos[env] >= 0) {
    c.flags &= ~(kFlagReadStep0 | kFlagDetEpisode);
    c.i_step = M.cursor_mode == DRL_CURSOR_STEPWISE ? A.inj_istep[env] : 0;
    c.pos = A.inj_pos[env];
  } else {                                     // straight:460-474 / base:79-85
    c.flags &= ~(kFlagReadStep0 | kFlagDetEpisode);
    const unsigned long long gid = (unsigned long long)(M.env_id_offset + env);
    const unsigned long long r = mix64(mix64(M.seed ^ (gid * 0xD1342543DE82EF95ull)) + (unsigned long long)c.resets);
    c.i_step = (int)(((r & 0xFFFFFFFFull) * (unsigned long long)M.n_steps) >> 32);
    c.pos = (int)(((r >> 32) * (unsigned long long)A.step_len[c.i_step]) >> 32);
  }
  c.resets += 1;
  c.rsi_step = c.i_step;
  dist = 0.f;
  zoff = 0.f;
  float rq, rv;
  ref_lookup(M, A, c, dist, zoff, l, G, L.isdof, rq, rv);
  q = rq; v = rv;
  // set_state + sim.forward: lowest foot-corner site (mimic_env.py:546-559)
  {
    float s = q - L.ref, cc = 1.f;
    if (L.isdof && L.type == 1) sincosf(L.sign * (q - L.ref), &s, &cc);
    if (L.isdof) *reinterpret_cast<float2*>(&E.cssn[l][0]) = make_float2(cc, s);
  }
  __syncwarp();
  float zO = M.root_z0;
  for (int j = 0; j < M.nslide; j++) zO = fmaf(M.dof_slide_z[j], E.cssn[j][1], zO);
  tree_kinematics<G>(M, E, l);
  float sz = 3.0e38f;
  for (int s = l; s < M.nsite; s += G) {
    const float* R = E.bodyR[M.site_body[s]];
    sz = fminf(sz, zO + R[11] + R[6] * M.site_pos[s][0] + R[7] * M.site_pos[s][1] + R[8] * M.site_pos[s][2]);
  }
  const float lowest = group_min<G>(sz);
  if (l == M.com_z_dof) q -= lowest;
  zoff = lowest;                               // refs.adjust_COM_Z_pos(lowest)
  cursor_next(M, A, c, dist);                  // mimic_env.py:568
  __syncwarp();
}

template <int NV, int G, bool RK4, bool DBG>
__global__ void __launch_bounds__(128, 4) mimic_step_kernel(const StepArgs A, const int do_reset_only) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  DevModel& M = *reinterpret_cast<DevModel*>(smem_raw);
  constexpr int kModelBytes = (sizeof(DevModel) + 15) / 16 * 16;
  EnvSmem<G>* envs = reinterpret_cast<EnvSmem<G>*>(smem_raw + kModelBytes);
  {
    const int4* src = reinterpret_cast<const int4*>(A.model);
    int4* dst = reinterpret_cast<int4*>(smem_raw);
    for (int i = threadIdx.x; i < (int)(sizeof(DevModel) / 16); i += blockDim.x) dst[i] = src[i];
  }
  __syncthreads();
  const int epb = blockDim.x / G;
  const int eib = threadIdx.x / G;
  const int env_raw = blockIdx.x * epb + eib;
  // a warp whose second environment is past the end recomputes the last valid one with all stores disabled
  const bool live = env_raw < A.num_envs;
  const int env = live ? env_raw : A.num_envs - 1;
  // (warps entirely past the end keep running on the clamped environment: the CTA-wide barriers in the substep loop
  //  need every warp, and only the last CTA has such warps)
  EnvSmem<G>& E = envs[eib];
  LaneConst L;
  L.l = threadIdx.x % G;
  const int l = L.l;
  L.emask = (G == 32) ? kFull : (0xFFFFu << (16 * ((threadIdx.x & 31) / 16)));
  L.isdof = l < M.nv;
  L.isbody = l < M.nb;
  const int jd = L.isdof ? l : 0;
  L.body = M.dof_body[jd]; L.type = M.dof_type[jd]; L.limited = M.dof_limited[jd]; L.last = M.dof_last[jd];
  L.sign = M.dof_sign[jd]; L.ref = M.dof_ref[jd]; L.damping = M.dof_damping[jd]; L.armature = M.dof_armature[jd];
  L.lo = M.dof_lo[jd]; L.hi = M.dof_hi[jd]; L.invw = M.dof_invw[jd];
  L.anc = M.dof_anc[jd]; L.desc = M.dof_desc[jd]; L.subb = M.dof_subbodies[jd];

  float* sf = A.state_f + (size_t)env * (4 * G);
  int* si = A.state_i + (size_t)env * kCurCount8;
  float q = sf[l], v = sf[G + l], a = sf[2 * G + l];
  Cursor c;
  c.i_step = si[kCurIstep]; c.pos = si[kCurPos]; c.count = si[kCurCount]; c.ep_dur = si[kCurEpDur];
  c.rsi_step = si[kCurRsiStep]; c.n_det = si[kCurNDet]; c.resets = si[kCurResets]; c.flags = si[kCurFlags];
  float dist = sf[3 * G + kMiscDist], zoff = sf[3 * G + kMiscZoff];
  Counters cnt = {0, 0, 0};
  // active set of the last evaluation of the previous step (bits 0-3 / 4-7: pyramid rows of the two candidates,
  // 8-9: candidates in contact, 10: limit row active, 11: limit violated)
  int* sa = A.state_as + (size_t)env * G;
  ActiveSet AS;
  {
    const unsigned pk = (unsigned)sa[l];
    AS.bits[0] = pk & 0xFu; AS.bits[1] = (pk >> 4) & 0xFu; AS.prev_act = (pk >> 8) & 3u;
    AS.lbit = (pk >> 10) & 1u; AS.prev_lim = (pk >> 11) & 1u;
  }
  float* obs_out = A.obs + (size_t)env * M.obs_dim;

  if (do_reset_only) {
    // VecEnv.reset(): (masked) reset of the episode state
    const bool want = live && (A.reset_mask == nullptr || A.reset_mask[env]);
    if (!__any_sync(kFull, want)) return;
    reset_env<G>(M, A, E, L, env, c, q, v, dist, zoff);
    float ph, dv;
    build_obs<G>(M, A, E, L, c, q, v, ph, dv);
    const bool left = M.mirror_policy && A.left_step[c.i_step];
    write_obs<G>(M, E, l, left, want, obs_out);
    if (want) {
      sf[l] = q; sf[G + l] = v; sf[2 * G + l] = 0.f;
      if (l == 0) {
        sf[3 * G + kMiscDist] = dist; sf[3 * G + kMiscZoff] = zoff;
        sf[3 * G + kMiscWalked] = 0.f; sf[3 * G + kMiscEpRet] = 0.f; sf[3 * G + kMiscEpTor] = 0.f;
        sf[3 * G + kMiscPrevPos] = 1.f; sf[3 * G + kMiscPrevVel] = 1.f; sf[3 * G + kMiscPrevCom] = 1.f;
        si[kCurIstep] = c.i_step; si[kCurPos] = c.pos; si[kCurCount] = c.count; si[kCurEpDur] = 0;
        si[kCurRsiStep] = c.rsi_step; si[kCurNDet] = c.n_det; si[kCurResets] = c.resets; si[kCurFlags] = c.flags;
      }
    }
    return;
  }

  // ---- actions -> joint torques (mimic_env.py:170-192, :483-489; MuJoCo ctrl/force clamps) ---------------------
  const bool left0 = M.mirror_policy && A.left_step[c.i_step];
  float force = 0.f;
  {
    float sc = 0.f;
    if (l < M.nu) {
      float act = A.actions[(size_t)env * M.act_dim + l];
      act = fminf(fmaxf(act, -1.f), 1.f);
      sc = act > 0.f ? __fmul_rn(act, M.act_chi[l]) : __fmul_rn(fabsf(act), M.act_clo[l]);
    }
    {
      const int src = l < M.nu ? M.mirror_act_idx[l] : 0;
      const float o = __shfl_sync(kFull, sc, src, G);
      if (left0) sc = l < M.nu ? M.mirror_act_sign[l] * o : 0.f;
    }
    E.tau[l] = 0.f;
    __syncwarp();
    if (l < M.nu) {
      const float cc = fminf(fmaxf(sc, M.act_clo[l]), M.act_chi[l]);
      force = fminf(fmaxf(M.act_gear[l] * cc, M.act_flo[l]), M.act_fhi[l]);
      E.tau[M.act_dof[l]] = M.act_gear[l] * force;
    }
    __syncwarp();
  }
  const float tau = E.tau[l];
  const float mean_abs_torque = group_sum<G>(l < M.nu ? fabsf(force) : 0.f) / (float)M.nu;

  // ---- physics: frame_skip x (RK4 | semi-implicit Euler) ---------------------------------------------------------
  const float h = M.timestep;
  bool bad = false;      // MuJoCo's mj_checkPos / mj_checkVel: non-finite or huge state -> MujocoException
  float* dbgp = (DBG && A.debug) ? A.debug + (size_t)env * 32 * 40 : nullptr;
  for (int sub = 0; sub < A.frame_skip; sub++) {
    {
      // (collectives must be reached by every lane: no short-circuit around env_any)
      const bool blown = env_any(L.isdof && (!(fabsf(q) <= 1e10f) || !(fabsf(v) <= 1e10f)), L.emask);
      bad = bad || blown;
    }
    if (bad) { q = L.ref; v = 0.f; a = 0.f; }      // park the environment on a harmless state; it resets below
    if (RK4) {
      const float q0 = q, v0 = v;
      float accq = 0.f, accv = 0.f;
#pragma unroll 1
      for (int st = 0; st < 4; st++) {
        // keep the warps of a CTA within one evaluation of each other: they then share instruction-cache lines
        // (the dynamics evaluation is ~80 KB of straight-line code); measured +3 %
        // (one barrier per substep instead of per stage loses the gain; a second barrier inside the evaluation adds none)
        __syncthreads();
        forward_dynamics<NV, G, DBG>(M, E, L, q, v, tau, a, AS, cnt, dbgp);
        const float bw = (st == 0 || st == 3) ? (1.f / 6.f) : (1.f / 3.f);
        accq = fmaf(bw, v, accq);
        accv = fmaf(bw, a, accv);
        if (st < 3) {
          const float aw = st == 2 ? 1.f : 0.5f;
          const float vn = fmaf(h * aw, a, v0);
          q = fmaf(h * aw, v, q0);
          v = vn;
        }
      }
      q = fmaf(h, accq, q0);
      v = fmaf(h, accv, v0);
    } else {
      forward_dynamics<NV, G, DBG>(M, E, L, q, v, tau, a, AS, cnt, dbgp);
      // mj_Euler with implicit joint damping: (M + h B) a' = M a   [M a = tau_total + J'f]
      // the column of M is rebuilt from E.S / E.Fd which are still valid
      __syncwarp();   // E.acc holds the constrained qacc of every dof
      float Hc[NV + 1];
      float Ma = 0.f;
      {
        // E.Mc still holds this lane's column from the evaluation above
#pragma unroll
        for (int r = 0; r < NV; r++) {
          const float m = E.Mc[r * G + l];
          Ma = fmaf(m, E.acc[r], Ma);
          Hc[r] = m + (r == l ? h * L.damping : 0.f);
        }
      }
      Hc[NV] = Ma;
      float an = ldl_solve_cols<NV, G>(Hc, l);
      if (!L.isdof) an = 0.f;
      v = fmaf(h, an, v);
      q = fmaf(h, v, q);
    }
  }
  {
    const bool blown = env_any(L.isdof && (!(fabsf(q) <= 1e10f) || !(fabsf(v) <= 1e10f)), L.emask);
    bad = bad || blown;
  }
  if (bad) { q = L.ref; v = 0.f; a = 0.f; }

  // ---- environment logic (computed for every lane; the blow-up path overrides the outcome) ------------------------
  float walked = sf[3 * G + kMiscWalked];
  float ep_ret = sf[3 * G + kMiscEpRet], ep_tor = sf[3 * G + kMiscEpTor];
  float pos_rew = sf[3 * G + kMiscPrevPos], vel_rew = sf[3 * G + kMiscPrevVel], com_rew = sf[3 * G + kMiscPrevCom];
  float reward, phase0 = 0.f, des0 = 0.f;
  bool done;
  const int ep_dur_before = c.ep_dur;
  cursor_next(M, A, c, dist);                                   // mimic_env.py:96
  // Monitor: refs._pos after the first step of an episode (monitor_wrapper.py:91-93), kept in the upper half of flags
  if (ep_dur_before == 0) c.flags = (c.flags & 0xFFFF) | (c.pos << 16);
  if (A.playback) {                                             // set_joint_kinematics_in_sim (mimic_env.py:273-293)
    float rq, rv;
    ref_lookup(M, A, c, dist, zoff, l, G, L.isdof, rq, rv);
    if (L.isdof) { q = rq; v = rv; }
  }
  build_obs<G>(M, A, E, L, c, q, v, phase0, des0);               // mimic_env.py:99
  c.ep_dur += 1;                                                // mimic_env.py:106
  {                                                             // mimic_env.py:131-139
    const float vx = __shfl_sync(kFull, v, 0, G), vy = __shfl_sync(kFull, v, 1, G);
    const float cx = fminf(fmaxf(vx, -5.5f), 5.5f), cy = fminf(fmaxf(vy, -5.5f), 5.5f);
    walked += sqrtf(cx * cx + cy * cy) * M.ctrl_freq_inv;
  }
  const float comz = __shfl_sync(kFull, q, M.com_z_dof, G);
  const bool timeout = c.ep_dur >= M.ep_dur_max;
  float rq, rv;
  ref_lookup(M, A, c, dist, zoff, l, G, L.isdof, rq, rv);
  // do_terminate_early (mimic_env.py:652-702, 3D branch): the reference computes nothing from it in step(); here the
  // reasons are counted and, when configured, end the episode
  bool et_low, et_trunk, et_drunk;
  {
    const int t0 = M.trunk_dof0;
    const float front_dev = fabsf(__shfl_sync(kFull, q - rq, t0, G));
    const float sag = __shfl_sync(kFull, q, t0 + 1, G), comy = __shfl_sync(kFull, q, 1, G);
    et_low = comz < 0.75f;
    et_trunk = (sag > 0.3f || sag < -0.05f) || front_dev > 0.2f;
    et_drunk = fabsf(comy) > 0.2f;
    if (A.playback) et_low = et_trunk = et_drunk = false;
  }
  const bool et_done = M.early_termination && (et_low || et_trunk || et_drunk);
  done = (comz < M.fall_z) || timeout || et_done;               // mimic_env.py:113-120
  {
    const float dq = L.isdof ? q - rq : 0.f, dv = L.isdof ? v - rv : 0.f;
    const bool iscom = (M.com_mask >> l) & 1u;
    const float sp = group_sum<G>(iscom ? 0.f : dq * dq);
    const float sv = group_sum<G>(iscom ? 0.f : dv * dv);
    const float sc = group_sum<G>(iscom ? dq * dq : 0.f);
    if (!done && !bad) {
      pos_rew = expf(-3.f * sp);                                // mimic_env.py:592-622
      vel_rew = expf(-0.05f * sv);
      com_rew = expf(-16.f * sc);
    }
  }
  if (bad) {
    // MujocoException path (mimic_env.py:86-91): reset, reward 0, done; the VecEnv then resets again (Q19) — here a
    // single reset serves both, and the terminal observation is the post-reset observation as in the reference
    done = true;
    reward = 0.f;
  } else if (done) {
    reward = timeout ? 0.f : -0.f;                              // _get_ET_reward always evaluates to +-0 (Q1)
  } else {
    reward = (M.w_pos * pos_rew + M.w_vel * vel_rew + M.w_com * com_rew) * M.rew_scale + M.alive_bonus;
  }
  {
    const bool left1 = M.mirror_policy && A.left_step[c.i_step];
    float* dst = (done && A.terminal_obs) ? A.terminal_obs + (size_t)env * M.obs_dim : obs_out;
    write_obs<G>(M, E, l, left1, live && !bad, dst);
  }
  // ---- Monitor.step (monitor_wrapper.py:88-166) --------------------------------------------------------------------
  ep_ret += reward;
  ep_tor += mean_abs_torque;
  double* sd = A.state_d + (size_t)env * 4;
  if (l == 0 && live) {
    sd[0] += (double)pos_rew; sd[1] += (double)vel_rew; sd[2] += (double)com_rew; sd[3] += 1.0;
    atomicAdd(&A.stats[DRL_STAT_ENV_STEPS], 1.0);
    atomicAdd(&A.stats[DRL_STAT_POS_REW_SUM], (double)pos_rew);
    atomicAdd(&A.stats[DRL_STAT_VEL_REW_SUM], (double)vel_rew);
    atomicAdd(&A.stats[DRL_STAT_COM_REW_SUM], (double)com_rew);
    atomicAdd(&A.stats[DRL_STAT_REW_STEPS], 1.0);
    atomicAdd(&A.stats[DRL_STAT_ABS_TORQUE_SUM], (double)mean_abs_torque);
    atomicAdd(&A.stats[DRL_STAT_SOLVER_ITERS], (double)cnt.iters);
    atomicAdd(&A.stats[DRL_STAT_DYN_EVALS], (double)cnt.evals);
    if (cnt.capped) atomicAdd(&A.stats[DRL_STAT_SOLVER_CAPPED], (double)cnt.capped);
    if (!bad) {
      if (et_low) atomicAdd(&A.stats[DRL_STAT_ET_COM_LOW], 1.0);
      if (et_trunk) atomicAdd(&A.stats[DRL_STAT_ET_TRUNK], 1.0);
      if (et_drunk) atomicAdd(&A.stats[DRL_STAT_ET_DRUNK], 1.0);
    }
    if (A.extras) {
      float* ex = A.extras + (size_t)env * 16;
      ex[0] = pos_rew; ex[1] = vel_rew; ex[2] = com_rew; ex[3] = walked; ex[4] = mean_abs_torque;
      ex[5] = des0; ex[6] = phase0; ex[7] = zoff;
    }
  }
  if (done && l == 0 && live) {
    const int ep_len = bad ? ep_dur_before + 1 : c.ep_dur;
    float* ms = sf + 3 * G;
    const int fl = c.flags;
    auto smooth = [&](int slot, int bit, float nv, float f) {
      ms[slot] = (fl >> bit) & 1 ? f * nv + (1.f - f) * ms[slot] : nv;   // utils.py:312-329
    };
    if (ep_len > 1) { smooth(kMiscMeanRewSm, 0, ep_ret / (float)(ep_len - 1), 0.9f); c.flags |= 1; }
    smooth(kMiscPosSm, 1, (float)(sd[0] / sd[3]), 0.9f);
    smooth(kMiscVelSm, 1, (float)(sd[1] / sd[3]), 0.9f);
    smooth(kMiscComSm, 1, (float)(sd[2] / sd[3]), 0.9f);
    smooth(kMiscEpRetSm, 1, ep_ret, 0.25f);
    smooth(kMiscEpLenSm, 1, (float)ep_len, 0.75f);
    smooth(kMiscTorSm, 1, ep_tor / (float)ep_len, 0.75f);
    c.flags |= 2;
    ms[kMiscMoved] = walked;
    atomicAdd(&A.stats[DRL_STAT_EPISODES], 1.0);
    atomicAdd(&A.stats[DRL_STAT_EP_LEN_SUM], (double)ep_len);
    atomicAdd(&A.stats[DRL_STAT_EP_RET_SUM], (double)ep_ret);
    if (ep_len > 1) atomicAdd(&A.stats[DRL_STAT_EP_MEAN_REW_SUM], (double)(ep_ret / (float)(ep_len - 1)));
    atomicAdd(&A.stats[DRL_STAT_MOVED_DISTANCE_SUM], (double)walked);
    atomicAdd(&A.stats[bad ? DRL_STAT_BLOWUPS : (timeout ? DRL_STAT_TIMEOUTS : DRL_STAT_FALLS)], 1.0);
    if (A.ring_cap > 0) {
      const unsigned long long slot = atomicAdd(A.ring_head, 1ull) % (unsigned long long)A.ring_cap;
      A.ring_len[slot] = ep_len;
      A.ring_ret[slot] = ep_ret;
      A.ring_rsi_pos[slot] = (int)((unsigned)c.flags >> 16);           // monitor_wrapper.py:91-93
      A.ring_et_pos[slot] = c.pos;                                      // :104-107
      A.ring_difficult[slot] = (float)ep_len < ms[kMiscEpLenSm] * 0.75f ? 1 : 0;   // :123-124 (after the smoothing)
    }
  }
  c.flags = __shfl_sync(kFull, c.flags, 0, G);
  // ---- auto-reset (DummyVecEnv.step_wait), warp-uniform: both environments compute it, only `done` ones commit -----
  if (__any_sync(kFull, done)) {
    Cursor c2 = c;
    float q2 = q, v2 = v, dist2 = dist, zoff2 = zoff;
    reset_env<G>(M, A, E, L, env, c2, q2, v2, dist2, zoff2);
    float ph, dv;
    build_obs<G>(M, A, E, L, c2, q2, v2, ph, dv);
    const bool left2 = M.mirror_policy && A.left_step[c2.i_step];
    write_obs<G>(M, E, l, left2, live && done, obs_out);
    write_obs<G>(M, E, l, left2, live && bad && A.terminal_obs != nullptr,
                 A.terminal_obs ? A.terminal_obs + (size_t)env * M.obs_dim : obs_out);
    if (done) {
      c = c2; q = q2; v = v2; dist = dist2; zoff = zoff2;
      a = 0.f;
      AS.bits[0] = AS.bits[1] = 0u; AS.prev_act = 0u; AS.lbit = AS.prev_lim = false;
      walked = 0.f; ep_ret = 0.f; ep_tor = 0.f;
      pos_rew = vel_rew = com_rew = 1.f;      // get_imitation_reward() inside reset_model (mimic_env.py:562)
    }
  }
  // ---- store -------------------------------------------------------------------------------------------------------
  if (live) {
    sf[l] = q; sf[G + l] = v; sf[2 * G + l] = a;
    sa[l] = (int)(AS.bits[0] | (AS.bits[1] << 4) | (AS.prev_act << 8) | ((AS.lbit ? 1u : 0u) << 10) |
                  ((AS.prev_lim ? 1u : 0u) << 11));
    if (l == 0) {
      A.rew[env] = reward;
      A.done[env] = done ? 1 : 0;
      float* ms = sf + 3 * G;
      ms[kMiscDist] = dist; ms[kMiscZoff] = zoff; ms[kMiscWalked] = walked; ms[kMiscEpRet] = ep_ret;
      ms[kMiscEpTor] = ep_tor; ms[kMiscPrevPos] = pos_rew; ms[kMiscPrevVel] = vel_rew; ms[kMiscPrevCom] = com_rew;
      si[kCurIstep] = c.i_step; si[kCurPos] = c.pos; si[kCurCount] = c.count; si[kCurEpDur] = c.ep_dur;
      si[kCurRsiStep] = c.rsi_step; si[kCurNDet] = c.n_det; si[kCurResets] = c.resets; si[kCurFlags] = c.flags;
    }
  }
}

// gather of the per-env extras (Monitor attributes served through VecEnv.get_attr)
__global__ void extras_kernel(const float* __restrict__ state_f, const float* __restrict__ last, float* out, int n,
                              int G) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * 16) return;
  const int env = i / 16, k = i % 16;
  float val;
  if (k < 8) val = last ? last[env * 16 + k] : 0.f;
  else val = state_f[(size_t)env * 4 * G + 3 * G + k];     // misc slots 8..15 line up with extras 8..15
  if (k == 3) val = state_f[(size_t)env * 4 * G + 3 * G + kMiscWalked];
  if (k == 7) val = state_f[(size_t)env * 4 * G + 3 * G + kMiscZoff];
  out[i] = val;
}

// strided copies between the padded state rows and dense [N][nv] user tensors
__global__ void state_copy_kernel(float* state_f, int* state_i, int* state_as, float* qpos, float* qvel, int* cursor,
                                  int n, int nv, int G, int to_state) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * G) return;
  const int env = i / G, l = i % G;
  float* sf = state_f + (size_t)env * 4 * G;
  if (l < nv) {
    if (to_state) {
      if (qpos) sf[l] = qpos[env * nv + l];
      if (qvel) sf[G + l] = qvel[env * nv + l];
      sf[2 * G + l] = 0.f;
      if (qpos || qvel) state_as[(size_t)env * G + l] = 0;
    } else {
      if (qpos) qpos[env * nv + l] = sf[l];
      if (qvel) qvel[env * nv + l] = sf[G + l];
    }
  }
  if (cursor && l < 4) {
    const int slot = l == 0 ? kCurIstep : (l == 1 ? kCurPos : (l == 2 ? kCurCount : kCurEpDur));
    if (to_state) state_i[env * kCurCount8 + slot] = cursor[env * 4 + l];
    else cursor[env * 4 + l] = state_i[env * kCurCount8 + slot];
  }
}

// ------------------------------------------------------------------------------------------------------------------
// launch helpers used by c_api.cu
// ------------------------------------------------------------------------------------------------------------------
size_t step_smem_bytes(int G, int envs_per_block) {
  const size_t model = (sizeof(DevModel) + 15) / 16 * 16;
  return model + (size_t)envs_per_block * (G == 16 ? sizeof(EnvSmem<16>) : sizeof(EnvSmem<32>));
}

template <int NV, int G, bool RK4, bool DBG>
static cudaError_t launch_one(const StepArgs& a, int reset_only, int block, cudaStream_t st) {
  const int epb = block / G;
  const size_t smem = step_smem_bytes(G, epb);
  auto kern = mimic_step_kernel<NV, G, RK4, DBG>;
  // function attributes are per device: set them once for every device this process launches on
  static unsigned long long attr_done = 0ull;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (!((attr_done >> (dev & 63)) & 1ull)) {
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return e;
    // all of the unified L1/shared array as shared memory: residency is bounded by the per-env scratch, not by L1 hits
    e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    attr_done |= 1ull << (dev & 63);
  }
  const int grid = (a.num_envs + epb - 1) / epb;
  kern<<<grid, block, smem, st>>>(a, reset_only);
  return cudaGetLastError();
}

// instantiated: walker3d (nv 14, 16 lanes) and walker_165cm_65kg (nv 19, 32 lanes), RK4 and Euler; the dump variant
// exists for the Euler kernels only (tests evaluate single forward passes with it)
cudaError_t launch_step(const StepArgs& a, int nv, int G, int rk4, int reset_only, int block, bool debug,
                        cudaStream_t st) {
  if (debug && rk4) return cudaErrorInvalidValue;
  if (nv == 14 && G == 16) {
    if (rk4) return launch_one<14, 16, true, false>(a, reset_only, block, st);
    return debug ? launch_one<14, 16, false, true>(a, reset_only, block, st)
                 : launch_one<14, 16, false, false>(a, reset_only, block, st);
  }
  if (nv == 19 && G == 32) {
    if (rk4) return launch_one<19, 32, true, false>(a, reset_only, block, st);
    return debug ? launch_one<19, 32, false, true>(a, reset_only, block, st)
                 : launch_one<19, 32, false, false>(a, reset_only, block, st);
  }
  return cudaErrorInvalidValue;
}

cudaError_t launch_extras(const float* state_f, const float* last, float* out, int n, int G, cudaStream_t st) {
  const int total = n * 16;
  extras_kernel<<<(total + 255) / 256, 256, 0, st>>>(state_f, last, out, n, G);
  return cudaGetLastError();
}

cudaError_t launch_state_copy(float* state_f, int* state_i, int* state_as, float* qpos, float* qvel, int* cursor, int n,
                              int nv, int G, int to_state, cudaStream_t st) {
  const int total = n * G;
  state_copy_kernel<<<(total + 255) / 256, 256, 0, st>>>(state_f, state_i, state_as, qpos, qvel, cursor, n, nv, G,
                                                         to_state);
  return cudaGetLastError();
}

}  // namespace drl
