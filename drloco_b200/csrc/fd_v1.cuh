// fd_v1.cuh — first-generation dynamics evaluation (lane <-> dof / body with per-lane ancestor loops, dense Hessian
// assembly).  Kept selectable (DRLOCO_B200_FD=1) as the A/B baseline of fd_v2.cuh.
#pragma once
#include "fd_common.cuh"

namespace drl {

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

}  // namespace drl
