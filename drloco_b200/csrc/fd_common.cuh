// fd_common.cuh — small device helpers of the step kernel (internal to libdrloco_b200).
#pragma once
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

// sin and cos of a joint angle.  Cody-Waite reduction by pi/2 in three steps and the minimax polynomials of the
// single-precision math library on [-pi/4, pi/4] (max error ~1 ulp for |x| up to a few thousand radians; joint angles
// are O(1) and a state beyond 1e10 is the blow-up path).  Unlike sincosf there is no large-argument slow path, which
// keeps ~200 cold instructions out of the evaluation loop's instruction-cache footprint.
__device__ __forceinline__ void sincos_joint(float x, float& sn, float& cs) {
  const float k = rintf(x * 0.636619772367581343f);
  float r = fmaf(k, -1.57079601e+00f, x);
  r = fmaf(k, -3.13916473e-07f, r);
  r = fmaf(k, -5.39030253e-15f, r);
  const int q = (int)k;
  const float r2 = r * r;
  float sp = fmaf(r2, -1.9515295891e-4f, 8.3321608736e-3f);
  sp = fmaf(sp, r2, -1.6666654611e-1f);
  sp = fmaf(sp * r2, r, r);
  float cp = fmaf(r2, 2.443315711809948e-5f, -1.388731625493765e-3f);
  cp = fmaf(cp, r2, 4.166664568298827e-2f);
  cp = fmaf(cp, r2, -0.5f);
  cp = fmaf(cp, r2, 1.f);
  float s0 = (q & 1) ? cp : sp, c0 = (q & 1) ? sp : cp;
  sn = (q & 2) ? -s0 : s0;
  cs = ((q + 1) & 2) ? -c0 : c0;
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
  int body, type;      // body and joint type (0 slide, 1 hinge) of this lane's dof; the other per-dof constants are
                       // read from the model block in shared memory where they are used (keeps them out of registers)
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

// does predicate p hold on any lane of this lane's environment?
__device__ __forceinline__ bool env_any(bool p, unsigned emask) { return (__ballot_sync(kFull, p) & emask) != 0u; }

}  // namespace drl
