// Canonical float32 primitives of the STAC solver (device side).
//
// Every operation is written out explicitly (fmaf where a fused multiply-add is meant,
// plain * and + elsewhere) and the translation unit is compiled with -fmad=false, so the
// compiler neither fuses nor splits anything: the arithmetic is the "canonical order"
// specified in DESIGN.md section 4 and reproduced independently by the CPU oracle.
// The formulas restate mujoco.mjx._src.math (rotate, quat_mul, axis_angle_to_quat,
// normalize) -- third-party code that is not part of the reference tree.
#pragma once
#include <cuda_runtime.h>

namespace stacb {

struct V3 { float x, y, z; };
struct Q4 { float w, x, y, z; };

__device__ __forceinline__ V3 mk3(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ Q4 mk4(float w, float x, float y, float z) { Q4 r; r.w = w; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ V3 add3(V3 a, V3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ V3 sub3(V3 a, V3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }

__device__ __forceinline__ float dot3(V3 a, V3 b) { return fmaf(a.z, b.z, fmaf(a.y, b.y, a.x * b.x)); }

__device__ __forceinline__ V3 cross3(V3 a, V3 b) {
  return mk3(fmaf(a.y, b.z, -(a.z * b.y)), fmaf(a.z, b.x, -(a.x * b.z)), fmaf(a.x, b.y, -(a.y * b.x)));
}

// math.rotate: 2*(u.v)u + (s^2 - u.u)v + 2s(u x v)
__device__ __forceinline__ V3 rotate(V3 v, Q4 q) {
  V3 u = mk3(q.x, q.y, q.z);
  float d = dot3(u, v);
  float uu = dot3(u, u);
  float k = fmaf(q.w, q.w, -uu);
  V3 c = cross3(u, v);
  float d2 = 2.0f * d, s2 = 2.0f * q.w;
  return mk3(fmaf(s2, c.x, fmaf(k, v.x, d2 * u.x)), fmaf(s2, c.y, fmaf(k, v.y, d2 * u.y)),
             fmaf(s2, c.z, fmaf(k, v.z, d2 * u.z)));
}

// math.quat_mul (Hamilton, w first)
__device__ __forceinline__ Q4 qmul(Q4 u, Q4 v) {
  Q4 r;
  r.w = fmaf(-u.z, v.z, fmaf(-u.y, v.y, fmaf(-u.x, v.x, u.w * v.w)));
  r.x = fmaf(-u.z, v.y, fmaf(u.y, v.z, fmaf(u.x, v.w, u.w * v.x)));
  r.y = fmaf(u.z, v.x, fmaf(u.y, v.w, fmaf(-u.x, v.z, u.w * v.y)));
  r.z = fmaf(u.z, v.w, fmaf(-u.y, v.x, fmaf(u.x, v.y, u.w * v.z)));
  return r;
}

// math.normalize: x / (n + 1e-6 * (n == 0)), canonically evaluated as x * (1 / d) with ONE IEEE division;
// the reciprocal is returned through *rinv_out (the reverse sweep multiplies by it as well).
__device__ __forceinline__ Q4 normalize4(Q4 q, float *rinv_out) {
  float n2 = fmaf(q.z, q.z, fmaf(q.y, q.y, fmaf(q.x, q.x, q.w * q.w)));
  float n = sqrtf(n2);
  float d = n + (n == 0.0f ? 1e-6f : 0.0f);
  float r = 1.0f / d;
  *rinv_out = r;
  return mk4(q.w * r, q.x * r, q.y * r, q.z * r);
}

// sin/cos by 3-term Cody-Waite reduction (pi/2) and degree-7/8 minimax polynomials;
// max error about 1 ulp on [-pi, pi], identical bit-for-bit to the oracle's c_sincos.
__device__ __forceinline__ void sincos_canon(float x, float *sp, float *cp) {
  float j = rintf(x * 0.636619747f);
  float r = fmaf(-j, 1.57079637e+00f, x);
  r = fmaf(-j, -4.37113883e-08f, r);
  r = fmaf(-j, -1.71512489e-15f, r);
  float r2 = r * r;
  float ps = fmaf(r2, -1.9515295891e-4f, 8.3321608736e-3f);
  ps = fmaf(ps, r2, -1.6666654611e-1f);
  float s = fmaf(r * r2, ps, r);
  float pc = fmaf(r2, 2.443315711809948e-5f, -1.388731625493765e-3f);
  pc = fmaf(pc, r2, 4.166664568298827e-2f);
  float c = fmaf(r2 * r2, pc, fmaf(-0.5f, r2, 1.0f));
  int q = ((int)j) & 3;
  float so = (q & 1) ? c : s, co = (q & 1) ? s : c;
  if (q == 1 || q == 2) co = -co;
  if (q >= 2) so = -so;
  *sp = so;
  *cp = co;
}

// The same normalisation with a division-free reciprocal square root that is REPRODUCIBLE on a CPU (no MUFU approximation):
// the classic integer seed, two Newton steps r <- r (1.5 - h r^2) and a final one in residual form r <- r + (r / 2)(1 - n2 r^2), all
// plain fused multiply-adds (within 1.3 ulp of 1 / sqrt(n2)); a zero quaternion maps to zero as with math.normalize.
// *rinv_out = 1 / |q| (the reverse sweep uses it).
__device__ __forceinline__ Q4 normalize4_nr(Q4 q, float *rinv_out) {
  const float n2 = fmaf(q.z, q.z, fmaf(q.y, q.y, fmaf(q.x, q.x, q.w * q.w)));
  const float h = 0.5f * n2;
  float r = __int_as_float(0x5f3759df - (__float_as_int(n2) >> 1));
#pragma unroll
  for (int i = 0; i < 2; i++) r = r * fmaf(-(h * r), r, 1.5f);
  const float e = fmaf(-(n2 * r), r, 1.0f);
  r = fmaf(0.5f * r, e, r);
  *rinv_out = r;
  return mk4(q.w * r, q.x * r, q.y * r, q.z * r);
}

// sin/cos up to a COMMON sign: 3-term Cody-Waite reduction by pi, then degree-9 / degree-8 minimax polynomials on
// [-pi/2, pi/2] (sin within 2 ulp, cos within 1 ulp of 1).  Returns ((-1)^j sin x, (-1)^j cos x) with j = rint(x / pi):
// a hinge's half-angle pair enters FK only through products of two of its members (cos t, sin t) and through the
// quaternion (c, a s), whose overall sign does not change the rotation, so no quadrant logic is needed.
// The CPU oracle's f_sincos is the same sequence of operations.
__device__ __forceinline__ void sincos_pi(float x, float *sp, float *cp) {
  const float j = rintf(x * 0.318309873f);
  float r = fmaf(-j, 3.14159274e+00f, x);
  r = fmaf(-j, -8.74227766e-08f, r);
  r = fmaf(-j, -3.43024902e-15f, r);
  const float r2 = r * r;
  float ps = fmaf(r2, 2.6056311526190257e-06f, -0.0001980953966267407f);
  ps = fmaf(ps, r2, 0.008333065547049046f);
  ps = fmaf(ps, r2, -0.16666659712791443f);
  *sp = fmaf(r * r2, ps, r);
  float pc = fmaf(r2, -2.619248391511064e-07f, 2.4769240553723648e-05f);
  pc = fmaf(pc, r2, -0.0013888567918911576f);
  pc = fmaf(pc, r2, 0.041666656732559204f);
  *cp = fmaf(r2 * r2, pc, fmaf(-0.5f, r2, 1.0f));
}

// math.axis_angle_to_quat
__device__ __forceinline__ Q4 axis_angle(V3 a, float ang) {
  float s, c;
  sincos_canon(ang * 0.5f, &s, &c);
  return mk4(c, a.x * s, a.y * s, a.z * s);
}

// jnp.clip(x, lo, hi) = min(max(x, lo), hi), written with comparisons (signed-zero stable)
__device__ __forceinline__ float clipf(float x, float lo, float hi) {
  float t = (x > lo) ? x : lo;
  return (t < hi) ? t : hi;
}

// the same clip through the hardware min / max (one FMNMX each); differs from clipf at most in the sign of a zero
__device__ __forceinline__ float clipm(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) v = v + __shfl_xor_sync(0xffffffffu, v, off);
  return v;
}

}  // namespace stacb
