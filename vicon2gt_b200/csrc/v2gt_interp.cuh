// vicon2gt_b200 kernel (2): vicon bracket search + SO(3) x R^3 interpolation (device functions).
//
// Replaces Interpolator::get_pose / get_pose_with_jacobian (src/meas/Interpolator.cpp:60-288).
// The pose store is a time-sorted, de-duplicated array (what the reference's std::set holds);
// the bracket is found with two bisections and reproduces std::set::equal_range exactly:
//   lo = first t >= tq, hi = first t > tq; fail if lo==begin, lo==end or hi==end;
//   idx0 = lo-1, idx1 = hi   (so an exact hit on sample k brackets with k-1 and k+1).
#pragma once
#include "v2gt_common.cuh"

namespace v2 {

struct InterpResult {
  int idx0, idx1;
  int ok;      // return value of the reference function
  int written; // outputs assigned (0 only for the gap-threshold failure)
  double lambda;
  double q[4], p[3];
  double Rtl[9], Rtr[9], Rbr[9]; // covariance blocks: R = [Rtl Rtr; Rtr^T Rbr]
  double Ht[6];                  // d pose / d t_off (get_pose_with_jacobian only)
};

// Batched-binary-search primitive: first index in [0,n) with t[i] >= tq (lower) / t[i] > tq (upper).
__device__ __forceinline__ long long lower_bound_d(const double *__restrict__ t, long long n, double tq) {
  long long lo = 0, hi = n;
  while (lo < hi) {
    const long long mid = (lo + hi) >> 1;
    if (__ldg(t + mid) < tq)
      lo = mid + 1;
    else
      hi = mid;
  }
  return lo;
}
__device__ __forceinline__ long long upper_bound_d(const double *__restrict__ t, long long n, double tq) {
  long long lo = 0, hi = n;
  while (lo < hi) {
    const long long mid = (lo + hi) >> 1;
    if (__ldg(t + mid) <= tq)
      lo = mid + 1;
    else
      hi = mid;
  }
  return lo;
}

// lower_bound with a guess `h` of the answer (the bracket of the same state at its previous evaluation: the query time
// t - t_off moves by less than a vicon period between evaluations once the optimiser has settled).  Checks h, then
// h -+ 1, with two / three loads instead of the ~15 dependent ones of the bisection; any other case falls back to it.
// The result is the bisection's, always.
__device__ __forceinline__ long long lower_bound_hint(const double *__restrict__ t, long long n, double tq, long long h) {
  if (h >= 1 && h < n) {
    const double a = __ldg(t + h - 1), b = __ldg(t + h);
    if (a < tq) {
      if (!(b < tq))
        return h; // t[h-1] < tq <= t[h]
      if (h + 1 == n)
        return n;
      if (!(__ldg(t + h + 1) < tq))
        return h + 1;
    } else {
      if (h == 1)
        return 0; // !(t[0] < tq)
      if (__ldg(t + h - 2) < tq)
        return h - 1;
    }
  }
  return lower_bound_d(t, n, tq);
}

// Interpolator.cpp:175-207 / :63-95.  Returns 0 ok, 1 "lo == begin", 2 "lo == end or hi == end".
__device__ __forceinline__ int vicon_bracket(const double *__restrict__ vt, long long n, double tq, int &idx0, int &idx1,
                                             long long hint = -1) {
  idx0 = -1;
  idx1 = -1;
  const long long lo = lower_bound_hint(vt, n, tq, hint);
  if (lo == 0)
    return 1;
  if (lo == n)
    return 2;
  // hi = upper bound: lo if t[lo] > tq, else lo+1 (keys are unique)
  const long long hi = (__ldg(vt + lo) > tq) ? lo : lo + 1;
  if (hi == n)
    return 2;
  idx0 = (int)(lo - 1);
  idx1 = (int)hi;
  return 0;
}

// Interpolator.cpp:239-287 (and :127-169).  vs: pose rows [V][8] (q4 p3 pad); cov: [V][18] or the
// trial-constant 18 doubles.
__device__ inline void vicon_interp(const double *__restrict__ vt, const double *__restrict__ vs, const double *__restrict__ vc,
                                    bool cov_const, long long n, double tq, double thresh, bool with_jac, InterpResult &o,
                                    long long hint = -1) {
  o.ok = 0;
  o.written = 1;
  o.lambda = 0;
  o.q[0] = 0; o.q[1] = 0; o.q[2] = 0; o.q[3] = 1;
  o.p[0] = 0; o.p[1] = 0; o.p[2] = 0;
#pragma unroll
  for (int i = 0; i < 9; i++) {
    o.Rtl[i] = 0; o.Rtr[i] = 0; o.Rbr[i] = 0;
  }
#pragma unroll
  for (int i = 0; i < 6; i++)
    o.Ht[i] = 0;
  const int br = vicon_bracket(vt, n, tq, o.idx0, o.idx1, hint);
  if (br != 0)
    return;
  const double t0 = __ldg(vt + o.idx0), t1 = __ldg(vt + o.idx1);
  if (fabs(tq - t0) > thresh || fabs(tq - t1) > thresh) {
    o.written = 0;
    return;
  }
  const double lambda = (tq - t0) / (t1 - t0);
  o.lambda = lambda;
  double q0[4], q1[4], p0[3], p1[3];
  {
    const double4 a0 = ldg4(vs + 8 * (long long)o.idx0);
    const double4 b0 = ldg4(vs + 8 * (long long)o.idx0 + 4);
    const double4 a1 = ldg4(vs + 8 * (long long)o.idx1);
    const double4 b1 = ldg4(vs + 8 * (long long)o.idx1 + 4);
    q0[0] = a0.x; q0[1] = a0.y; q0[2] = a0.z; q0[3] = a0.w;
    p0[0] = b0.x; p0[1] = b0.y; p0[2] = b0.z;
    q1[0] = a1.x; q1[1] = a1.y; q1[2] = a1.z; q1[3] = a1.w;
    p1[0] = b1.x; p1[1] = b1.y; p1[2] = b1.z;
  }
  double R0[9], R1[9], R01[9], r01[3], lr[3], R0i[9], Ri[9];
  quat_2_Rot(q0, R0);
  quat_2_Rot(q1, R1);
  mm3_nt(R1, R0, R01);
  log_so3(R01, r01);
  lr[0] = lambda * r01[0]; lr[1] = lambda * r01[1]; lr[2] = lambda * r01[2];
  exp_so3(lr, R0i);
  mm3(R0i, R0, Ri);
  rot_2_quat(Ri, o.q);
#pragma unroll
  for (int d = 0; d < 3; d++)
    o.p[d] = (1 - lambda) * p0[d] + lambda * p1[d];

  // covariance propagation (Interpolator.cpp:252-276).  JRinv_r01 is inverted twice there, i.e. it is
  // Jr(r01); JRneg_r0i = Jr(-lambda log(R01^T)) and log(R01^T) == -log(R01) exactly, so it equals JR_r0i.
  double JR[9], JRi[9], Mx[9], T1[9];
  Jr_so3(lr, JR);
  Jr_so3(r01, JRi);
#pragma unroll
  for (int i = 0; i < 9; i++)
    T1[i] = JR[i] * lambda;
  mm3(T1, JRi, Mx);
  double Ha[9], Hb[9], T2[9];
#pragma unroll
  for (int i = 0; i < 9; i++)
    T2[i] = Mx[i] - ((i % 4 == 0) ? 1.0 : 0.0);
  mm3(R0i, T2, Ha);
#pragma unroll
  for (int i = 0; i < 9; i++)
    Ha[i] = -Ha[i];
  mm3(R0i, Mx, Hb);
  double Rq0[9], Rq1[9], Rp1[9];
  {
    const double *c0 = cov_const ? vc : vc + 18 * (long long)o.idx0;
    const double *c1 = cov_const ? vc : vc + 18 * (long long)o.idx1;
#pragma unroll
    for (int i = 0; i < 9; i++) {
      Rq0[i] = __ldg(c0 + i);
      Rq1[i] = __ldg(c1 + i);
      Rp1[i] = __ldg(c1 + 9 + i);
    }
  }
  // Hu = [Ha 0 Hb 0 ; 0 0 (1-l)I lI] (the (1-l)I block sits on pose-1 ORIENTATION noise, Interpolator.cpp:266)
  double A1[9], A2[9], B1[9];
  mm3(Ha, Rq0, T1);
  mm3_nt(T1, Ha, A1);
  mm3(Hb, Rq1, B1); // Hb Rq1
  mm3_nt(B1, Hb, A2);
  const double oml = 1 - lambda;
#pragma unroll
  for (int i = 0; i < 9; i++) {
    o.Rtl[i] = A1[i] + A2[i];
    o.Rtr[i] = B1[i] * oml;
    o.Rbr[i] = (oml * Rq1[i]) * oml + (lambda * Rp1[i]) * lambda;
  }
  if (with_jac) {
    const double hl = -1.0 / (t1 - t0);
    double T3[9], v[3];
    mm3(R0i, JR, T3);
    mv3(T3, r01, v);
#pragma unroll
    for (int d = 0; d < 3; d++) {
      o.Ht[d] = (-v[d]) * hl;
      o.Ht[3 + d] = (p1[d] - p0[d]) * hl;
    }
  }
  o.ok = 1;
}

// Expand the block form into the dense 6x6 (row-major) covariance.
__device__ __forceinline__ void interp_dense_cov(const InterpResult &r, double *R36) {
#pragma unroll
  for (int a = 0; a < 3; a++)
#pragma unroll
    for (int b = 0; b < 3; b++) {
      R36[6 * a + b] = r.Rtl[3 * a + b];
      R36[6 * a + 3 + b] = r.Rtr[3 * a + b];
      R36[6 * (3 + a) + b] = r.Rtr[3 * b + a];
      R36[6 * (3 + a) + 3 + b] = r.Rbr[3 * a + b];
    }
}

} // namespace v2
