// ORACLE (test infrastructure, NOT product code) -- pinned on the reference's own sources (oracle/_ref, tests/test_ref_pin_cpu.py).
//
// Measurement stores: IMU store + interval selection (Propagator) and the vicon
// pose store with SO(3)xR^3 interpolation (Interpolator), restated from the
// reference's src/meas/Propagator.{h,cpp} and src/meas/Interpolator.{h,cpp}.
#pragma once
#include "cpi.hpp"
#include <algorithm>
#include <limits>
#include <vector>

namespace orc {

// src/meas/Propagator.h:31-35
struct ImuData {
  double timestamp;
  Vec3 wm, am;
};

class Propagator {
public:
  // src/meas/Propagator.h:41-46
  Propagator(double sigmaw, double sigmawb, double sigmaa, double sigmaab)
      : sigma_w(sigmaw), sigma_wb(sigmawb), sigma_a(sigmaa), sigma_ab(sigmaab) {}

  // src/meas/Propagator.cpp:21-31 (appended, assumed time sorted, never sorted)
  void feed_imu(double timestamp, const Vec3 &wm, const Vec3 &am) { imu_data.push_back(ImuData{timestamp, wm, am}); }

  // src/meas/Propagator.h:63-73
  static ImuData interpolate_data(const ImuData &imu_1, const ImuData &imu_2, double timestamp) {
    const double lambda = (timestamp - imu_1.timestamp) / (imu_2.timestamp - imu_1.timestamp);
    ImuData data;
    data.timestamp = timestamp;
    data.am = (1 - lambda) * imu_1.am + lambda * imu_2.am;
    data.wm = (1 - lambda) * imu_1.wm + lambda * imu_2.wm;
    return data;
  }

  // Sample selection of src/meas/Propagator.cpp:33-132.  `faithful` keeps the
  // reference's scan from i = 0 (O(M) per interval); otherwise the scan starts at
  // the last sample at or before time0 (same visited cases, found by bisection).
  bool select(double time0, double time1, std::vector<ImuData> &prop_data, bool faithful) const {
    prop_data.clear();
    if (imu_data.empty())
      return false;
    size_t i_begin = 0;
    if (!faithful) {
      // lb = first sample with t >= time0.  For i < lb-1 both t[i] and t[i+1] are < time0 and
      // none of the three cases below can fire, so starting at lb-1 visits the same cases.
      auto it = std::lower_bound(imu_data.begin(), imu_data.end(), time0,
                                 [](const ImuData &d, double t) { return d.timestamp < t; });
      size_t lb = (size_t)(it - imu_data.begin());
      i_begin = (lb >= 1) ? lb - 1 : 0;
    }
    for (size_t i = i_begin; i + 1 < imu_data.size(); i++) {
      const ImuData &d0 = imu_data[i];
      const ImuData &d1 = imu_data[i + 1];
      // CASE 1: start of the integration period -> split the current measurement
      if (d1.timestamp > time0 && d0.timestamp < time0) {
        prop_data.push_back(interpolate_data(d0, d1, time0));
        continue;
      }
      // CASE 2: fully inside
      if (d0.timestamp >= time0 && d1.timestamp <= time1) {
        prop_data.push_back(d0);
        continue;
      }
      // CASE 3: end of the period
      if (d1.timestamp > time1) {
        if (d0.timestamp > time1 && i == 0) {
          break;
        } else if (d0.timestamp > time1) {
          prop_data.push_back(interpolate_data(imu_data[i - 1], d0, time1));
        } else {
          prop_data.push_back(d0);
        }
        if (prop_data.back().timestamp != time1) {
          prop_data.push_back(interpolate_data(d0, d1, time1));
        }
        break;
      }
    }
    if (prop_data.empty())
      return false;
    // drop zero-dt entries (Propagator.cpp:118-126)
    for (size_t i = 0; i + 1 < prop_data.size(); i++) {
      if (std::abs(prop_data[i + 1].timestamp - prop_data[i].timestamp) < 1e-12) {
        prop_data.erase(prop_data.begin() + i);
        i--;
      }
    }
    return prop_data.size() >= 2;
  }

  // src/meas/Propagator.cpp:33-163
  bool propagate(double time0, double time1, const Vec3 &bg_lin, const Vec3 &ba_lin, CpiV1 &integration,
                 bool faithful = false) const {
    std::vector<ImuData> prop_data;
    if (!select(time0, time1, prop_data, faithful))
      return false;
    integration = CpiV1(sigma_w, sigma_wb, sigma_a, sigma_ab, true);
    integration.setLinearizationPoints(bg_lin, ba_lin);
    for (size_t i = 0; i + 1 < prop_data.size(); i++) {
      integration.feed_IMU(prop_data[i].timestamp, prop_data[i + 1].timestamp, prop_data[i].wm, prop_data[i].am,
                           prop_data[i + 1].wm, prop_data[i + 1].am);
    }
    return true;
  }

  // src/meas/Propagator.cpp:165-193.  For a time-sorted store this is
  // t[0] <= timestamp <= t[M-1]; `faithful` keeps the linear scan.
  bool has_bounding_imu(double timestamp, bool faithful = false) const {
    if (imu_data.empty())
      return false;
    if (!faithful) {
      bool has_lower = false, has_upper = false;
      // Equivalent closed form of the scan below for ANY ordering of the store up to the first
      // sample >= timestamp is not available, so only use the shortcut when the store is sorted.
      if (sorted_checked_ == 0) {
        sorted_checked_ = 1;
        for (size_t i = 0; i + 1 < imu_data.size(); i++)
          if (imu_data[i + 1].timestamp < imu_data[i].timestamp) {
            sorted_checked_ = 2;
            break;
          }
      }
      if (sorted_checked_ == 1) {
        has_lower = imu_data.front().timestamp <= timestamp;
        has_upper = imu_data.back().timestamp >= timestamp;
        return has_lower && has_upper;
      }
    }
    bool has_lower = false, has_upper = false;
    for (size_t i = 0; i < imu_data.size(); i++) {
      if (imu_data[i].timestamp <= timestamp)
        has_lower = true;
      if (imu_data[i].timestamp >= timestamp)
        has_upper = true;
      if (has_lower && has_upper)
        break;
    }
    return has_lower && has_upper;
  }

  std::vector<ImuData> imu_data;
  double sigma_w, sigma_wb, sigma_a, sigma_ab;

private:
  mutable int sorted_checked_ = 0;
};

// src/meas/Interpolator.h:30-46 (pose part only; feed_odom has no call sites)
struct PoseData {
  double timestamp = -1;
  Vec4 q;
  Vec3 p;
  Mat3 R_q, R_p;
};

class Interpolator {
public:
  // src/meas/Interpolator.cpp:21-38.  The reference keeps a std::set keyed by
  // timestamp: sorted, and a duplicate key is silently dropped (first one wins).
  void feed_pose(double timestamp, const Vec4 &q, const Vec3 &p, const Mat3 &R_q, const Mat3 &R_p) {
    PoseData d;
    d.timestamp = timestamp;
    d.q = q;
    d.p = p;
    d.R_q = R_q;
    d.R_p = R_p;
    auto it = std::lower_bound(pose_data.begin(), pose_data.end(), timestamp,
                               [](const PoseData &a, double t) { return a.timestamp < t; });
    if (it == pose_data.end() || it->timestamp != timestamp)
      pose_data.insert(it, d);
    time_min = std::min(time_min, timestamp);
    time_max = std::max(time_max, timestamp);
  }

  // Bracket search shared by get_pose / get_pose_with_jacobian
  // (src/meas/Interpolator.cpp:63-95 and :175-207): equal_range, fail when
  // lower==begin, lower==end or upper==end; idx0 = lower-1, idx1 = upper.
  // Returns 0 ok, 1 = "before first" failure, 2 = "no upper/at end" failure.
  int bracket(double timestamp, long &idx0, long &idx1) const {
    idx0 = -1;
    idx1 = -1;
    auto lo = std::lower_bound(pose_data.begin(), pose_data.end(), timestamp,
                               [](const PoseData &a, double t) { return a.timestamp < t; });
    auto hi = std::upper_bound(pose_data.begin(), pose_data.end(), timestamp,
                               [](double t, const PoseData &a) { return t < a.timestamp; });
    if (lo == pose_data.begin())
      return 1;
    if (lo == pose_data.end() || hi == pose_data.end())
      return 2;
    idx0 = (long)(lo - pose_data.begin()) - 1;
    idx1 = (long)(hi - pose_data.begin());
    return 0;
  }

  struct Result {
    bool ok = false;       // the function's return value
    bool written = false;  // outputs were assigned (false only for the gap-threshold failure)
    long idx0 = -1, idx1 = -1;
    double lambda = 0;
    Vec4 q;
    Vec3 p;
    Mat6 R;
    Vec6 H_toff;
  };

  // src/meas/Interpolator.cpp:60-170 (thresh 5.0 s) and :172-288 (thresh 0.1 s, + H_toff)
  Result interpolate(double timestamp, double thresh_sec, bool with_jacobian) const {
    Result out;
    long i0, i1;
    const int br = bracket(timestamp, i0, i1);
    if (br != 0) {
      out.q = Vec4{0, 0, 0, 1};
      out.written = true; // q=[0 0 0 1], p=0, R=0 are assigned
      return out;
    }
    out.idx0 = i0;
    out.idx1 = i1;
    const PoseData &pose0 = pose_data[(size_t)i0];
    const PoseData &pose1 = pose_data[(size_t)i1];
    // (the "exact" branch of Interpolator.cpp:98/:210 cannot fire for unique keys)
    if (std::abs(timestamp - pose0.timestamp) > thresh_sec || std::abs(timestamp - pose1.timestamp) > thresh_sec) {
      return out; // nothing written
    }
    const double lambda = (timestamp - pose0.timestamp) / (pose1.timestamp - pose0.timestamp);
    out.lambda = lambda;
    const Mat3 R_Gto0 = quat_2_Rot(pose0.q);
    const Mat3 R_Gto1 = quat_2_Rot(pose1.q);
    const Mat3 R_0to1 = R_Gto1 * R_Gto0.transpose();
    const Vec3 r01 = log_so3(R_0to1);
    const Mat3 R_0toi = exp_so3(lambda * r01);
    const Mat3 R_interp = R_0toi * R_Gto0;
    const Vec3 p_interp = (1 - lambda) * pose0.p + lambda * pose1.p;

    const Mat3 eye33 = Mat3::Identity();
    const Mat3 JR_r0i = Jr_so3(lambda * r01);
    // Interpolator.cpp:256-257 inverts twice, i.e. JRinv_r01 == Jr_so3(log(R_0to1))
    const Mat3 JRinv_r01 = Jr_so3(r01);
    const Mat3 JRneg_r0i = Jr_so3(-lambda * log_so3(R_0to1.transpose()));

    Mat<6, 12> Hu;
    Hu.setBlock(0, 0, -(R_0toi * (JR_r0i * lambda * JRinv_r01 - eye33)));
    Hu.setBlock(0, 6, R_0toi * (JRneg_r0i * lambda * JRinv_r01));
    Hu.setBlock(3, 6, (1 - lambda) * eye33); // column 6 (not 3) as in Interpolator.cpp:266
    Hu.setBlock(3, 9, lambda * eye33);
    Mat<12, 12> R_12;
    R_12.setBlock(0, 0, pose0.R_q);
    R_12.setBlock(3, 3, pose0.R_p);
    R_12.setBlock(6, 6, pose1.R_q);
    R_12.setBlock(9, 9, pose1.R_p);
    out.R = Hu * R_12 * Hu.transpose();

    if (with_jacobian) {
      const double H_lambda2toff = -1.0 / (pose1.timestamp - pose0.timestamp);
      const Vec3 top = -((R_0toi * JR_r0i) * r01) * H_lambda2toff;
      const Vec3 bot = (pose1.p - pose0.p) * H_lambda2toff;
      for (int i = 0; i < 3; i++) {
        out.H_toff(i) = top(i);
        out.H_toff(3 + i) = bot(i);
      }
    }
    out.q = rot_2_quat(R_interp);
    out.p = p_interp;
    out.ok = true;
    out.written = true;
    return out;
  }

  Result get_pose(double t) const { return interpolate(t, 5.0, false); }
  Result get_pose_with_jacobian(double t) const { return interpolate(t, 0.1, true); }

  double get_time_min() const { return time_min; }
  double get_time_max() const { return time_max; }
  size_t size() const { return pose_data.size(); }

  std::vector<PoseData> pose_data; // sorted, unique timestamps (std::set semantics)
  double time_min = std::numeric_limits<double>::infinity();
  double time_max = -std::numeric_limits<double>::infinity();
};

} // namespace orc
