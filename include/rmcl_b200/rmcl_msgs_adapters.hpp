// rmcl_msgs_adapters.hpp -- the wire formats of the path (rmcl_msgs/msg/{ScanInfo,DepthInfo,O1DnInfo,OnDnInfo,RangeData,Scan,Depth,O1Dn,OnDn}.msg)
// fed straight into the B200 correspondence classes.  Duck-typed templates: they accept the rosidl-generated C++ structs (rmcl_msgs::msg::*)
// as well as any struct with the same field names (tests/test_cpp_boundary.py uses plain PODs), so this header does not depend on ROS 2.
//
// What each function replaces in the reference:
//   setModel(rcc, info)      rmcl::convert(info, sensor_model_) + setModel  (rmcl_ros/src/util/conversions.cpp:22-34, 48-60, 74-94, 96-120;
//                                                                          MICPSphericalSensorCPU.cpp:155-160)
//   setData(rcc, data)       the per-ray loop of MICP*Sensor*::unpackMessage (rmcl_ros/src/micpl/MICPSphericalSensorCPU.cpp:193-228): here one
//                            512 KiB upload + the unpack kernel on the device (dataset point = direction * range (+ origin), mask = range in [min, max];
//                            RangeData.mask is ignored exactly like the reference ignores it: "TODOs: use input mask values", :189-191)
//   unpackMessage(rcc, msg)  both, for a whole Scan / Depth / O1Dn / OnDn message (or its *Stamped payload)
//   fillSensorStats(...)     rmcl_msgs/MICPSensorStats as published at rmcl_ros/src/nodes/micp_localization.cpp:1009-1015
#pragma once
#include <vector>

#include "rcc_b200.hpp"

namespace rmcl {
namespace b200 {

template <typename ScanInfoT> inline void setModel(RCCB200Spherical& rcc, const ScanInfoT& info)
{
    rm::SphericalModel m;                                                   // conversions.cpp:22-34
    m.phi.min = info.phi_min; m.phi.inc = info.phi_inc; m.phi.size = info.phi_n;
    m.theta.min = info.theta_min; m.theta.inc = info.theta_inc; m.theta.size = info.theta_n;
    m.range.min = info.range_min; m.range.max = info.range_max;
    rcc.setModel(m);
}
template <typename DepthInfoT> inline void setModel(RCCB200Pinhole& rcc, const DepthInfoT& info)
{
    rm::PinholeModel m;                                                     // conversions.cpp:48-60
    m.width = info.width; m.height = info.height; m.f[0] = info.fx; m.f[1] = info.fy; m.c[0] = info.cx; m.c[1] = info.cy;
    m.range.min = info.range_min; m.range.max = info.range_max;
    rcc.setModel(m);
}
// O1Dn / OnDn: the Point32 arrays go to the C ABI as packed xyz floats (no intermediate rm::Memory)
template <typename O1DnInfoT> inline void setModel(RCCB200O1Dn& rcc, const O1DnInfoT& info)
{
    std::vector<float> dirs(3 * info.dirs.size());                          // conversions.cpp:74-94
    for (size_t i = 0; i < info.dirs.size(); i++) { dirs[3 * i] = info.dirs[i].x; dirs[3 * i + 1] = info.dirs[i].y; dirs[3 * i + 2] = info.dirs[i].z; }
    const float orig[3] = {info.orig.x, info.orig.y, info.orig.z};
    if (dirs.size() != 3 * size_t(info.width) * info.height) throw std::runtime_error("O1DnInfo: dirs.size() != width * height");
    b2_check(b2_rcc_set_model_o1dn(rcc.handle(), info.width, info.height, orig, dirs.data(), info.range_min, info.range_max), "setModel(O1DnInfo)");
    rcc.noteModelSize(info.width, info.height);
}
template <typename OnDnInfoT> inline void setModel(RCCB200OnDn& rcc, const OnDnInfoT& info)
{
    if (info.origs.size() != info.dirs.size() || info.dirs.size() != size_t(info.width) * info.height) throw std::runtime_error("OnDnInfo: origs / dirs size != width * height");
    std::vector<float> origs(3 * info.origs.size()), dirs(3 * info.dirs.size());                    // conversions.cpp:96-120
    for (size_t i = 0; i < info.dirs.size(); i++) {
        origs[3 * i] = info.origs[i].x; origs[3 * i + 1] = info.origs[i].y; origs[3 * i + 2] = info.origs[i].z;
        dirs[3 * i] = info.dirs[i].x; dirs[3 * i + 1] = info.dirs[i].y; dirs[3 * i + 2] = info.dirs[i].z;
    }
    b2_check(b2_rcc_set_model_ondn(rcc.handle(), info.width, info.height, origs.data(), dirs.data(), info.range_min, info.range_max), "setModel(OnDnInfo)");
    rcc.noteModelSize(info.width, info.height);
}

// RangeData -> dataset (device-side unpackMessage)
template <typename RangeDataT> inline void setData(CorrespondencesB200& rcc, const RangeDataT& data) { rcc.setRanges(data.ranges.data(), data.ranges.size()); }

// Scan / Depth / O1Dn / OnDn: { info, data }
template <typename RCC, typename MsgT> inline void unpackMessage(RCC& rcc, const MsgT& msg) { setModel(rcc, msg.info); setData(rcc, msg.data); }

// rmcl_msgs/MICPSensorStats (micp_localization.cpp:1009-1015): valid_matches = Cmerged_o.n_meas, cov_trace = trace(Cmerged_o.covariance)
template <typename StatsMsgT> inline void fillSensorStats(StatsMsgT& stats, const rm::CrossStatistics& Cmerged_o, uint32_t total_measurements, uint32_t valid_measurements)
{
    stats.total_measurements = total_measurements; stats.valid_measurements = valid_measurements;
    stats.valid_matches = Cmerged_o.n_meas; stats.cov_trace = Cmerged_o.covariance.trace();
}

}  // namespace b200
}  // namespace rmcl
