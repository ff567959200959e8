// gen_ref_golden.cpp -- drives the REFERENCE's own classes (rmcl::RCCEmbree*, rmagine simulators / statistics / umeyama, Embree through
// rmagine's EmbreeMap) on the scenarios written by make_ref_inputs.py and dumps what they return.  Test infrastructure: never shipped, never
// linked by the product; it only links the reference.  Cannot be built in the authoring image (no rmagine / Embree): see CMakeLists.txt.
//
// Every block names the reference code it calls (paths relative to the rmcl checkout).
#include <rmcl/registration/RCCEmbree.hpp>                 // rmcl/include/rmcl/registration/RCCEmbree.hpp:18-83
#include <rmcl/registration/CPCEmbree.hpp>                 // rmcl/include/rmcl/registration/CPCEmbree.hpp:20-54

#include <rmagine/map/EmbreeMap.hpp>
#include <rmagine/map/embree/embree_shapes.h>
#include <rmagine/math/statistics.h>                       // rm::statistics_p2l (called at rmcl/src/rmcl/registration/CorrespondencesCPU.cpp:26-30)
#include <rmagine/math/linalg.h>                           // rm::umeyama_transform (called at rmcl_ros/src/nodes/micp_localization.cpp:952-953)
#include <rmagine/types/sensor_models.h>

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <stdexcept>
#include <string>
#include <vector>

namespace rm = rmagine;

// ---- tiny binary container: magic "B2REF1\0\0", then records { char name[32]; uint32 dtype (0 f32, 1 u8, 2 u32, 3 f64); uint32 count; payload } ----
struct Writer {
    std::ofstream f;
    explicit Writer(const std::string& path) : f(path, std::ios::binary) { if (!f) throw std::runtime_error("cannot write " + path); f.write("B2REF1\0\0", 8); }
    void put(const char* name, uint32_t dtype, const void* p, uint32_t count, size_t elem)
    {
        char nm[32] = {0}; strncpy(nm, name, 31);
        f.write(nm, 32); f.write((const char*)&dtype, 4); f.write((const char*)&count, 4); f.write((const char*)p, (std::streamsize)(elem * count));
    }
    void f32(const char* n, const float* p, size_t c) { put(n, 0, p, (uint32_t)c, 4); }
    void u8(const char* n, const uint8_t* p, size_t c) { put(n, 1, p, (uint32_t)c, 1); }
    void u32(const char* n, const uint32_t* p, size_t c) { put(n, 2, p, (uint32_t)c, 4); }
};
struct Reader {
    std::vector<char> buf;
    explicit Reader(const std::string& path)
    {
        std::ifstream f(path, std::ios::binary); if (!f) throw std::runtime_error("cannot read " + path);
        buf.assign(std::istreambuf_iterator<char>(f), {});
        if (buf.size() < 8 || memcmp(buf.data(), "B2REF1", 6)) throw std::runtime_error("bad magic in " + path);
    }
    const char* find(const char* name, uint32_t& count) const
    {
        size_t o = 8;
        while (o + 40 <= buf.size()) {
            uint32_t dtype, cnt; memcpy(&dtype, &buf[o + 32], 4); memcpy(&cnt, &buf[o + 36], 4);
            const size_t elem = dtype == 1 ? 1 : (dtype == 3 ? 8 : 4);
            if (!strncmp(&buf[o], name, 32)) { count = cnt; return &buf[o + 40]; }
            o += 40 + elem * cnt;
        }
        throw std::runtime_error(std::string("record not found: ") + name);
    }
    std::vector<float> f32(const char* n) const { uint32_t c; const char* p = find(n, c); std::vector<float> v(c); memcpy(v.data(), p, 4 * (size_t)c); return v; }
};

static rm::Transform tf_from(const std::vector<float>& v, size_t i = 0)       // [qx qy qz qw tx ty tz] like rmcl_b200.synth.TRANSFORM_DTYPE
{
    rm::Transform T;
    T.R.x = v[8 * i + 0]; T.R.y = v[8 * i + 1]; T.R.z = v[8 * i + 2]; T.R.w = v[8 * i + 3];
    T.t.x = v[8 * i + 4]; T.t.y = v[8 * i + 5]; T.t.z = v[8 * i + 6];
    return T;
}
static void put_tf(Writer& w, const char* n, const rm::Transform& T)
{
    const float v[8] = {T.R.x, T.R.y, T.R.z, T.R.w, T.t.x, T.t.y, T.t.z, 0.f};
    w.f32(n, v, 8);
}
static void put_stats(Writer& w, const char* n, const rm::CrossStatistics& s)
{
    float v[16];
    v[0] = s.dataset_mean.x; v[1] = s.dataset_mean.y; v[2] = s.dataset_mean.z; v[3] = s.model_mean.x; v[4] = s.model_mean.y; v[5] = s.model_mean.z;
    for (int c = 0; c < 3; c++) for (int r = 0; r < 3; r++) v[6 + c * 3 + r] = s.covariance(r, c);       // column-major like rm::Matrix3x3 storage
    uint32_t n_meas = s.n_meas; memcpy(&v[15], &n_meas, 4);
    w.f32(n, v, 16);
}

// One MICP-L scenario through the reference's classes: find, computeCrossStatistics, umeyama, and the correctOnce loop
// (rmcl_ros/src/nodes/micp_localization.cpp:899-984 with MICPSensor_::findCorrespondences / computeCrossStatistics,
//  rmcl_ros/include/rmcl_ros/micpl/MICPSensor.hpp:146-184, restated around the real objects because the node itself needs ROS 2).
template <typename RCC, typename Model>
static void run_micp(const std::string& tag, rm::EmbreeMapPtr map, const Model& model, const Reader& in, Writer& out)
{
    const rm::Transform Tsb = tf_from(in.f32((tag + ".Tsb").c_str()));
    const rm::Transform Tbo = tf_from(in.f32((tag + ".Tbo").c_str()));
    const rm::Transform Tom = tf_from(in.f32((tag + ".Tom").c_str()));
    const rm::Transform Tgt = tf_from(in.f32((tag + ".Tgt").c_str()));
    const std::vector<float> ranges = in.f32((tag + ".ranges").c_str());
    const std::vector<float> prm = in.f32((tag + ".params").c_str());          // max_dist, adaptive_max_dist_min, convergence_progress, iterations

    RCC rcc(map);                                                              // RCCEmbree.hpp:25-26
    rcc.setTsb(Tsb);                                                           // RCCEmbree.cpp:15-19
    rcc.setModel(model);                                                       // RCCEmbree.cpp:21-24
    rcc.params.max_dist = prm[0];                                              // Correspondences.hpp:22
    rcc.adaptive_max_dist_min = prm[1];                                        // Correspondences.hpp:23

    // ---- simulate at the ground-truth pose: what rm::*SimulatorEmbree::simulate returns (RCCEmbree.cpp:35) ----
    rcc.find(Tgt);
    {
        auto mv = rcc.modelView();                                             // Correspondences.hpp:47-54
        const size_t n = mv.points.size();
        out.f32((tag + ".gt.points").c_str(), reinterpret_cast<const float*>(mv.points.raw()), 3 * n);
        out.f32((tag + ".gt.normals").c_str(), reinterpret_cast<const float*>(mv.normals.raw()), 3 * n);
        out.u8((tag + ".gt.hits").c_str(), reinterpret_cast<const uint8_t*>(mv.mask.raw()), n);
    }

    // ---- dataset from the scan exactly like MICPSphericalSensorCPU::unpackMessage (rmcl_ros/src/micpl/MICPSphericalSensorCPU.cpp:193-228) ----
    const size_t n = ranges.size();
    rcc.dataset.points.resize(n);
    rcc.dataset.mask.resize(n);
    for (unsigned int vid = 0; vid < model.getHeight(); vid++)
        for (unsigned int hid = 0; hid < model.getWidth(); hid++) {
            const unsigned int loc_id = model.getBufferId(vid, hid);
            const float real_range = ranges[loc_id];
            rcc.dataset.points[loc_id] = model.getDirection(vid, hid) * real_range;
            rcc.dataset.mask[loc_id] = (real_range < model.range.min || real_range > model.range.max) ? 0 : 1;
        }
    out.f32((tag + ".dataset.points").c_str(), reinterpret_cast<const float*>(rcc.dataset.points.raw()), 3 * n);
    out.u8((tag + ".dataset.mask").c_str(), reinterpret_cast<const uint8_t*>(rcc.dataset.mask.raw()), n);

    // ---- find at the pose guess, one reduction, one Umeyama ----
    const rm::Transform Tbm = Tom * Tbo;                                       // MICPSensor.hpp:148
    rcc.find(Tbm);
    {
        auto mv = rcc.modelView();
        out.f32((tag + ".guess.points").c_str(), reinterpret_cast<const float*>(mv.points.raw()), 3 * n);
        out.f32((tag + ".guess.normals").c_str(), reinterpret_cast<const float*>(mv.normals.raw()), 3 * n);
        out.u8((tag + ".guess.hits").c_str(), reinterpret_cast<const uint8_t*>(mv.mask.raw()), n);
    }
    const double cp = prm[2];
    const rm::CrossStatistics s0 = rcc.computeCrossStatistics(rm::Transform::Identity(), cp);       // CorrespondencesCPU.cpp:10-39
    put_stats(out, (tag + ".stats0").c_str(), s0);
    put_tf(out, (tag + ".umeyama0").c_str(), rm::umeyama_transform(s0));                            // micp_localization.cpp:952-953

    // ---- correctOnce (single sensor, merge weight 1): micp_localization.cpp:910-984 ----
    rm::Transform T_onew_oold = rm::Transform::Identity();
    rm::CrossStatistics Cmerged_o = rm::CrossStatistics::Identity();
    const unsigned iterations = (unsigned)prm[3];
    for (unsigned i = 0; i < iterations; i++) {
        Cmerged_o = rm::CrossStatistics::Identity();
        rm::CrossStatistics Cmerged_weighted_o = rm::CrossStatistics::Identity();
        const rm::Transform T_bnew_bold = ~Tbo * T_onew_oold * Tbo;                                 // :926
        const rm::Transform T_snew_sold = ~Tsb * T_bnew_bold * Tsb;                                 // MICPSensor.hpp:178
        const rm::CrossStatistics stats_s = rcc.computeCrossStatistics(T_snew_sold, cp);            // MICPSensor.hpp:179-180
        const rm::CrossStatistics Cs_b = Tsb * stats_s;                                             // MICPSensor.hpp:182
        const rm::CrossStatistics Cs_o = Tbo * Cs_b;                                                // :931
        rm::CrossStatistics Cs_weighted_o = Cs_o;
        Cs_weighted_o.n_meas *= 1.0;                                                                // :933-934 (merge_weight_multiplier default 1.0)
        Cmerged_o += Cs_o;                                                                          // :936
        Cmerged_weighted_o += Cs_weighted_o;                                                        // :937
        const rm::Transform T_inner = rm::umeyama_transform(Cmerged_weighted_o);                    // :952-953
        T_onew_oold = T_onew_oold * T_inner;                                                        // :963
    }
    rm::Transform Tom_new = Tom * T_onew_oold;                                                      // :972
    if (Cmerged_o.n_meas > 0) Tom_new.R.normalizeInplace(); else Tom_new = Tom;                     // :974-984
    put_tf(out, (tag + ".Tom_new").c_str(), Tom_new);
    put_tf(out, (tag + ".T_onew_oold").c_str(), T_onew_oold);
    put_stats(out, (tag + ".Cmerged_o").c_str(), Cmerged_o);
}

// The PF's direct Embree call (rmcl_ros/src/rmcl/PCDSensorUpdaterEmbree.cpp:30-47): rtcIntersect1(tnear 0, tfar inf) on the map's scene;
// geomID / tfar / raw Ng per ray.  (The updater class itself needs ROS 2; this is its ray query verbatim.)
static void run_rays(const std::string& tag, rm::EmbreeMapPtr map, const Reader& in, Writer& out)
{
    const std::vector<float> o = in.f32((tag + ".origs").c_str()), d = in.f32((tag + ".dirs").c_str());
    const size_t n = o.size() / 3;
    std::vector<float> t(n), ng(3 * n); std::vector<uint8_t> hit(n); std::vector<uint32_t> prim(n);
    for (size_t i = 0; i < n; i++) {
        RTCRayHit rayhit;
        rayhit.ray.org_x = o[3 * i]; rayhit.ray.org_y = o[3 * i + 1]; rayhit.ray.org_z = o[3 * i + 2];
        rayhit.ray.dir_x = d[3 * i]; rayhit.ray.dir_y = d[3 * i + 1]; rayhit.ray.dir_z = d[3 * i + 2];
        rayhit.ray.tnear = 0;
        rayhit.ray.tfar = std::numeric_limits<float>::infinity();
        rayhit.ray.mask = -1; rayhit.ray.flags = 0;
        rayhit.hit.geomID = RTC_INVALID_GEOMETRY_ID; rayhit.hit.instID[0] = RTC_INVALID_GEOMETRY_ID;
        rtcIntersect1(map->scene->handle(), &rayhit);                                               // :44
        hit[i] = rayhit.hit.geomID != RTC_INVALID_GEOMETRY_ID;
        t[i] = rayhit.ray.tfar; prim[i] = rayhit.hit.primID;
        ng[3 * i] = rayhit.hit.Ng_x; ng[3 * i + 1] = rayhit.hit.Ng_y; ng[3 * i + 2] = rayhit.hit.Ng_z;
    }
    out.f32((tag + ".t").c_str(), t.data(), n); out.u8((tag + ".hit").c_str(), hit.data(), n);
    out.u32((tag + ".prim").c_str(), prim.data(), n); out.f32((tag + ".ng").c_str(), ng.data(), 3 * n);
}

int main(int argc, char** argv)
{
    if (argc < 3) { std::cerr << "usage: gen_ref_golden <inputs dir> <outputs dir>\n"; return 2; }
    const std::string in_dir = argv[1], out_dir = argv[2];
    try {
        // ---- C1: 32 x 32 spherical on the 10 092-triangle cube (BASELINE.json configs[0]) ----
        {
            Reader in(in_dir + "/c1.b2ref");
            Writer out(out_dir + "/c1.b2ref");
            rm::EmbreeMapPtr map = rm::import_embree_map(in_dir + "/cube29.ply");                   // micp_localization.cpp:188
            const std::vector<float> m = in.f32("c1.model");                                         // phi_min, phi_inc, phi_n, theta_min, theta_inc, theta_n, range_min, range_max
            rm::SphericalModel model;                                                                // filled like rmcl_ros/src/util/conversions.cpp:22-34
            model.phi.min = m[0]; model.phi.inc = m[1]; model.phi.size = (uint32_t)m[2];
            model.theta.min = m[3]; model.theta.inc = m[4]; model.theta.size = (uint32_t)m[5];
            model.range.min = m[6]; model.range.max = m[7];
            run_micp<rmcl::RCCEmbreeSpherical>("c1", map, model, in, out);
            run_rays("c1rays", map, in, out);
        }
        // ---- a pinhole sensor in the small building (the C4 code path at a size the oracle and Embree both finish in a blink) ----
        {
            Reader in(in_dir + "/pin.b2ref");
            Writer out(out_dir + "/pin.b2ref");
            rm::EmbreeMapPtr map = rm::import_embree_map(in_dir + "/building60k.ply");
            const std::vector<float> m = in.f32("pin.model");                                        // width, height, fx, fy, cx, cy, range_min, range_max
            rm::PinholeModel model;                                                                  // conversions.cpp:48-60
            model.width = (uint32_t)m[0]; model.height = (uint32_t)m[1]; model.f[0] = m[2]; model.f[1] = m[3]; model.c[0] = m[4]; model.c[1] = m[5];
            model.range.min = m[6]; model.range.max = m[7];
            run_micp<rmcl::RCCEmbreePinhole>("pin", map, model, in, out);
        }
    } catch (const std::exception& e) {
        std::cerr << "gen_ref_golden: " << e.what() << "\n";
        return 1;
    }
    std::cout << "reference outputs written to " << out_dir << "\n";
    return 0;
}
