// cpp_dropin.cpp -- the legacy benchmark scenario (rmcl_ros/src/benchmarks/lidar_corrector_embree_benchmark.cpp:73-135) written against
// the drop-in C++ classes of include/rmcl_b200/rcc_b200.hpp: sphere map, vlp16_900 with range.min = 0, T_curr = I with z += 0.2,
// ten correct() calls; plus one v2-style find / computeCrossStatistics / umeyama round and a particle update.
// Prints machine-readable lines that tests/test_gpu_cpp.py compares with the Python path.
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <rmcl_b200/rcc_b200.hpp>
#include <rmcl_b200/rmcl_msgs_adapters.hpp>

namespace rm = rmagine;

// plain stand-ins for the rosidl-generated rmcl_msgs structs (same field names as rmcl_msgs/msg/*.msg): the adapters are duck-typed
namespace mock_msgs {
struct ScanInfo { float phi_min, phi_inc; uint32_t phi_n; float theta_min, theta_inc; uint32_t theta_n; float range_min, range_max; };
struct RangeData { std::vector<float> ranges; std::vector<bool> mask; };
struct Scan { ScanInfo info; RangeData data; };
struct Point32 { float x, y, z; };
struct O1DnInfo { uint32_t width, height; float range_min, range_max; Point32 orig; std::vector<Point32> dirs; };
struct O1Dn { O1DnInfo info; RangeData data; };
struct MICPSensorStats { uint32_t total_measurements, valid_measurements, valid_matches; float cov_trace; };
}

static void make_sphere(unsigned A, unsigned B, float radius, std::vector<float>& V, std::vector<uint32_t>& F)
{
    const double eps = 1e-3, pi = 3.14159265358979323846;
    for (unsigned i = 0; i <= A; i++)
        for (unsigned j = 0; j <= B; j++) {
            const double pol = eps + (pi - 2 * eps) * i / A, az = 2 * pi * j / B;
            V.push_back((float)(radius * std::sin(pol) * std::cos(az))); V.push_back((float)(radius * std::sin(pol) * std::sin(az))); V.push_back((float)(radius * std::cos(pol)));
        }
    for (unsigned i = 0; i < A; i++)
        for (unsigned j = 0; j < B; j++) {
            const uint32_t a = i * (B + 1) + j, b = a + 1, c = a + (B + 1), d = c + 1;
            F.insert(F.end(), {a, c, d, a, d, b});
        }
}

int main(int argc, char** argv)
{
    const unsigned A = argc > 1 ? (unsigned)atoi(argv[1]) : 40, B = argc > 2 ? (unsigned)atoi(argv[2]) : 60;
    const size_t Nposes = argc > 3 ? (size_t)atoi(argv[3]) : 100;
    std::vector<float> V; std::vector<uint32_t> F;
    if (argc > 4) {                                             // raw mesh file: u32 nv, u32 nf, nv*3 f32, nf*3 u32 (written by the test)
        FILE* fp = fopen(argv[4], "rb");
        uint32_t nv = 0, nf = 0;
        if (!fp || fread(&nv, 4, 1, fp) != 1 || fread(&nf, 4, 1, fp) != 1) { fprintf(stderr, "ERROR: cannot read %s\n", argv[4]); return 1; }
        V.resize(3 * (size_t)nv); F.resize(3 * (size_t)nf);
        if (fread(V.data(), 4, V.size(), fp) != V.size() || fread(F.data(), 4, F.size(), fp) != F.size()) { fprintf(stderr, "ERROR: short mesh file\n"); return 1; }
        fclose(fp);
    } else make_sphere(A, B, 10.0f, V, F);
    try {
        auto map = std::make_shared<rmcl::B200Map>(V.data(), (uint32_t)(V.size() / 3), F.data(), (uint32_t)(F.size() / 3));
        rmcl::SphereCorrectorB200 correct(map);
        correct.setTsb(rm::Transform::Identity());
        rm::SphericalModel model;                               // vlp16_900()
        model.phi = {-15.0f * (float)M_PI / 180.0f, 2.0f * (float)M_PI / 180.0f, 16};
        model.theta = {-(float)M_PI, 2.0f * (float)M_PI / 900.0f, 900};
        model.range = {0.0f, 130.0f};                           // model.range.min = 0.0 (:90)
        correct.setModel(model);

        // simulate the data that would be recorded at destination (:117-118): find at identity, ranges back as input data
        correct.find(rm::Transform::Identity());
        std::vector<float> ranges(model.size());
        b2_rcc_download_model(correct.handle(), nullptr, nullptr, nullptr, nullptr, ranges.data());
        correct.setInputData(ranges.data(), ranges.size());

        std::vector<rm::Transform> T_curr(Nposes, rm::Transform::Identity());
        for (auto& T : T_curr) T.t.z += 0.2f;
        const auto t0 = std::chrono::steady_clock::now();
        for (int run = 0; run < 10; run++) {
            auto res = correct.correct(T_curr);
            for (size_t i = 0; i < Nposes; i++) {               // T_curr = multNxN(T_curr, Tdelta) for pure translations + tiny rotations: compose
                const rm::Transform& a = T_curr[i]; const rm::Transform& d = res.Tdelta[i];
                // quaternion product and rotated translation (host convenience)
                rm::Quaternion q{a.R.w * d.R.x + a.R.x * d.R.w + a.R.y * d.R.z - a.R.z * d.R.y, a.R.w * d.R.y - a.R.x * d.R.z + a.R.y * d.R.w + a.R.z * d.R.x,
                                 a.R.w * d.R.z + a.R.x * d.R.y - a.R.y * d.R.x + a.R.z * d.R.w, a.R.w * d.R.w - a.R.x * d.R.x - a.R.y * d.R.y - a.R.z * d.R.z};
                const float x = d.t.x, y = d.t.y, z = d.t.z, qx = a.R.x, qy = a.R.y, qz = a.R.z, qw = a.R.w;
                const float tx = 2 * (qy * z - qz * y), ty = 2 * (qz * x - qx * z), tz = 2 * (qx * y - qy * x);
                rm::Vector3f t{x + qw * tx + (qy * tz - qz * ty) + a.t.x, y + qw * ty + (qz * tx - qx * tz) + a.t.y, z + qw * tz + (qx * ty - qy * tx) + a.t.z};
                T_curr[i] = rm::Transform{q, t, 0};
            }
            if (run == 0) printf("RUN0 ncorr %u tdelta_z %.9g\n", res.Ncorr[0], res.Tdelta[0].t.z);
        }
        const double el = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() / 10.0;
        printf("FINAL z %.9g\n", T_curr[0].t.z);
        printf("RUNTIME %zu,%zu,%.6f\n", F.size() / 3, Nposes, el);

        // v2 style: find -> computeCrossStatistics -> umeyama
        rm::Transform Tg = rm::Transform::Identity(); Tg.t.z = 0.2f;
        correct.params.max_dist = 1.0f;
        correct.find(Tg);
        const rm::CrossStatistics cs = correct.computeCrossStatistics(rm::Transform::Identity());
        const rm::Transform Tu = rmcl::umeyama_transform(cs);
        printf("V2 n_meas %u cov_trace %.9g tz %.9g\n", cs.n_meas, cs.covariance.trace(), Tu.t.z);

        // the same v2 round fed from rmcl_msgs-shaped messages (rmcl_msgs_adapters.hpp): Scan for the spherical sensor, O1Dn with the same rays
        {
            mock_msgs::Scan msg;
            msg.info = {model.phi.min, model.phi.inc, model.phi.size, model.theta.min, model.theta.inc, model.theta.size, model.range.min, model.range.max};
            msg.data.ranges = ranges;
            rmcl::RCCB200Spherical rcc(map);
            rcc.setTsb(rm::Transform::Identity()); rcc.params.max_dist = 1.0f;
            rmcl::b200::unpackMessage(rcc, msg);
            rcc.find(Tg);
            const rm::CrossStatistics c1 = rcc.computeCrossStatistics(rm::Transform::Identity());
            mock_msgs::O1Dn omsg;
            omsg.info.width = model.theta.size; omsg.info.height = model.phi.size; omsg.info.range_min = model.range.min; omsg.info.range_max = model.range.max; omsg.info.orig = {0.f, 0.f, 0.f};
            for (uint32_t vid = 0; vid < model.phi.size; vid++)
                for (uint32_t hid = 0; hid < model.theta.size; hid++) {
                    const float ph = model.phi.min + (float)vid * model.phi.inc, th = model.theta.min + (float)hid * model.theta.inc;
                    omsg.info.dirs.push_back({cosf(ph) * cosf(th), cosf(ph) * sinf(th), sinf(ph)});
                }
            omsg.data.ranges = ranges;
            rmcl::RCCB200O1Dn orcc(map);
            orcc.setTsb(rm::Transform::Identity()); orcc.params.max_dist = 1.0f;
            rmcl::b200::unpackMessage(orcc, omsg);
            orcc.find(Tg);
            const rm::CrossStatistics c2 = orcc.computeCrossStatistics(rm::Transform::Identity());
            mock_msgs::MICPSensorStats st{};
            rmcl::b200::fillSensorStats(st, c1, (uint32_t)ranges.size(), (uint32_t)ranges.size());
            printf("MSG n_meas %u cov_trace %.9g o1dn_n_meas %u o1dn_cov_trace %.9g stats_matches %u\n", c1.n_meas, c1.covariance.trace(), c2.n_meas, c2.covariance.trace(), st.valid_matches);
        }

        // particle update with a handful of beams
        rmcl::PCDSensorUpdaterB200 up(map);
        up.setTsb(rm::Transform::Identity());
        std::vector<rmcl::RangeMeasurement> beams;
        for (int k = 0; k < 8; k++) {
            const float th = 0.7f * k; rmcl::RangeMeasurement m{}; m.orig = {0, 0, 0}; m.dir = {std::cos(th), std::sin(th), 0.0f}; m.range = 10.0f; beams.push_back(m);
        }
        up.setBeams(beams);
        std::vector<rm::Transform> poses(4, rm::Transform::Identity());
        poses[1].t.x = 1.0f; poses[2].t.y = -2.0f; poses[3].t.z = 0.5f;
        std::vector<rmcl::ParticleAttributes> attrs(4);
        for (auto& a : attrs) { a.likelihood = rm::Gaussian1D::Identity(); a.likelihood.mean = 1.0f; for (float& s : a.state_sigma) s = 0.1f; }
        up.update(rm::MemoryView<rm::Transform, rm::RAM>{poses.data(), poses.size()}, rm::MemoryView<rmcl::ParticleAttributes, rm::RAM>{attrs.data(), attrs.size()});
        for (size_t i = 0; i < attrs.size(); i++) printf("PF %zu mean %.9g sigma %.9g n %u\n", i, attrs[i].likelihood.mean, attrs[i].likelihood.sigma, attrs[i].likelihood.n_meas);
    } catch (const std::exception& e) {
        fprintf(stderr, "ERROR: %s\n", e.what());
        return 1;
    }
    return 0;
}
