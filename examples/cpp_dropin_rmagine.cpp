// cpp_dropin_rmagine.cpp -- the drop-in classes used THROUGH THE REFERENCE'S OWN INTERFACE TYPES.  Compiled with -DRMCL_B200_WITH_RMAGINE
// -DRMCL_B200_WITH_RMCL_ROS against the reference's unmodified headers (rmcl/registration/Correspondences.hpp, rmcl_ros/rmcl/SensorUpdater.hpp)
// and an rmagine on the include path -- the real one, or the shape-mirroring stand-ins of tests/stubs in the authoring image.
//
// What the reference's node does with a backend object, done here with ours:
//   micp_localization.cpp:616-626   correspondences_ = std::make_shared<RCCOptixSpherical>(map)   (a shared_ptr<Correspondences_<VRAM_CUDA>>)
//   MICPSphericalSensorCPU.cpp:155-160   dynamic_pointer_cast<rm::ModelSetter<rm::SphericalModel>>(correspondences_)->setModel(model)
//   MICPSphericalSensorCUDA.cpp:230-232   dataset written through the public member `correspondences_->dataset`
//   MICPSensor.hpp:146-184   correspondences_->find(Tbm), ->computeCrossStatistics(T_snew_sold, cp), ->modelView()
//   rmcl_localization.cpp:538-541   sensor_updater_->update(poses(0, n), attrs(0, n))  through SensorUpdater<VRAM_CUDA>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <vector>

#include <rmcl_b200/rcc_b200.hpp>

namespace rm = rmagine;

int main(int argc, char** argv)
{
    if (argc < 2) { fprintf(stderr, "usage: cpp_dropin_rmagine <raw mesh file>\n"); return 2; }
    std::vector<float> V; std::vector<uint32_t> F;
    {
        FILE* fp = fopen(argv[1], "rb");
        uint32_t nv = 0, nf = 0;
        if (!fp || fread(&nv, 4, 1, fp) != 1 || fread(&nf, 4, 1, fp) != 1) { fprintf(stderr, "ERROR: cannot read %s\n", argv[1]); return 1; }
        V.resize(3 * (size_t)nv); F.resize(3 * (size_t)nf);
        if (fread(V.data(), 4, V.size(), fp) != V.size() || fread(F.data(), 4, F.size(), fp) != F.size()) { fprintf(stderr, "ERROR: short mesh file\n"); return 1; }
        fclose(fp);
    }
    try {
        auto map = std::make_shared<rmcl::B200Map>(V.data(), (uint32_t)(V.size() / 3), F.data(), (uint32_t)(F.size() / 3));
        rm::SphericalModel model;                               // vlp16_900(), range.min = 0
        model.phi = {-15.0f * (float)M_PI / 180.0f, 2.0f * (float)M_PI / 180.0f, 16};
        model.theta = {-(float)M_PI, 2.0f * (float)M_PI / 900.0f, 900};
        model.range = {0.0f, 130.0f};
        const size_t n = model.size();

        // ---- the backend object behind the reference's base-class pointer ----
        std::shared_ptr<rmcl::Correspondences_<rm::VRAM_CUDA>> correspondences_ = std::make_shared<rmcl::RCCB200Spherical>(map);
        if (auto model_setter = std::dynamic_pointer_cast<rm::ModelSetter<rm::SphericalModel>>(correspondences_)) model_setter->setModel(model);
        else { fprintf(stderr, "ERROR: no ModelSetter<SphericalModel>\n"); return 1; }
        correspondences_->setTsb(rm::Transform::Identity());
        correspondences_->params.max_dist = 1.0f;
        correspondences_->adaptive_max_dist_min = 0.15f;

        // the scan: simulated at identity with the v1 corrector (simulate(T, ranges), lidar_corrector_optix_benchmark.cpp:117)
        rmcl::SphereCorrectorB200 v1(map);
        v1.setTsb(rm::Transform::Identity()); v1.setModel(model);
        std::vector<float> ranges(n);
        v1.simulate(rm::Transform::Identity(), ranges.data());

        // dataset through the PUBLIC MEMBER, host loop of unpackMessage (MICPSphericalSensorCPU.cpp:193-228) + upload (…CUDA.cpp:230-232)
        rm::Memory<rm::Vector, rm::RAM> pts(n); rm::Memory<uint8_t, rm::RAM> mask(n);
        for (unsigned vid = 0; vid < model.getHeight(); vid++)
            for (unsigned hid = 0; hid < model.getWidth(); hid++) {
                const unsigned loc = model.getBufferId(vid, hid);
                const float r = ranges[loc];
                pts[loc] = model.getDirection(vid, hid) * r;
                mask[loc] = (r < model.range.min || r > model.range.max) ? 0 : 1;
            }
        correspondences_->dataset.points.resize(n);
        correspondences_->dataset.mask.resize(n);
        cudaMemcpy(correspondences_->dataset.points.raw(), pts.raw(), sizeof(rm::Vector) * n, cudaMemcpyHostToDevice);
        cudaMemcpy(correspondences_->dataset.mask.raw(), mask.raw(), n, cudaMemcpyHostToDevice);
        correspondences_->outdated = true;

        rm::Transform Tg = rm::Transform::Identity(); Tg.t.z = 0.2f;
        correspondences_->find(Tg);
        correspondences_->outdated = false;
        const rm::CrossStatistics cs = correspondences_->computeCrossStatistics(rm::Transform::Identity(), 0.0);
        const rm::Transform Tu = rmcl::umeyama_transform(cs);
        // inherited modelView() over the protected model_buffers_ the kernel wrote into
        auto mv = correspondences_->modelView();
        std::vector<uint8_t> hits(mv.mask.size());
        cudaMemcpy(hits.data(), mv.mask.raw(), hits.size(), cudaMemcpyDeviceToHost);
        size_t n_hits = 0; for (uint8_t h : hits) n_hits += h;
        printf("V2R n_meas %u cov_trace %.9g tz %.9g model_n %zu hits %zu dataset_n %zu\n", cs.n_meas, cs.covariance.trace(), Tu.t.z, mv.points.size(), n_hits,
               correspondences_->datasetView().points.size());

        // the same through the stand-alone setters of the same object type: identical numbers
        rmcl::RCCB200Spherical plain(map);
        plain.setTsb(rm::Transform::Identity()); plain.setModel(model); plain.params.max_dist = 1.0f;
        plain.setRanges(ranges.data(), n);
        plain.find(Tg);
        const rm::CrossStatistics cs2 = plain.computeCrossStatistics(rm::Transform::Identity(), 0.0);
        printf("V2P n_meas %u cov_trace %.9g\n", cs2.n_meas, cs2.covariance.trace());

        // v1 benchmark(): stage split
        v1.setInputData(ranges.data(), n);
        std::vector<rm::Transform> T_curr(16, Tg);
        const rmcl::BenchmarkResultB200 br = v1.benchmark(T_curr, 3);
        printf("BENCH sim %.6g red %.6g svd %.6g\n", br.sim, br.red, br.svd);

        // ---- particle filter through SensorUpdater<VRAM_CUDA> ----
        auto concrete = std::make_shared<rmcl::PCDSensorUpdaterB200>(map);
        std::shared_ptr<rmcl::SensorUpdater<rm::VRAM_CUDA>> sensor_updater_ = concrete;
        sensor_updater_->init();
        concrete->setTsb(rm::Transform::Identity());
        std::vector<rmcl::RangeMeasurement> beams;
        for (int k = 0; k < 8; k++) {
            const float th = 0.7f * k; rmcl::RangeMeasurement m{}; m.orig = {0, 0, 0}; m.dir = {std::cos(th), std::sin(th), 0.0f}; m.range = 10.0f; beams.push_back(m);
        }
        concrete->setBeams(beams);
        rm::Memory<rm::Transform, rm::RAM> poses(4); rm::Memory<rmcl::ParticleAttributes, rm::RAM> attrs(4);
        for (size_t i = 0; i < 4; i++) {
            poses[i] = rm::Transform::Identity();
            attrs[i].likelihood = rm::Gaussian1D::Identity(); attrs[i].likelihood.mean = 1.0f;
            for (unsigned k = 0; k < 6; k++) attrs[i].state_sigma(k, 0) = 0.1f;
        }
        poses[1].t.x = 1.0f; poses[2].t.y = -2.0f; poses[3].t.z = 0.5f;
        rm::Memory<rm::Transform, rm::VRAM_CUDA> poses_d(4); rm::Memory<rmcl::ParticleAttributes, rm::VRAM_CUDA> attrs_d(4);
        cudaMemcpy(poses_d.raw(), poses.raw(), sizeof(rm::Transform) * 4, cudaMemcpyHostToDevice);
        cudaMemcpy(attrs_d.raw(), attrs.raw(), sizeof(rmcl::ParticleAttributes) * 4, cudaMemcpyHostToDevice);
        sensor_updater_->update(poses_d(0, 4), attrs_d(0, 4));                                         // rmcl_localization.cpp:538-541
        cudaDeviceSynchronize();
        cudaMemcpy(attrs.raw(), attrs_d.raw(), sizeof(rmcl::ParticleAttributes) * 4, cudaMemcpyDeviceToHost);
        for (size_t i = 0; i < 4; i++) printf("PF %zu mean %.9g sigma %.9g n %u\n", i, attrs[i].likelihood.mean, attrs[i].likelihood.sigma, attrs[i].likelihood.n_meas);
        sensor_updater_->reset();
    } catch (const std::exception& e) {
        fprintf(stderr, "ERROR: %s\n", e.what());
        return 1;
    }
    return 0;
}
