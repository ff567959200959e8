// rcc_b200.hpp -- C++ host classes with the reference's names and signatures for the ray-casting-correspondence path, implemented over
// the C ABI (include/rmcl_b200.h).  These are what rmcl_ros' loadSensor / RmclNode would instantiate for a "b200" backend string
// (INTEGRATION.md).  Header-only; link with -lrmcl_b200.
//
//   rmcl::Correspondences_<MemT>            rmcl/include/rmcl/registration/Correspondences.hpp:16-88
//   rmcl::RCCB200{Spherical,Pinhole,O1Dn,OnDn}  twins of rmcl::RCCEmbree* / RCCOptix*  (RCCEmbree.hpp:18-83, RCCOptix.hpp:18-93)
//   rmcl::{Sphere,Pinhole,O1Dn,OnDn}CorrectorB200  v1 API used by rmcl_ros/src/benchmarks/lidar_corrector_{embree,optix}_benchmark.cpp:86-155
//   rmcl::PCDSensorUpdaterB200              rmcl_ros/include/rmcl_ros/rmcl/ParticleUpdater.hpp:39-43 + PCDSensorUpdaterEmbree.cpp:244-352
#pragma once
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../rmcl_b200.h"
#include "rmagine_compat.hpp"
#ifdef RMCL_B200_WITH_RMAGINE
// Built inside the reference's tree (or against an installed rmcl): the classes below DERIVE from the reference's own interface, so that
//   std::shared_ptr<rmcl::Correspondences_<rm::VRAM_CUDA>> correspondences_ = std::make_shared<rmcl::RCCB200Spherical>(map);
// compiles exactly like the OptiX line at rmcl_ros/src/nodes/micp_localization.cpp:616-626.
#include <rmagine/types/MemoryCuda.hpp>
#include <rmagine/simulation/SimulationResults.hpp>
#include <rmcl/registration/Correspondences.hpp>          // rmcl/include/rmcl/registration/Correspondences.hpp:16-88
#ifdef RMCL_B200_WITH_RMCL_ROS
#include <rmcl_ros/rmcl/SensorUpdater.hpp>                 // rmcl_ros/include/rmcl_ros/rmcl/SensorUpdater.hpp:18-42, ParticleUpdater.hpp:24-43
#include <rmcl_ros/rmcl/RangeMeasurement.hpp>              // rmcl_ros/include/rmcl_ros/rmcl/RangeMeasurement.hpp:10-50
#endif
#define B2_OVERRIDE override
#else
#define B2_OVERRIDE
#endif

namespace rmcl {

namespace rm = rmagine;

// O1Dn / OnDn tables are rm::Memory<Vector, RAM> in rmagine, std::vector in the stand-alone layer
template <typename C> inline auto b2_table_ptr(const C& c) -> decltype(c.raw()) { return c.raw(); }
template <typename T> inline const T* b2_table_ptr(const std::vector<T>& c) { return c.data(); }

inline void b2_check(int rc, const char* what)
{
    // error convention of the reference: std::runtime_error for map / configuration failures (micp_localization.cpp:124, PCDSensorUpdaterEmbree.cpp:165)
    if (rc != B2_OK) throw std::runtime_error(std::string(what) + ": " + b2_last_error());
}

// stands in for rm::EmbreeMap / rm::OptixMap; shared between sensors and plugins like the reference's map container (micp_localization.cpp:545)
class B200Map {
public:
    B200Map(const float* verts_xyz, uint32_t n_vertices, const uint32_t* faces_ijk, uint32_t n_faces, int device = 0, int build_mode = B2_BUILD_DEVICE_LBVH)
    { b2_check(b2_mesh_create(verts_xyz, n_vertices, faces_ijk, n_faces, device, build_mode, &m_), "B200Map"); }
    // rm::import_embree_map(file) twin (micp_localization.cpp:188): .ply / .obj
    explicit B200Map(const std::string& mesh_file, int device = 0, int build_mode = B2_BUILD_DEVICE_LBVH)
    { b2_check(b2_mesh_create_from_file(mesh_file.c_str(), device, build_mode, &m_), "B200Map"); }
    ~B200Map() { b2_mesh_destroy(m_); }
    B200Map(const B200Map&) = delete; B200Map& operator=(const B200Map&) = delete;
    // the map's vertices moved (same faces): refit the resident BVH; dependants set their `outdated` flag like Correspondences.hpp:26-31
    void refit(const float* verts_xyz, uint32_t n_vertices, bool on_device = false) { b2_check(b2_mesh_refit(m_, verts_xyz, n_vertices, on_device ? 1 : 0), "B200Map::refit"); }
    b2_mesh* handle() const { return m_; }
    b2_mesh_info info() const { b2_mesh_info i; b2_check(b2_mesh_get_info(m_, &i), "B200Map::info"); return i; }
private:
    b2_mesh* m_ = nullptr;
};
using B200MapPtr = std::shared_ptr<B200Map>;

// device-resident point cloud views (rm::PointCloudView_<VRAM_CUDA>)
struct PointCloudViewB200 { rm::MemoryView<rm::Vector3f, rm::VRAM_CUDA> points; rm::MemoryView<uint8_t, rm::VRAM_CUDA> mask; rm::MemoryView<rm::Vector3f, rm::VRAM_CUDA> normals; };

#if !(defined(RMCL_B200_WITH_RMAGINE) && defined(RMCL_B200_WITH_RMCL_ROS))
struct ParticleAttributes { rm::Gaussian1D likelihood; float state_sigma[6]; };                // ParticleAttributes.hpp:18-32
struct RangeMeasurement { rm::Vector3f orig, dir; float range; rm::Matrix3x3 cov; };            // RangeMeasurement.hpp:10-21
struct ParticleUpdateResults {};
struct ParticleUpdateConfig {};
#endif
static_assert(sizeof(ParticleAttributes) == 36 && sizeof(RangeMeasurement) == 64, "rmcl layouts");

// --- Correspondences_ interface, device flavour --------------------------------------------------------------------------------
class CorrespondencesB200
#ifdef RMCL_B200_WITH_RMAGINE
    : public Correspondences_<rm::VRAM_CUDA>            // public: params, adaptive_max_dist_min, dataset, outdated; protected: model_buffers_, Tsb_
#endif
{
public:
#ifndef RMCL_B200_WITH_RMAGINE
    rm::UmeyamaReductionConstraints params{1.0f};      // Correspondences.hpp:22
    float adaptive_max_dist_min = 0.15f;               // :23
    bool outdated = true;                              // :31
#endif

    explicit CorrespondencesB200(B200MapPtr map) : map_(std::move(map))
    {
        if (!map_) throw std::runtime_error("NO MAP");                                           // PCDSensorUpdaterOptix.cpp:179-185
        b2_check(b2_rcc_create(map_->handle(), &h_), "RCCB200");
#ifdef RMCL_B200_WITH_RMAGINE
        params.max_dist = 1.0f; adaptive_max_dist_min = 0.15f;
#endif
    }
    virtual ~CorrespondencesB200() { b2_rcc_destroy(h_); }
    CorrespondencesB200(const CorrespondencesB200&) = delete; CorrespondencesB200& operator=(const CorrespondencesB200&) = delete;

    virtual void setTsb(const rm::Transform& Tsb) B2_OVERRIDE { Tsb_ = Tsb; b2_check(b2_rcc_set_tsb(h_, tf(&Tsb)), "setTsb"); }     // :33-36
    // Stand-alone builds: the public `dataset` field of the reference becomes two setters (the buffers live in HBM, owned by the handle).
    // rmagine builds: the inherited `dataset` member (rm::PointCloud_<VRAM_CUDA>, written directly by the sensor wrapper like
    // MICPSphericalSensorCUDA.cpp:230-232) is what find / computeCrossStatistics / correctOnce read -- bound, not copied; the setters below
    // stay available as the device-side unpackMessage and then take precedence until `dataset` is resized again.
    void setDataset(const rm::Vector3f* points, const uint8_t* mask, size_t n, bool on_device = false)
    { b2_check(b2_rcc_set_dataset(h_, reinterpret_cast<const float*>(points), mask, (uint32_t)n, on_device ? 1 : 0), "setDataset"); outdated = true; own_dataset_ = true; }
    void setRanges(const float* ranges, size_t n, bool on_device = false)                          // MICPSphericalSensorCPU.cpp:181-233 on the device
    { b2_check(b2_rcc_set_ranges(h_, ranges, (uint32_t)n, on_device ? 1 : 0), "setRanges"); outdated = true; own_dataset_ = true; }

    virtual void find(const rm::Transform& Tbm_est) B2_OVERRIDE                                    // :42-44
    { sync_params(); bind_members(); b2_check(b2_rcc_find(h_, tf(&Tbm_est)), "find"); outdated = false; }

    virtual rm::CrossStatistics computeCrossStatistics(const rm::Transform& T_snew_sold, double convergence_progress = 0.0) const B2_OVERRIDE   // :75-77
    {
        sync_params(); const_cast<CorrespondencesB200*>(this)->bind_members();
        rm::CrossStatistics out;
        b2_check(b2_rcc_cross_statistics(h_, tf(&T_snew_sold), convergence_progress, reinterpret_cast<b2_cross_stats*>(&out)), "computeCrossStatistics");
        return out;
    }
#ifndef RMCL_B200_WITH_RMAGINE            // rmagine builds inherit modelView() / datasetView() over model_buffers_ / dataset (Correspondences.hpp:47-62)
    PointCloudViewB200 modelView()                                                                 // :47-54
    {
        float *p = nullptr, *nr = nullptr; uint8_t* hi = nullptr; uint32_t n = 0;
        b2_check(b2_rcc_model_view(h_, &p, &nr, &hi, nullptr, nullptr, &n), "modelView");
        return PointCloudViewB200{{reinterpret_cast<rm::Vector3f*>(p), n}, {hi, n}, {reinterpret_cast<rm::Vector3f*>(nr), n}};
    }
    PointCloudViewB200 datasetView()                                                               // :56-62
    {
        float* p = nullptr; uint8_t* m = nullptr; uint32_t n = 0;
        b2_check(b2_rcc_dataset_view(h_, &p, &m, &n), "datasetView");
        return PointCloudViewB200{{reinterpret_cast<rm::Vector3f*>(p), n}, {m, n}, {nullptr, 0}};
    }
#endif
    // one MICPLocalizationNode::correctOnce for this sensor on the device (micp_localization.cpp:899-984)
    rm::Transform correctOnce(const rm::Transform& Tom, const rm::Transform& Tbo, unsigned iterations = 5, double convergence_progress = 0.0,
                              rm::Transform* T_onew_oold = nullptr, rm::CrossStatistics* Cmerged_o = nullptr)
    {
        sync_params(); bind_members();
        rm::Transform out;
        b2_check(b2_rcc_correct_once(h_, tf(&Tom), tf(&Tbo), iterations, convergence_progress, reinterpret_cast<b2_transform*>(&out),
                                     reinterpret_cast<b2_transform*>(T_onew_oold), reinterpret_cast<b2_cross_stats*>(Cmerged_o)), "correctOnce");
        outdated = false;
        return out;
    }
    // same, with the scan's ranges on the HOST (end-to-end entry): a pinned buffer is read in place by the kernel, a pageable one is uploaded
    rm::Transform correctOnceRanges(const float* ranges_host, size_t n, const rm::Transform& Tom, const rm::Transform& Tbo, unsigned iterations = 5,
                                    double convergence_progress = 0.0, rm::Transform* T_onew_oold = nullptr, rm::CrossStatistics* Cmerged_o = nullptr)
    {
        sync_params(); own_dataset_ = true;
        rm::Transform out;
        b2_check(b2_rcc_correct_once_ranges(h_, ranges_host, (uint32_t)n, tf(&Tom), tf(&Tbo), iterations, convergence_progress, reinterpret_cast<b2_transform*>(&out),
                                            reinterpret_cast<b2_transform*>(T_onew_oold), reinterpret_cast<b2_cross_stats*>(Cmerged_o)), "correctOnceRanges");
        outdated = false;
        return out;
    }
    // ScanMapSegmentationEmbreeNode::scanCB classification (scan_map_segmentation_embree.cpp:110-187): after setRanges(real scan) + find(pose)
    struct Segmentation { std::vector<rm::Vector3f> outlier_scan, outlier_map; };
    Segmentation segment(float min_dist_outlier_scan, float min_dist_outlier_map)
    {
        uint32_t n = 0, ns = 0, nm = 0;
        b2_check(b2_rcc_model_view(h_, nullptr, nullptr, nullptr, nullptr, nullptr, &n), "segment");
        Segmentation r; r.outlier_scan.resize(n); r.outlier_map.resize(n);
        b2_check(b2_rcc_segment(h_, min_dist_outlier_scan, min_dist_outlier_map, reinterpret_cast<float*>(r.outlier_scan.data()), n, &ns,
                                reinterpret_cast<float*>(r.outlier_map.data()), n, &nm, nullptr), "segment");
        r.outlier_scan.resize(ns); r.outlier_map.resize(nm);
        return r;
    }
    void setStream(void* cuda_stream) { b2_check(b2_rcc_set_stream(h_, cuda_stream), "setStream"); }
    b2_rcc* handle() const { return h_; }
    void noteModelSize(uint32_t width, uint32_t height) { note_model(width, height); }      // for model tables set through the C ABI directly (rmcl_msgs_adapters.hpp)

protected:
    static const b2_transform* tf(const rm::Transform* T) { return reinterpret_cast<const b2_transform*>(T); }
    void sync_params() const { b2_rcc_set_params(h_, params.max_dist, adaptive_max_dist_min); }
    // model size of the sensor model last given to setModel (RCCOptix keeps a model_cache_ for the same purpose, RCCOptix.hpp:39)
    void note_model(uint32_t width, uint32_t height) { model_w_ = width; model_h_ = height; }
#ifdef RMCL_B200_WITH_RMAGINE
    // Point the handle at the members the reference's interface exposes: model_buffers_ grown like RCCOptix.cpp:30-40 and bound as the find()
    // output, `dataset` bound as the reduction input (unless the device-side unpack of setRanges / setDataset is in charge).
    void bind_members()
    {
        const size_t n = size_t(model_w_) * model_h_;
        if (n > 0) {
            if (n > model_buffers_.points.size()) rm::resize_memory_bundle<rm::VRAM_CUDA>(model_buffers_, model_h_, model_w_, 1);
            b2_check(b2_rcc_bind_model_buffers(h_, reinterpret_cast<float*>(model_buffers_.points.raw()), reinterpret_cast<float*>(model_buffers_.normals.raw()),
                                               model_buffers_.hits.raw(), (uint32_t)model_buffers_.points.size()), "bind model_buffers_");
        }
        if (dataset.points.size() > 0 && (!own_dataset_ || dataset.points.raw() != bound_dataset_ || dataset.points.size() != bound_n_)) {
            if (dataset.mask.size() < dataset.points.size()) throw std::runtime_error("dataset.mask smaller than dataset.points");
            b2_check(b2_rcc_bind_dataset(h_, reinterpret_cast<const float*>(dataset.points.raw()), dataset.mask.raw(), (uint32_t)dataset.points.size()), "bind dataset");
            bound_dataset_ = dataset.points.raw(); bound_n_ = dataset.points.size(); own_dataset_ = false;
        }
    }
    const void* bound_dataset_ = nullptr; size_t bound_n_ = 0;
#else
    void bind_members() {}
    rm::Transform Tsb_ = rm::Transform::Identity();
#endif
    bool own_dataset_ = false;
    uint32_t model_w_ = 0, model_h_ = 0;
    B200MapPtr map_;
    b2_rcc* h_ = nullptr;
};

class RCCB200Spherical : public CorrespondencesB200, public rm::ModelSetter<rm::SphericalModel> {
public:
    explicit RCCB200Spherical(B200MapPtr map) : CorrespondencesB200(std::move(map)) {}
    void setModel(const rm::SphericalModel& m) override                                           // RCCEmbree.cpp:21-24
    {
        b2_spherical_model s{m.phi.min, m.phi.inc, m.phi.size, m.theta.min, m.theta.inc, m.theta.size, m.range.min, m.range.max};
        b2_check(b2_rcc_set_model_spherical(h_, &s), "setModel"); note_model(m.theta.size, m.phi.size);
    }
};
class RCCB200Pinhole : public CorrespondencesB200, public rm::ModelSetter<rm::PinholeModel> {
public:
    explicit RCCB200Pinhole(B200MapPtr map) : CorrespondencesB200(std::move(map)) {}
    void setModel(const rm::PinholeModel& m) override                                             // RCCEmbree.cpp:53-56
    {
        b2_pinhole_model p{m.width, m.height, m.f[0], m.f[1], m.c[0], m.c[1], m.range.min, m.range.max};
        b2_check(b2_rcc_set_model_pinhole(h_, &p), "setModel"); note_model(m.width, m.height);
    }
};
class RCCB200O1Dn : public CorrespondencesB200, public rm::ModelSetter<rm::O1DnModel> {
public:
    explicit RCCB200O1Dn(B200MapPtr map) : CorrespondencesB200(std::move(map)) {}
    void setModel(const rm::O1DnModel& m) override                                                // RCCEmbree.cpp:84-87
    { b2_check(b2_rcc_set_model_o1dn(h_, m.width, m.height, &m.orig.x, reinterpret_cast<const float*>(b2_table_ptr(m.dirs)), m.range.min, m.range.max), "setModel"); note_model(m.width, m.height); }
};
class RCCB200OnDn : public CorrespondencesB200, public rm::ModelSetter<rm::OnDnModel> {
public:
    explicit RCCB200OnDn(B200MapPtr map) : CorrespondencesB200(std::move(map)) {}
    void setModel(const rm::OnDnModel& m) override                                                // RCCEmbree.cpp:116-119
    {
        b2_check(b2_rcc_set_model_ondn(h_, m.width, m.height, reinterpret_cast<const float*>(b2_table_ptr(m.origs)), reinterpret_cast<const float*>(b2_table_ptr(m.dirs)),
                                       m.range.min, m.range.max), "setModel");
        note_model(m.width, m.height);
    }
};

// rmcl::CPCEmbree twin (rmcl/include/rmcl/registration/CPCEmbree.hpp:20-54): closest-point correspondences on the same map BVH.
// find() = one closest-point query per dataset point (CPCEmbree.cpp:17-43); no sensor model.
class CPCB200 : public CorrespondencesB200 {
public:
    explicit CPCB200(B200MapPtr map) : CorrespondencesB200(std::move(map)) { b2_check(b2_rcc_set_correspondence_type(h_, B2_CORR_CPC), "CPCB200"); }
#ifdef RMCL_B200_WITH_RMAGINE
    void find(const rm::Transform& Tbm_est) override                     // CPCEmbree.cpp:17-43: one model entry per dataset point
    { note_model((uint32_t)dataset.points.size(), dataset.points.size() ? 1u : 0u); CorrespondencesB200::find(Tbm_est); }
#endif
};

// --- v1 batched corrector API (lidar_corrector_embree_benchmark.cpp:86-133) -------------------------------------------------------
struct CorrectionResultsB200 { std::vector<rm::Transform> Tdelta; std::vector<uint32_t> Ncorr; };
struct BenchmarkResultB200 { double sim = 0.0, red = 0.0, svd = 0.0; };                          // lidar_corrector_optix_benchmark.cpp:143-155
template <typename RCC> class CorrectorB200 : public RCC {
public:
    using RCC::RCC;
    // v1 simulate(Tbm, ranges): trace the sensor at Tbm and hand out the simulated ranges (HOST buffer of model-size floats; :117)
    void simulate(const rm::Transform& Tbm, float* ranges_host)
    {
        this->find(Tbm);
        uint32_t n = 0;
        b2_check(b2_rcc_model_view(this->h_, nullptr, nullptr, nullptr, nullptr, nullptr, &n), "simulate");
        b2_check(b2_rcc_download_model(this->h_, nullptr, nullptr, nullptr, nullptr, ranges_host), "simulate");
    }
    // v1 benchmark(Tbm, Nruns) -> {sim, red, svd} seconds: the stages of correct() run unfused and timed separately (:143-155)
    BenchmarkResultB200 benchmark(const std::vector<rm::Transform>& Tbm, size_t Nruns = 100)
    {
        this->sync_params();
        BenchmarkResultB200 r;
        b2_check(b2_rcc_benchmark_batch(this->h_, reinterpret_cast<const b2_transform*>(Tbm.data()), (uint32_t)Tbm.size(), (uint32_t)Nruns, &r.sim, &r.red, &r.svd), "benchmark");
        return r;
    }
    void setInputData(const float* ranges, size_t n, bool on_device = false) { this->setRanges(ranges, n, on_device); }       // :118
    CorrectionResultsB200 correct(const std::vector<rm::Transform>& Tbm)                                                       // :127-133
    {
        this->sync_params();
        CorrectionResultsB200 r; r.Tdelta.resize(Tbm.size()); r.Ncorr.resize(Tbm.size());
        b2_check(b2_rcc_correct_batch(this->h_, reinterpret_cast<const b2_transform*>(Tbm.data()), (uint32_t)Tbm.size(), 0,
                                      reinterpret_cast<b2_transform*>(r.Tdelta.data()), r.Ncorr.data(), nullptr, 0), "correct");
        return r;
    }
};
using SphereCorrectorB200 = CorrectorB200<RCCB200Spherical>;
using PinholeCorrectorB200 = CorrectorB200<RCCB200Pinhole>;
using O1DnCorrectorB200 = CorrectorB200<RCCB200O1Dn>;
using OnDnCorrectorB200 = CorrectorB200<RCCB200OnDn>;

inline rm::Transform umeyama_transform(const rm::CrossStatistics& s, int device = 0)              // micp_localization.cpp:952-953
{
    rm::Transform T;
    b2_check(b2_umeyama_batch(reinterpret_cast<const b2_cross_stats*>(&s), 1, reinterpret_cast<b2_transform*>(&T), 0, device, nullptr), "umeyama_transform");
    return T;
}

// --- particle filter sensor update ------------------------------------------------------------------------------------------------
class PCDSensorUpdaterB200
#if defined(RMCL_B200_WITH_RMAGINE) && defined(RMCL_B200_WITH_RMCL_ROS)
    : public SensorUpdater<rm::VRAM_CUDA>               // = SensorUpdaterBase (init / reset) + ParticleUpdater<VRAM_CUDA> (update), SensorUpdater.hpp:37-42
#define B2_PF_OVERRIDE override
#else
#define B2_PF_OVERRIDE
#endif
{
public:
    // SensorUpdaterBase::init / reset (SensorUpdater.hpp:18-25).  The Embree updater loads its map in init() (PCDSensorUpdaterEmbree.cpp:136-189);
    // here the map arrives through the constructor, so init() only checks it and reset() drops the beams of the last cloud.
    void init() B2_PF_OVERRIDE { if (!map_) throw std::runtime_error("NO MAP"); }
    void reset() B2_PF_OVERRIDE { beams_.clear(); }
    b2_pf_params config{2.0f, 100.0f, 100.0f, 0.0f, 0.05f, 80.0f, 0, 0};                         // PCDSensorUpdaterEmbree.cpp:122-134
    explicit PCDSensorUpdaterB200(B200MapPtr map) : map_(std::move(map))
    {
        if (!map_) throw std::runtime_error("NO MAP");
        b2_check(b2_pf_create(map_->handle(), &h_), "PCDSensorUpdaterB200");
    }
    ~PCDSensorUpdaterB200() { b2_pf_destroy(h_); }
    void setTsb(const rm::Transform& Tsb) { Tsb_ = Tsb; }
    // beams replace the random_device sampling of :276-327 (quirk D5): the caller samples the cloud and passes RangeMeasurements
    void setBeams(const std::vector<RangeMeasurement>& beams) { beams_ = beams; }
    // how rays are mapped to lanes (a schedule, results identical): 0 beams of one particle, 1 particles, 2 particles sorted by pose, 3 (default) by timing
    void setMapping(int mode) { b2_check(b2_pf_set_mapping(h_, mode), "setMapping"); }
    // ParticleUpdater<RAM>::update
    ParticleUpdateResults update(rm::MemoryView<rm::Transform, rm::RAM> poses, rm::MemoryView<ParticleAttributes, rm::RAM> attrs, const ParticleUpdateConfig& = {})
    {
        b2_check(b2_pf_sensor_update_host(h_, reinterpret_cast<const b2_transform*>(poses.raw()), reinterpret_cast<b2_particle_attr*>(attrs.raw()), (uint32_t)poses.size(),
                                          reinterpret_cast<const b2_transform*>(&Tsb_), reinterpret_cast<const b2_range_meas*>(beams_.data()), (uint32_t)beams_.size(), &config), "update");
        return {};
    }
    // ParticleUpdater<VRAM_CUDA>::update (ParticleUpdater.hpp:39-43)
    ParticleUpdateResults update(rm::MemoryView<rm::Transform, rm::VRAM_CUDA> poses, rm::MemoryView<ParticleAttributes, rm::VRAM_CUDA> attrs, const ParticleUpdateConfig& = {}) B2_PF_OVERRIDE
    {
        b2_check(b2_pf_sensor_update(h_, reinterpret_cast<const b2_transform*>(poses.raw()), reinterpret_cast<b2_particle_attr*>(attrs.raw()), (uint32_t)poses.size(),
                                     reinterpret_cast<const b2_transform*>(&Tsb_), reinterpret_cast<const b2_range_meas*>(beams_.data()), (uint32_t)beams_.size(), &config), "update");
        return {};
    }
    // rest of the cycle on the device (particles stay in HBM): TFMotionUpdaterGPU (particle_motion.cu:36-46) and compute_stats (resampling.cu:84-92)
    void motionUpdate(rm::MemoryView<rm::Transform, rm::VRAM_CUDA> poses, rm::MemoryView<ParticleAttributes, rm::VRAM_CUDA> attrs, const rm::Transform& T_bnew_bold, double forget_rate,
                      bool check_collision = false)       // true: the wall check of TFMotionUpdaterCPU (TFMotionUpdaterCPU.cpp:205-216)
    {
        b2_check(b2_pf_motion_update(h_, reinterpret_cast<b2_transform*>(poses.raw()), reinterpret_cast<b2_particle_attr*>(attrs.raw()), (uint32_t)poses.size(),
                                     reinterpret_cast<const b2_transform*>(&T_bnew_bold), forget_rate, check_collision ? 1 : 0), "motionUpdate");
    }
    struct SimpleLikelihoodStats { float sum = 0.0f; float max = -1.0f; };                       // resampling.cuh:26-30
    SimpleLikelihoodStats computeStats(rm::MemoryView<ParticleAttributes, rm::VRAM_CUDA> attrs)
    {
        SimpleLikelihoodStats s;
        b2_check(b2_pf_likelihood_stats(h_, reinterpret_cast<const b2_particle_attr*>(attrs.raw()), (uint32_t)attrs.size(), &s.sum, &s.max), "computeStats");
        return s;
    }
    // GladiatorResamplerGPU::resample (resampling.cu:201-221; config GladiatorResamplerConfig.hpp:7-20).  `poses/attrs` = all particles
    // opponents are drawn from; the outputs receive champions first .. first+poses_new.size()-1 (single GPU: first = 0, same sizes).
    b2_gladiator_config resampler_config{0.03f, 0.03f, 0.0f, 0.0f, 0.0f, 0.01f, 0.3f, 0.2f};
    void resample(rm::MemoryView<rm::Transform, rm::VRAM_CUDA> poses, rm::MemoryView<ParticleAttributes, rm::VRAM_CUDA> attrs,
                  rm::MemoryView<rm::Transform, rm::VRAM_CUDA> poses_new, rm::MemoryView<ParticleAttributes, rm::VRAM_CUDA> attrs_new,
                  uint64_t seed = 1234, uint32_t step = 0, uint32_t first = 0)
    {
        b2_check(b2_pf_resample_gladiator(h_, reinterpret_cast<const b2_transform*>(poses.raw()), reinterpret_cast<const b2_particle_attr*>(attrs.raw()), (uint32_t)poses.size(),
                                          first, (uint32_t)poses_new.size(), reinterpret_cast<b2_transform*>(poses_new.raw()), reinterpret_cast<b2_particle_attr*>(attrs_new.raw()),
                                          &resampler_config, seed, step, nullptr, nullptr), "resample");
    }
private:
    B200MapPtr map_;
    b2_pf* h_ = nullptr;
    rm::Transform Tsb_ = rm::Transform::Identity();
    std::vector<RangeMeasurement> beams_;
};

}  // namespace rmcl
