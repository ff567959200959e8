/*
 * rmcl_b200.h -- C ABI of the B200-native ray-casting-correspondence library (librmcl_b200.so).
 *
 * This is the drop-in boundary for ONE hot path of uos/rmcl: plain pointers and sizes, no C++
 * or torch types.  Every entry point cites the reference interface it replaces (paths relative to the rmcl repo).
 * The reference-side bindings a maintainer would add are shown in INTEGRATION.md; the C++ shim classes with the
 * reference's names live in include/rmcl_b200/ (C++ headers).
 *
 * Conventions
 *  - every function returns B2_OK (0) or a negative B2_ERR_*; b2_last_error() gives the thread-local message.
 *  - nothing here falls back to the CPU: if no CUDA device / kernel image is usable the call fails with B2_ERR_CUDA.
 *  - "device pointer" arguments are CUDA device addresses on the mesh's device; all work is enqueued on the handle's
 *    stream (b2_*_set_stream, default: the legacy default stream).  Calls that return data to HOST memory synchronise
 *    that stream; calls that take/return only device pointers do not.
 *  - layouts are byte-compatible with rmagine / rmcl.
 */
#ifndef RMCL_B200_H
#define RMCL_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define B2_API
#else
#define B2_API __attribute__((visibility("default")))
#endif

/* ---------------------------------------------------------------- POD types ---------------------------------- */
typedef struct { float x, y, z; } b2_vec3;                                   /* rmagine::Vector3f   (12 B) */
typedef struct { float x, y, z, w; } b2_quat;                                /* rmagine::Quaternion (16 B) */
typedef struct { b2_quat R; b2_vec3 t; uint32_t stamp; } b2_transform;       /* rmagine::Transform  (32 B) */
typedef struct { float m[9]; } b2_mat3;                                      /* rmagine::Matrix3x3, column-major */
typedef struct { b2_vec3 dataset_mean, model_mean; b2_mat3 covariance; uint32_t n_meas; } b2_cross_stats; /* rmagine::CrossStatistics (64 B) */
typedef struct { float mean, sigma; uint32_t n_meas; } b2_gaussian1d;        /* rmagine::Gaussian1D (12 B) */
typedef struct { b2_gaussian1d likelihood; float state_sigma[6]; } b2_particle_attr;  /* rmcl::ParticleAttributes, rmcl_ros/include/rmcl_ros/rmcl/ParticleAttributes.hpp:18-32 (36 B) */
typedef struct { b2_vec3 orig, dir; float range; b2_mat3 cov; } b2_range_meas;         /* rmcl::RangeMeasurement,  rmcl_ros/include/rmcl_ros/rmcl/RangeMeasurement.hpp:10-21 (64 B) */

/* rmagine::SphericalModel as filled at rmcl_ros/src/util/conversions.cpp:22-34 */
typedef struct { float phi_min, phi_inc; uint32_t phi_size; float theta_min, theta_inc; uint32_t theta_size; float range_min, range_max; } b2_spherical_model;
/* rmagine::PinholeModel as filled at rmcl_ros/src/util/conversions.cpp:48-60 */
typedef struct { uint32_t width, height; float fx, fy, cx, cy; float range_min, range_max; } b2_pinhole_model;

/* PCDSensorUpdaterEmbree config, rmcl_ros/src/rmcl/PCDSensorUpdaterEmbree.cpp:122-134 (defaults in comments) */
typedef struct {
    float dist_sigma;                 /* 2.0   */
    float real_hit_sim_miss_error;    /* 100.0 */
    float real_miss_sim_hit_error;    /* 100.0 */
    float real_miss_sim_miss_error;   /* 0.0   */
    float range_min, range_max;       /* sensor_range 0.05 .. 80.0 */
    int   ng_mode;                    /* 0: raw un-normalised geometric normal (Embree path, :57-68); 1: unit normal (OptiX path, BeamEvaluateProgram.cu:104-113) */
    int   correspondence_type;        /* 0: ray casting (evaluate_rcc, :18-86); 1: closest point (evaluate_cpc, :88-95; selected at :219-222) */
} b2_pf_params;

typedef struct {
    uint32_t n_faces, n_vertices, n_nodes, n_leaf_tris, max_depth;
    uint64_t bvh_bytes;               /* nodes + leaf triangle records resident in HBM */
    float    build_ms;                /* wall time of b2_mesh_create (host build + upload, or device build) */
    int      device;
    int      build_mode;
    float    sah_cost;
} b2_mesh_info;

typedef struct b2_mesh b2_mesh;       /* immutable map: triangle mesh + in-HBM BVH.  Stands in for rmagine::EmbreeMap / OptixMap */
typedef struct b2_rcc  b2_rcc;        /* one ray-casting-correspondence set == one rmcl::RCC..{Spherical,Pinhole,O1Dn,OnDn} object */
typedef struct b2_pf   b2_pf;         /* one particle-filter sensor updater == rmcl::PCDSensorUpdater{Embree,Optix} */

enum {
    B2_OK = 0,
    B2_ERR_INVALID = -1,      /* bad argument / call order (e.g. find before set_model) */
    B2_ERR_CUDA = -2,         /* CUDA runtime error (message in b2_last_error) */
    B2_ERR_NO_MAP = -3,       /* empty mesh: mirrors "NO MAP"/"EMPTY MAP" of PCDSensorUpdaterOptix.cpp:179-192 */
    B2_ERR_OOM = -4,
    B2_ERR_UNSUPPORTED = -5
};
enum { B2_BUILD_HOST_SAH = 0, B2_BUILD_DEVICE_LBVH = 1 };
/* correspondence type of a b2_rcc handle: ray casting (RCCEmbree*, default) or closest point (CPCEmbree) */
enum { B2_CORR_RCC = 0, B2_CORR_CPC = 1 };

B2_API const char* b2_last_error(void);
B2_API int         b2_version(void);
B2_API int         b2_device_count(int* n);

/* ---------------------------------------------------------------- map ---------------------------------------- */
/* replaces rm::import_embree_map + Embree scene commit (rmcl_ros/src/nodes/micp_localization.cpp:188,
 * rmcl_ros/src/rmcl/PCDSensorUpdaterEmbree.cpp:158): vertices/faces in HOST memory -> BVH resident in HBM of `device`. */
B2_API int b2_mesh_create(const float* verts_xyz, uint32_t n_vertices, const uint32_t* faces_ijk, uint32_t n_faces,
                          int device, int build_mode, b2_mesh** out);
/* rm::import_embree_map(file) (rmcl_ros/src/nodes/micp_localization.cpp:188, rmcl_ros/src/rmcl/PCDSensorUpdaterEmbree.cpp:158): mesh file ->
 * map.  Self-contained readers (no assimp) for Stanford PLY (ascii / binary_little_endian), Wavefront OBJ and COLLADA .dae (geometry library +
 * visual-scene node transforms; the format of the reference's example maps, docs/MICPL.md:46-49); polygons are fan-triangulated. */
B2_API int b2_mesh_create_from_file(const char* path, int device, int build_mode, b2_mesh** out);
/* the import step alone (host only, no device needed): malloc'ed vertex / face arrays, released with b2_mesh_file_free */
B2_API int b2_mesh_file_load(const char* path, float** verts_xyz, uint32_t* nv, uint32_t** faces_ijk, uint32_t* nf);
B2_API void b2_mesh_file_free(float* verts_xyz, uint32_t* faces_ijk);
/* Dynamic maps: the vertices moved, faces unchanged -> refit the resident BVH instead of rebuilding it (the reference
 * re-commits its Embree scene and flags dependants `outdated`, Correspondences.hpp:26-31).  Maps built with B2_BUILD_DEVICE_LBVH only; the call
 * synchronises the device first, so no handle may be tracing concurrently from another thread.  Results afterwards equal a fresh build's. */
B2_API int b2_mesh_refit(b2_mesh* m, const float* verts_xyz, uint32_t nv, int src_is_device);
/* BVH blob: the built map as one host buffer (header + nodes + leaf triangle records), so that a map is built once and
 * shipped to the other ranks (broadcast) or cached on disk; b2_mesh_create_from_blob validates the header and every index before uploading. */
B2_API int b2_mesh_blob_size(const b2_mesh* m, uint64_t* bytes);
B2_API int b2_mesh_export_blob(const b2_mesh* m, void* dst_host, uint64_t capacity);
B2_API int b2_mesh_create_from_blob(const void* blob_host, uint64_t bytes, int device, b2_mesh** out);
/* Drops the creator's reference.  b2_rcc / b2_pf handles created on the map hold their own references (the reference shares its map through
 * rm::EmbreeMapPtr, a shared_ptr, micp_localization.cpp:545), so map and handles may be destroyed in any order. */
B2_API int b2_mesh_destroy(b2_mesh* m);
B2_API int b2_mesh_get_info(const b2_mesh* m, b2_mesh_info* info);
/* closest hit for arbitrary rays (host arrays in, host arrays out): t in (0,tfar], tie -> smaller face id.
 * Replaces rtcIntersect1 as used at rmcl_ros/src/rmcl/PCDSensorUpdaterEmbree.cpp:30-47. Any output may be NULL. */
B2_API int b2_mesh_intersect(const b2_mesh* m, const float* origs_xyz, const float* dirs_xyz, uint32_t n, float tfar,
                             float* t_out, uint32_t* face_out, float* ng_out_xyz, uint8_t* hit_out);
/* traversal counters of the same rays (instrumented build of the same kernel): mean nodes visited / triangles tested per ray */
B2_API int b2_mesh_intersect_stats(const b2_mesh* m, const float* origs_xyz, const float* dirs_xyz, uint32_t n, float tfar,
                                   double* mean_nodes, double* mean_tris);

/* ---------------------------------------------------------------- MICP-L: RCC ------------------------------ */
/* ctor(map): rmcl/include/rmcl/registration/RCCEmbree.hpp:25-26 */
B2_API int b2_rcc_create(b2_mesh* map, b2_rcc** out);
B2_API int b2_rcc_destroy(b2_rcc* h);
B2_API int b2_rcc_set_stream(b2_rcc* h, void* cuda_stream);
/* Correspondences_::setTsb, rmcl/include/rmcl/registration/Correspondences.hpp:33-36 */
B2_API int b2_rcc_set_tsb(b2_rcc* h, const b2_transform* Tsb);
/* ModelSetter<ModelT>::setModel, rmcl/src/rmcl/registration/RCCEmbree.cpp:21-24,53-56,84-87,116-119 */
B2_API int b2_rcc_set_model_spherical(b2_rcc* h, const b2_spherical_model* model);
B2_API int b2_rcc_set_model_pinhole(b2_rcc* h, const b2_pinhole_model* model);
B2_API int b2_rcc_set_model_o1dn(b2_rcc* h, uint32_t width, uint32_t height, const float orig_xyz[3], const float* dirs_xyz, float range_min, float range_max);
B2_API int b2_rcc_set_model_ondn(b2_rcc* h, uint32_t width, uint32_t height, const float* origs_xyz, const float* dirs_xyz, float range_min, float range_max);
/* public fields params.max_dist / adaptive_max_dist_min, Correspondences.hpp:22-23 */
B2_API int b2_rcc_set_params(b2_rcc* h, float max_dist, float adaptive_max_dist_min);
/* public field dataset (points + mask), Correspondences.hpp:24; src_is_device != 0: pointers are device addresses */
B2_API int b2_rcc_set_dataset(b2_rcc* h, const float* points_xyz, const uint8_t* mask, uint32_t n, int src_is_device);
/* The same two members as CALLER-OWNED device memory, nothing copied: a subclass of rmcl::Correspondences_<rm::VRAM_CUDA> keeps the reference's public
 * `dataset` (rm::PointCloud_<VRAM_CUDA>, Correspondences.hpp:24) and protected `model_buffers_` (Correspondences.hpp:81-85) and binds their device
 * pointers here; find() then writes points / normals / hits straight into the bound buffers (capacity entries; grow and re-bind like
 * RCCOptix.cpp:36-40), the reductions read the bound dataset.  b2_rcc_set_dataset / set_ranges / correct_once_ranges switch back to the handle's
 * own dataset buffers; binding with capacity 0 unbinds the model buffers. */
B2_API int b2_rcc_bind_dataset(b2_rcc* h, const float* points_xyz_dev, const uint8_t* mask_dev, uint32_t n);
B2_API int b2_rcc_bind_model_buffers(b2_rcc* h, float* points_xyz_dev, float* normals_xyz_dev, uint8_t* hits_dev, uint32_t capacity);
/* MICP..Sensor..::unpackMessage (rmcl_ros/src/micpl/MICPSphericalSensorCPU.cpp:181-233) on the device: dataset = dir*range (+orig),
 * mask = range in [range.min, range.max].  Also the v1 setInputData(ranges) (lidar_corrector_embree_benchmark.cpp:118). */
B2_API int b2_rcc_set_ranges(b2_rcc* h, const float* ranges, uint32_t n, int src_is_device);
/* RCC..::find(Tbm_est), rmcl/src/rmcl/registration/RCCEmbree.cpp:26-36,58-68,89-99,121-131 */
B2_API int b2_rcc_find(b2_rcc* h, const b2_transform* Tbm_est);
/* The three semantics of rm::*SimulatorEmbree::simulate that SURVEY.md Appendix A.3 cannot settle without rmagine's source, as switches (all 0 = the
 * defaults stated there): tfar_mode 1: rays are traced to +inf instead of model.range.max; min_mode 1: a closest hit nearer than model.range.min is
 * reported as a miss; miss_fill 1: points / normals of a miss are zeros instead of NaN.  Affects b2_rcc_find, correct_once and correct_batch. */
B2_API int b2_rcc_set_sim_options(b2_rcc* h, int tfar_mode, int min_mode, int miss_fill);
/* CPCEmbree (rmcl/include/rmcl/registration/CPCEmbree.hpp:20-54, find at rmcl/src/rmcl/registration/CPCEmbree.cpp:17-43): with
 * B2_CORR_CPC, b2_rcc_find runs one closest-point query per DATASET point (mask not consulted, as in the reference) instead of tracing
 * the sensor model: Pm = Tsm*d_i; cp = map.closestPoint(Pm); hits = cp.d <= max_dist; points = Tms*cp.p; normals = Tms.R*cp.n.
 * No sensor model is needed; the model buffers get n_dataset entries (ranges = cp.d), so b2_rcc_cross_statistics and
 * b2_rcc_correct_once work unchanged.  b2_rcc_correct_once_ranges is refused in this mode. */
B2_API int b2_rcc_set_correspondence_type(b2_rcc* h, int type);
/* closest-point mode only.  skip_masked != 0: dataset points whose mask is 0 are not queried (their model entry reads hits = 0, NaN).  The reference
 * queries every point and never uses those entries (statistics_p2l tests dataset.mask); a dropped beam unpacked to range.max + 1 lies far outside the
 * map and its query is the most expensive of the scan.  Default 0 = the reference's literal behaviour. */
B2_API int b2_rcc_set_cpc_options(b2_rcc* h, int skip_masked);
/* Correspondences{CPU,CUDA}::computeCrossStatistics, rmcl/src/rmcl/registration/CorrespondencesCPU.cpp:10-39 (CUDA twin CorrespondencesCUDA.cpp:9-30) */
B2_API int b2_rcc_cross_statistics(b2_rcc* h, const b2_transform* T_snew_sold, double convergence_progress, b2_cross_stats* out_host);
/* Scan-vs-map segmentation: the classification of ScanMapSegmentationEmbreeNode::scanCB (rmcl_ros/src/nodes/filter/
 * scan_map_segmentation_embree.cpp:84-187).  Call after b2_rcc_set_ranges (the real scan) and b2_rcc_find(T_sensor_map with Tsb = identity,
 * or Tbm with the handle's Tsb): compares real and simulated ranges / normals per ray and returns the two outlier clouds in raster order
 * (HOST buffers of cap_* points x 3 floats; n_* receive the full counts even when the capacity is smaller) and optionally one label per
 * ray (0 none, 1 outlier_scan, 2 outlier_map). */
B2_API int b2_rcc_segment(b2_rcc* h, float min_dist_outlier_scan, float min_dist_outlier_map, float* outlier_scan_host, uint32_t cap_scan, uint32_t* n_scan,
                          float* outlier_map_host, uint32_t cap_map, uint32_t* n_map, uint8_t* labels_host);
/* modelView()/datasetView(), Correspondences.hpp:47-62: device pointers (points/normals packed xyz, hits/mask u8) + our extra face ids / ranges */
B2_API int b2_rcc_model_view(b2_rcc* h, float** points, float** normals, uint8_t** hits, uint32_t** face_ids, float** ranges, uint32_t* n);
B2_API int b2_rcc_dataset_view(b2_rcc* h, float** points, uint8_t** mask, uint32_t* n);
/* host copies of the model buffers (synchronises); any pointer may be NULL */
B2_API int b2_rcc_download_model(b2_rcc* h, float* points, float* normals, uint8_t* hits, uint32_t* face_ids, float* ranges);
B2_API int b2_rcc_download_dataset(b2_rcc* h, float* points, uint8_t* mask);

/* One MICPLocalizationNode::correctOnce for this sensor, entirely on the device
 * (rmcl_ros/src/nodes/micp_localization.cpp:899-984 + rmcl_ros/include/rmcl_ros/micpl/MICPSensor.hpp:146-184):
 * find(Tom*Tbo), then `iterations` x { P2L cross statistics -> frame changes -> Umeyama -> compose }.  Outputs to HOST (may be NULL). */
B2_API int b2_rcc_correct_once(b2_rcc* h, const b2_transform* Tom, const b2_transform* Tbo, uint32_t iterations,
                               double convergence_progress, b2_transform* Tom_new, b2_transform* T_onew_oold, b2_cross_stats* Cmerged_o);
/* same, but the scan arrives as HOST ranges and is uploaded inside the call (end-to-end entry point used by bench.py "e2e").  A pinned buffer
 * (cudaHostAlloc / cudaHostRegister) is moved by a copy engine while find runs and unpacked by the loop kernel; a pageable one is staged and
 * uploaded on a side stream.  The buffer must stay unchanged until the call returns. */
B2_API int b2_rcc_correct_once_ranges(b2_rcc* h, const float* ranges_host, uint32_t n, const b2_transform* Tom, const b2_transform* Tbo,
                                      uint32_t iterations, double convergence_progress, b2_transform* Tom_new, b2_transform* T_onew_oold,
                                      b2_cross_stats* Cmerged_o);

/* The same step split into enqueue / collect: _async launches the kernels on the handle's stream and returns at once, _wait blocks until the
 * result of the OLDEST call in flight has landed (mapped pinned memory, no stream synchronise) and hands it out.  Up to 8 calls may be in flight
 * per handle (exec mode 0: one); the synchronous entry points refuse to run while any is.  Lets a caller keep the GPU queue full (bench.py's
 * device-timed loop) or overlap its own host work with the correction.  The dataset must not be changed while calls are in flight. */
B2_API int b2_rcc_correct_once_async(b2_rcc* h, const b2_transform* Tom, const b2_transform* Tbo, uint32_t iterations, double convergence_progress);
B2_API int b2_rcc_correct_once_wait(b2_rcc* h, b2_transform* Tom_new, b2_transform* T_onew_oold, b2_cross_stats* Cmerged_o);
/* MICPLocalizationNode::correctOnce over ALL sensors of the node (rmcl_ros/src/nodes/micp_localization.cpp:899-984): per sensor k
 * find(Tom * Tbo[k]) (:900-908), then `iterations` x { per sensor: P2L cross statistics under T_bnew_bold = ~Tbo[k] * T_onew_oold * Tbo[k] (:926-929),
 * Cs_o = Tbo[k] * Cs_b (:931), weighted copy n_meas *= merge_weights[k] (u32 *= double, :933-934), Cmerged_o += Cs_o, Cmerged_weighted_o +=
 * Cs_weighted_o (:936-937) } -> umeyama(Cmerged_weighted_o) -> compose (:952-963).  1..4 sensors (RCC or CPC handles on one device); all inner
 * iterations of all sensors run in one kernel.  merge_weights NULL = all 1.0 (MICPSensor.hpp:103); ranges_host NULL or per-sensor NULL = use the
 * resident dataset, else that sensor's scan is uploaded / read zero-copy inside the call.  Outputs to HOST (may be NULL). */
B2_API int b2_micp_correct_once(b2_rcc* const* sensors, const b2_transform* Tbo, const double* merge_weights, const float* const* ranges_host, uint32_t n_sensors,
                                const b2_transform* Tom, uint32_t iterations, double convergence_progress,
                                b2_transform* Tom_new, b2_transform* T_onew_oold, b2_cross_stats* Cmerged_o);
/* How b2_rcc_correct_once* runs the inner iterations: 2 (default) one kernel for all of them, launched programmatically behind find; the block
 * sums cross the grid as 64-bit fixed-point atomics (every block must be co-resident: one block per SM).  1 the same kernel through a cooperative
 * launch, the sums as FP64 slots behind a grid sync.  0 one reduction launch per inner iteration in the reference's own frame-algebra order
 * (MICPSensor.hpp:178-182).  Results agree within float rounding.  The environment variable B2_FUSED sets the default of new handles.  A mode-2
 * call whose blocks cannot become co-resident, or whose sums leave the fixed-point range (|block partial| >= 2^46), is run again through mode 1
 * by the library itself. */
B2_API int b2_rcc_set_exec_mode(b2_rcc* h, int mode);

/* v1 {Sphere,Pinhole,O1Dn}Corrector{Embree,Optix}::correct(Tbm[N]) -> {Tdelta[N], Ncorr[N]}
 * (shape: rmcl_ros/src/benchmarks/lidar_corrector_embree_benchmark.cpp:86-133, ..._optix_benchmark.cpp:85-155):
 * fused trace -> P2L gate -> per-pose cross statistics -> batched Umeyama, one launch sequence for all poses.
 * poses_on_device / out_on_device select HOST or DEVICE pointers. Outputs may be NULL. */
B2_API int b2_rcc_correct_batch(b2_rcc* h, const b2_transform* Tbm, uint32_t n_poses, int poses_on_device,
                                b2_transform* Tdelta, uint32_t* ncorr, b2_cross_stats* stats_b, int out_on_device);
/* v1 corrector.benchmark(Tbm, Nruns) -> {sim, red, svd} seconds (rmcl_ros/src/benchmarks/lidar_corrector_optix_benchmark.cpp:143-155): the stages of
 * correct() run unfused (trace of all poses -> P2L reduction -> Umeyama) and timed separately with CUDA events, summed over n_runs.  Poses on the HOST. */
B2_API int b2_rcc_benchmark_batch(b2_rcc* h, const b2_transform* Tbm_host, uint32_t n_poses, uint32_t n_runs, double* sim_s, double* red_s, double* svd_s);
/* rm::umeyama_transform for n statistics (rmcl_ros/src/nodes/micp_localization.cpp:952-953) */
B2_API int b2_umeyama_batch(const b2_cross_stats* stats, uint32_t n, b2_transform* out, int on_device, int device, void* cuda_stream);

/* ---------------------------------------------------------------- particle filter ---------------------------- */
B2_API int b2_pf_create(b2_mesh* map, b2_pf** out);
B2_API int b2_pf_destroy(b2_pf* h);
B2_API int b2_pf_set_stream(b2_pf* h, void* cuda_stream);
/* How b2_pf_sensor_update maps rays to lanes (a schedule: results are bit-identical).  0: the 32 lanes of a warp trace 32 beams of ONE particle --
 * the reference's loop order seen from a particle (PCDSensorUpdaterEmbree.cpp:290-342) and the better choice for particle sets spread over the map.
 * 1: the lanes trace the SAME beam for 32 particles; 2: the same after sorting the particles by (heading, map cell) on the device, coherent and up to
 * 1.7x faster once the cloud has converged.  3 (default): 0 or 2, whichever was faster when last timed (both on the first two updates, the slower one
 * again every 64 updates).  The environment variable B2_PF_MAP sets the default of new handles. */
B2_API int b2_pf_set_mapping(b2_pf* h, int mode);
B2_API int b2_pf_get_mapping(b2_pf* h, int* mode, int* current);
/* ParticleUpdater<VRAM_CUDA>::update (rmcl_ros/include/rmcl_ros/rmcl/ParticleUpdater.hpp:39-43) with the hot loop of
 * PCDSensorUpdaterEmbree::update (rmcl_ros/src/rmcl/PCDSensorUpdaterEmbree.cpp:290-342): all beams x all particles in ONE launch,
 * per-particle likelihood merged in beam order, attrs read-modified-written once.  poses/attrs: DEVICE pointers; beams: HOST. */
B2_API int b2_pf_sensor_update(b2_pf* h, const b2_transform* poses_dev, b2_particle_attr* attrs_dev, uint32_t n_particles,
                               const b2_transform* Tsb, const b2_range_meas* beams_host, uint32_t n_beams, const b2_pf_params* params);
/* ParticleUpdater<RAM>::update: HOST poses/attrs; copies in, updates, copies attrs back (end-to-end entry point).  Sets of 32 768 particles and more
 * are processed in 8 chunks alternating between two streams: with pinned host arrays the transfers run under the kernel (C3: 5.2 ms for 5.13 ms of
 * kernel).  attrs_host is updated in place. */
B2_API int b2_pf_sensor_update_host(b2_pf* h, const b2_transform* poses_host, b2_particle_attr* attrs_host, uint32_t n_particles,
                                    const b2_transform* Tsb, const b2_range_meas* beams_host, uint32_t n_beams, const b2_pf_params* params);

/* rest of the particle-filter cycle on the device, so that particles never leave HBM between stages:
 * TFMotionUpdaterGPU / particle_move_and_forget (rmcl_ros/src/rmcl/particle_motion.cu:11-46): pose = pose * T_bnew_bold, n_meas -= forget_rate * n_meas.
 * check_collision != 0 adds the wall check of the CPU updater (TFMotionUpdaterCPU.cpp:17-50,205-216): a ray from the old to the new position; a hit
 * sets the particle's likelihood to {0, 0, MAX_N_MEAS}. */
B2_API int b2_pf_motion_update(b2_pf* h, b2_transform* poses_dev, b2_particle_attr* attrs_dev, uint32_t n_particles, const b2_transform* T_bnew_bold, double forget_rate,
                               int check_collision);
/* compute_stats (rmcl_ros/src/rmcl/resampling.cu:41-92): sum and max (initial 0) of likelihood.mean over the LOCAL particles; with particles
 * sharded across GPUs the caller all-reduces the 8 bytes (SUM, MAX) -- the one exchange step of the cycle.  Results to HOST. */
B2_API int b2_pf_likelihood_stats(b2_pf* h, const b2_particle_attr* attrs_dev, uint32_t n_particles, float* sum_out, float* max_out);

/* GladiatorResamplerGPU / gladiator_resample (rmcl_ros/src/rmcl/resampling.cu:108-221; config GladiatorResamplerConfig.hpp:7-20): each
 * champion i draws a random opponent; if the opponent's likelihood.mean is larger, the champion's slot receives a perturbed copy of the
 * opponent (Gaussian noise on t and on the Euler angles) whose n_meas is reduced by the forget rate; else it keeps its own state.
 * poses/attrs hold ALL n_all particles (on one GPU: the local ones; sharded: the all-gathered set -- the one exchange step of this stage);
 * this call produces the champions first .. first+n_local-1 into poses_new/attrs_new (n_local entries; must not alias the inputs).
 * Draws come from Philox4x32-10 keyed by (seed, step, global particle index) unless raw_dev/normals_dev (n_local, n_local x 6) supply them. */
typedef struct {
    float min_noise_tx, min_noise_ty, min_noise_tz, min_noise_roll, min_noise_pitch, min_noise_yaw;
    float likelihood_forget_per_meter, likelihood_forget_per_radian;     /* reference defaults 0.3, 0.2 */
} b2_gladiator_config;
B2_API int b2_pf_resample_gladiator(b2_pf* h, const b2_transform* poses_dev, const b2_particle_attr* attrs_dev, uint32_t n_all, uint32_t first, uint32_t n_local,
                                    b2_transform* poses_new_dev, b2_particle_attr* attrs_new_dev, const b2_gladiator_config* cfg, uint64_t seed, uint32_t step,
                                    const uint32_t* raw_dev, const float* normals_dev);
/* The same resampling with the particles SHARDED over GPUs (one process per GPU) and no all-gather: every rank publishes its shard in a buffer
 * the other ranks map over NVLink (CUDA IPC); a champion reads its opponent's 4-byte likelihood from the owner's HBM and fetches the 68-byte
 * record only when the opponent wins (the reference draws opponents from all particles, resampling.cu:137).  Equal shard sizes; results equal the
 * single-GPU resampling of the concatenated set bit for bit.  init -> (exchange the 128-byte handles of all ranks) -> connect; per step:
 * publish, cross-rank barrier, resample_p2p, cross-rank barrier.  remote_bytes_out (may be NULL) receives the bytes read from other GPUs. */
B2_API int b2_pf_p2p_init(b2_pf* h, uint32_t n_per_rank, void* handles_out_128_bytes);
B2_API int b2_pf_p2p_connect(b2_pf* h, const void* all_handles_world_x_128, uint32_t world, uint32_t rank, uint32_t n_per_rank);
B2_API int b2_pf_p2p_publish(b2_pf* h, const b2_transform* poses_dev, const b2_particle_attr* attrs_dev, uint32_t n_local);
B2_API int b2_pf_resample_gladiator_p2p(b2_pf* h, b2_transform* poses_new_dev, b2_particle_attr* attrs_new_dev, const b2_gladiator_config* cfg, uint64_t seed, uint32_t step,
                                        uint64_t* remote_bytes_out);
/* test hook: several shards on one device (poses_all / attrs_all hold world * n_per_rank particles; this handle plays `rank`) */
B2_API int b2_pf_p2p_connect_local(b2_pf* h, const b2_transform* poses_all_dev, const b2_particle_attr* attrs_all_dev, uint32_t world, uint32_t rank, uint32_t n_per_rank);
/* the draws b2_pf_resample_gladiator would use (replaces init_curand / curand(), resampling.cu:13-30,134-143): raw u32 + 6 normals per particle */
B2_API int b2_pf_gladiator_randoms(b2_pf* h, uint64_t seed, uint32_t step, uint32_t first, uint32_t n, uint32_t* raw_dev, float* normals_dev);

/* ---------------------------------------------------------------- introspection ------------------------------ */
/* CUDA-event timing of the kernels inside b2_rcc_correct_once*(): when enabled, events are recorded on the handle's stream around the
 * find kernel and around the reduction/Umeyama kernels; b2_rcc_last_timing returns the two durations of the most recent call (ms). */
B2_API int b2_rcc_enable_timing(b2_rcc* h, int enable);
B2_API int b2_rcc_last_timing(b2_rcc* h, float* find_ms, float* reduce_ms);
/* the calling thread's pending CUDA runtime error as text ("" if none), without clearing it: no entry point of this library leaves one behind */
B2_API const char* b2_peek_cuda_error(void);
/* memory-system micro-benchmark for bench.py's roofline denominators: read bandwidth (GB/s) of a `bytes` working set, 128-bit loads from all SMs,
 * `iters` timed launches after a warm-up.  Below the L2 capacity it measures the L2, far above it the HBM. */
B2_API int b2_debug_read_bandwidth(int device, uint64_t bytes, int iters, double* gbytes_per_s);
/* number of kernels this library has launched in this process (bench.py's gpu_launches) */
B2_API uint64_t b2_kernel_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* RMCL_B200_H */
