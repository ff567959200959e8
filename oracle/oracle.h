/*
 * oracle.h -- CPU restatement of RMCL's ray-casting-correspondence hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under rmcl_b200/ (the product) may include,
 * link or call this.  Allowed users: tests/, __graft_entry__.smoke(), and the
 * cpu_baseline / --impl reference legs of bench.py.
 *
 * PARITY UNPINNED: the arithmetic of this path lives in rmagine (+Embree), which
 * is not vendored under /root/reference (source_dependencies.yaml:4-7, unpinned
 * `main`; rmcl/CMakeLists.txt:62-73 asks for rmagine >= 2.4.0) and cannot be built
 * here; the reference ships no tests or golden vectors.
 * Every function below cites the in-tree reference call site it follows; rmagine
 * semantics are restated from its published behaviour and
 * pinned only against analytic cases (tests/test_oracle_*.py).
 */
#ifndef RMCL_ORACLE_H
#define RMCL_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- POD layouts at the boundary ---- */
typedef struct { float x, y, z; } orc_vec3;                       /* rm::Vector3f, 12 B   */
typedef struct { float x, y, z, w; } orc_quat;                    /* rm::Quaternion, 16 B */
typedef struct { orc_quat R; orc_vec3 t; uint32_t stamp; } orc_transform; /* rm::Transform 32 B */
typedef struct { float m[9]; } orc_mat3;                          /* rm::Matrix3x3, column-major m[c*3+r] */
typedef struct {
    orc_vec3 dataset_mean; orc_vec3 model_mean; orc_mat3 covariance; uint32_t n_meas;
} orc_cross_stats;                                                /* rm::CrossStatistics 64 B */
typedef struct { float mean, sigma; uint32_t n_meas; } orc_gaussian1d;  /* rm::Gaussian1D 12 B */
/* rmcl_ros/include/rmcl_ros/rmcl/ParticleAttributes.hpp:18-32 */
typedef struct { orc_gaussian1d likelihood; float state_sigma[6]; } orc_particle_attr; /* 36 B */
/* rmcl_ros/include/rmcl_ros/rmcl/RangeMeasurement.hpp:10-21 */
typedef struct { orc_vec3 orig, dir; float range; orc_mat3 cov; } orc_range_meas;       /* 64 B */

/* rmcl_ros/src/rmcl/PCDSensorUpdaterEmbree.cpp:122-134 (defaults there) */
typedef struct {
    float dist_sigma;                 /* 2.0   */
    float real_hit_sim_miss_error;    /* 100.0 */
    float real_miss_sim_hit_error;    /* 100.0 */
    float real_miss_sim_miss_error;   /* 0.0   */
    float range_min, range_max;       /* 0.05, 80.0 */
    int   ng_mode;                    /* 0 = raw (un-normalised) Ng like the Embree path (quirk D1), 1 = unit normal like the OptiX path */
    int   correspondence_type;        /* 0 = ray casting (evaluate_rcc, :18-86), 1 = closest point (evaluate_cpc, :88-95, selected at :219-222) */
} orc_pf_params;

typedef struct orc_scene orc_scene;

/* ---- scene (stands in for rm::EmbreeMap: closest-hit over a triangle mesh) ---- */
orc_scene* orc_scene_create(const float* verts_xyz, uint32_t nv, const uint32_t* faces_ijk, uint32_t nf);
void       orc_scene_destroy(orc_scene* s);
uint32_t   orc_scene_num_faces(const orc_scene* s);

/* closest hit, t in (0, tfar]; tie -> smaller face id.  brute != 0: test every triangle (no BVH).
 * Restates rtcIntersect1 as used at PCDSensorUpdaterEmbree.cpp:30-47. ng = raw geometric normal (v1-v0)x(v2-v0). */
int  orc_intersect(const orc_scene* s, const float o[3], const float d[3], float tfar, int brute,
                   float* t_out, uint32_t* face_out, float ng_out[3]);
void orc_intersect_batch(const orc_scene* s, uint32_t n, const float* origs, const float* dirs, float tfar, int brute,
                         float* t_out, uint32_t* face_out, float* ng_out, uint8_t* hit_out);

/* ---- sensor models (rmagine SphericalModel / PinholeModel::getDirection; witness rmcl_ros/src/util/conversions.cpp:174-188) ---- */
void orc_spherical_dirs(float phi_min, float phi_inc, uint32_t phi_n, float theta_min, float theta_inc, uint32_t theta_n, float* dirs_out);
void orc_pinhole_dirs(uint32_t width, uint32_t height, float fx, float fy, float cx, float cy, float* dirs_out);

/* ---- math ---- */
void orc_transform_mul(const orc_transform* a, const orc_transform* b, orc_transform* out);
void orc_transform_inv(const orc_transform* a, orc_transform* out);
void orc_transform_point(const orc_transform* T, const float p[3], float out[3]);
void orc_quat_rotate(const orc_quat* q, const float v[3], float out[3]);

/* ---- RCCEmbree*::find == X SimulatorEmbree::simulate (rmcl/src/rmcl/registration/RCCEmbree.cpp:26-36,58-68,89-99,121-131) ----
 * rays given as sensor-frame tables: dirs[n], origs[n_origs] with n_origs in {1 (shared origin: spherical/pinhole/O1Dn), n (OnDn)}.
 * Outputs in the SENSOR frame; misses: hits=0, points/normals NaN, face id 0xFFFFFFFF, range = range_max + 1. Any out pointer may be NULL. */
void orc_simulate(const orc_scene* s, const orc_transform* Tbm, const orc_transform* Tsb,
                  uint32_t n, const float* origs_s, uint32_t n_origs, const float* dirs_s, float range_max,
                  float* points, float* normals, uint8_t* hits, uint32_t* face_ids, float* ranges);

/* The three rmagine semantics SURVEY.md Appendix A.3 leaves open (rmagine's source is not in the tree), as switches; 0 = the stated default. */
typedef struct {
    int tfar_mode;     /* 0: ray.tfar = model.range.max (hits beyond it are misses); 1: tfar = +inf (every hit is reported, like PCDSensorUpdaterEmbree.cpp:38) */
    int min_mode;      /* 0: a closest hit with t < model.range.min is a hit; 1: it is a miss (sim_hit needs t > range.min at PCDSensorUpdaterEmbree.cpp:47) */
    int miss_fill;     /* 0: points / normals of a miss are NaN; 1: zeros */
} orc_sim_options;
void orc_simulate_opt(const orc_scene* s, const orc_transform* Tbm, const orc_transform* Tsb,
                      uint32_t n, const float* origs_s, uint32_t n_origs, const float* dirs_s, float range_min, float range_max, const orc_sim_options* opt,
                      float* points, float* normals, uint8_t* hits, uint32_t* face_ids, float* ranges);

/* MICP*Sensor*::unpackMessage (rmcl_ros/src/micpl/MICPSphericalSensorCPU.cpp:181-233): dataset point = dir*range (+orig), mask = range in [min,max] */
void orc_dataset_from_ranges(uint32_t n, const float* origs_s, uint32_t n_origs, const float* dirs_s, const float* ranges,
                             float range_min, float range_max, float* points, uint8_t* mask, uint32_t* n_valid);

/* ---- rm::statistics_p2l as called at rmcl/src/rmcl/registration/CorrespondencesCPU.cpp:26-30 (formula witness rmcl_ros/src/micpl/MICPSensorCPU.cpp:71-98) ----
 * orc_statistics_p2l:     FP32, sequential `stats += CrossStatistics(Di,Mi)` merges (restates the reference arithmetic, one thread).
 * orc_statistics_p2l_f64: identical FP32 per-element math and gating (=> identical n_meas), accumulation in double sum form (precision reference). */
void orc_statistics_p2l(const orc_transform* Tpre, uint32_t n, const float* dpts, const uint8_t* dmask,
                        const float* mpts, const float* mnrm, const uint8_t* mmask, float max_dist, orc_cross_stats* out);
void orc_statistics_p2l_par(const orc_transform* Tpre, uint32_t n, const float* dpts, const uint8_t* dmask,
                            const float* mpts, const float* mnrm, const uint8_t* mmask, float max_dist, orc_cross_stats* out);   /* OpenMP chunks, FP32: timing baseline */
void orc_statistics_p2l_f64(const orc_transform* Tpre, uint32_t n, const float* dpts, const uint8_t* dmask,
                            const float* mpts, const float* mnrm, const uint8_t* mmask, float max_dist, orc_cross_stats* out);
/* max_dist interpolation of CorrespondencesCPU.cpp:21-23 */
float orc_adaptive_max_dist(float max_dist, float adaptive_max_dist_min, double convergence_progress);

void orc_cross_stats_identity(orc_cross_stats* s);
void orc_cross_stats_merge(const orc_cross_stats* a, const orc_cross_stats* b, orc_cross_stats* out);            /* rm::CrossStatistics::operator+= */
void orc_cross_stats_transform(const orc_transform* T, const orc_cross_stats* s, orc_cross_stats* out);            /* Transform * CrossStatistics */

/* rm::umeyama_transform(CrossStatistics) as called at rmcl_ros/src/nodes/micp_localization.cpp:952-953 */
void orc_umeyama(const orc_cross_stats* s, orc_transform* out);

/* One MICPLocalizationNode::correctOnce for ONE sensor (micp_localization.cpp:899-984 + MICPSensor.hpp:146-184).
 * Returns Tom_new; optional outputs: T_onew_oold, last merged stats. */
void orc_micp_correct_once(const orc_scene* s,
                           uint32_t n, const float* origs_s, uint32_t n_origs, const float* dirs_s, float range_max,
                           const float* dataset_pts, const uint8_t* dataset_mask,
                           const orc_transform* Tom, const orc_transform* Tbo, const orc_transform* Tsb,
                           uint32_t optimization_iterations, float max_dist, float adaptive_max_dist_min,
                           double convergence_progress, int f64_accum /* 0: FP32 sequential, 1: FP64 sums, 2: FP32 OpenMP chunks */,
                           orc_transform* Tom_new, orc_transform* T_onew_oold, orc_cross_stats* Cmerged_o);

/* The same over ALL sensors of the node (micp_localization.cpp:899-984): per sensor find, then per inner iteration the per-sensor statistics are
 * moved to the odom frame, merged un-weighted (Cmerged_o, :936) and weighted (n_meas *= merge_weight as u32 *= double, :933-937); Umeyama runs on the
 * weighted merge.  dirs_s == NULL selects closest-point correspondences for that sensor. */
typedef struct {
    uint32_t n; uint32_t n_origs;
    const float* origs_s; const float* dirs_s; const float* dataset_pts; const uint8_t* dataset_mask;
    orc_transform Tbo, Tsb;
    float range_max, max_dist, adaptive_max_dist_min, pad_;
    double merge_weight;
} orc_micp_sensor;
void orc_micp_correct_once_multi(const orc_scene* s, uint32_t n_sensors, const orc_micp_sensor* sensors, const orc_transform* Tom,
                                 uint32_t optimization_iterations, double convergence_progress, int f64_accum,
                                 orc_transform* Tom_new, orc_transform* T_onew_oold, orc_cross_stats* Cmerged_o);

/* v1 SphereCorrectorEmbree::correct(Tbm[N]) shape (rmcl_ros/src/benchmarks/lidar_corrector_embree_benchmark.cpp:117-133):
 * per pose: simulate at Tbm[p], P2L against the dataset built from `ranges`, one Umeyama step; Tdelta is in the BASE frame. */
void orc_correct_batch(const orc_scene* s, uint32_t n_poses, const orc_transform* Tbm, const orc_transform* Tsb,
                       uint32_t n, const float* origs_s, uint32_t n_origs, const float* dirs_s, float range_min, float range_max,
                       const float* ranges, float max_dist, int f64_accum,
                       orc_transform* Tdelta, uint32_t* ncorr, orc_cross_stats* stats_b);

/* ---- particle filter: PCDSensorUpdaterEmbree::update hot loop (PCDSensorUpdaterEmbree.cpp:290-342) with beams as input ---- */
float orc_pf_evaluate_rcc(const orc_scene* s, const orc_range_meas* meas_m, const orc_pf_params* p);               /* :18-86 */
void  orc_pf_sensor_update_one(const orc_scene* s, const orc_transform* Tsm, const orc_range_meas* meas_s,
                               const orc_pf_params* p, orc_particle_attr* attr_inout);                              /* :197-241 */
void  orc_pf_update(const orc_scene* s, uint32_t n_particles, const orc_transform* poses, orc_particle_attr* attrs,
                    const orc_transform* Tsb, uint32_t n_beams, const orc_range_meas* beams_s, const orc_pf_params* p);
void  orc_gaussian1d_add(orc_gaussian1d* a, const orc_gaussian1d* b);                                               /* rm::Gaussian1D::operator+= */

/* closest-point correspondences: rm::EmbreeMap::closestPoint + CPCEmbree::find (rmcl/src/rmcl/registration/CPCEmbree.cpp:17-43) */
int   orc_closest_point(const orc_scene* s, const float q[3], int brute, float* d_out, float p_out[3], float n_out[3], uint32_t* face_out);
void  orc_cpc_find(const orc_scene* s, const orc_transform* Tbm, const orc_transform* Tsb, uint32_t n, const float* dataset_pts, float max_dist, int brute,
                   float* points, float* normals, uint8_t* hits, uint32_t* face_ids, float* dists);

/* rest of the PF cycle: motion update (particle_motion.cu:11-46) and likelihood statistics (resampling.cu:41-92) */
void  orc_pf_motion_update(uint32_t n, orc_transform* poses, orc_particle_attr* attrs, const orc_transform* T_bnew_bold, double forget_rate);
void  orc_pf_motion_update_collide(const orc_scene* s, uint32_t n, orc_transform* poses, orc_particle_attr* attrs, const orc_transform* T_bnew_bold, double forget_rate);   /* + wall check, TFMotionUpdaterCPU.cpp:17-50,184-224 */
void  orc_pf_likelihood_stats(uint32_t n, const orc_particle_attr* attrs, float* sum_out, float* max_out);

/* Gladiator resampling, device variant (rmcl_ros/src/rmcl/resampling.cu:108-199; config GladiatorResamplerConfig.hpp:7-20).
 * The reference draws from cuRAND XORWOW (curand_init(1234, idx, 0), :20); its skip-ahead tables are not in the tree, and its own CPU
 * variant uses std::mt19937 (GladiatorResamplerCPU.cpp:96-109), so the random stream is an implementation detail: here the draws are
 * Philox4x32-10 (Salmon et al., SC'11) with counter (global particle index, 0, step, block) and key = seed, one raw u32 (opponent) and
 * six Box-Muller normals per particle.  The resampling itself is a pure function of (particles, draws). */
typedef struct {
    float min_noise_tx, min_noise_ty, min_noise_tz, min_noise_roll, min_noise_pitch, min_noise_yaw;
    float likelihood_forget_per_meter, likelihood_forget_per_radian;     /* defaults 0.3, 0.2 */
} orc_gladiator_config;
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
void orc_pf_gladiator_randoms(uint64_t seed, uint32_t step, uint32_t first, uint32_t n, uint32_t* raw_out, float* normals_out /* n x 6 */);
/* champions first .. first+n_local-1 of the n_all particles; opponent = raw % n_all; outputs have n_local entries */
void orc_pf_gladiator_resample(uint32_t n_all, const orc_transform* poses, const orc_particle_attr* attrs, uint32_t first, uint32_t n_local,
                               const uint32_t* raw, const float* normals, const orc_gladiator_config* cfg,
                               orc_transform* poses_new, orc_particle_attr* attrs_new);

/* scan-vs-map segmentation: the classification loop of ScanMapSegmentationEmbreeNode (rmcl_ros/src/nodes/filter/
 * scan_map_segmentation_embree.cpp:110-187) given the simulated ranges + normals (sensor frame) and the real ranges.  Outputs the two
 * outlier clouds in raster order (n x 3 capacity each) and, optionally, a per-ray label (0 none, 1 scan outlier, 2 map outlier). */
void orc_segment(uint32_t n, const float* origs_s, uint32_t n_origs, const float* dirs_s, float range_min, float range_max,
                 const float* ranges_real, const float* ranges_sim, const float* normals_sim, float min_dist_outlier_scan, float min_dist_outlier_map,
                 float* outlier_scan, uint32_t* n_scan, float* outlier_map, uint32_t* n_map, uint8_t* labels);

int orc_num_threads(void);
void orc_set_num_threads(int n);

#ifdef __cplusplus
}
#endif
#endif
