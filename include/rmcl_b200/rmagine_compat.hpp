// rmagine_compat.hpp -- the handful of rmagine types that cross the boundary of the ray-casting-correspondence path, layout-compatible
// with rmagine 2.4, for builds WITHOUT rmagine.  With rmagine available, define RMCL_B200_WITH_RMAGINE and the
// real headers are used instead (the shim classes in rcc_b200.hpp only rely on the members named here).
#pragma once
#ifdef RMCL_B200_WITH_RMAGINE
#include <rmagine/math/types.h>
#include <rmagine/types/Memory.hpp>
#include <rmagine/types/PointCloud.hpp>
#include <rmagine/types/sensor_models.h>
#include <rmagine/types/UmeyamaReductionConstraints.hpp>
#else
#include <cstddef>
#include <cstdint>
#include <memory>
#include <vector>

namespace rmagine {

struct Vector3f { float x, y, z; };
using Vector = Vector3f;
struct Quaternion { float x, y, z, w; };
struct Transform { Quaternion R; Vector3f t; uint32_t stamp;
    static Transform Identity() { return Transform{{0.f, 0.f, 0.f, 1.f}, {0.f, 0.f, 0.f}, 0u}; } };
struct Matrix3x3 { float data[9]; float& operator()(int r, int c) { return data[c * 3 + r]; } float operator()(int r, int c) const { return data[c * 3 + r]; }
    float trace() const { return data[0] + data[4] + data[8]; } };
struct CrossStatistics { Vector3f dataset_mean, model_mean; Matrix3x3 covariance; uint32_t n_meas; };
struct Gaussian1D { float mean, sigma; uint32_t n_meas; static Gaussian1D Identity() { return Gaussian1D{0.f, 0.f, 0u}; } };
struct Interval { float min, max; bool inside(float v) const { return min <= v && v <= max; } };
struct DiscreteInterval { float min, inc; uint32_t size; float operator[](uint32_t i) const { return min + static_cast<float>(i) * inc; } };
struct UmeyamaReductionConstraints { float max_dist; };
static_assert(sizeof(Transform) == 32 && sizeof(CrossStatistics) == 64 && sizeof(Gaussian1D) == 12, "rmagine layouts");

struct SphericalModel { DiscreteInterval phi, theta; Interval range;
    uint32_t getWidth() const { return theta.size; } uint32_t getHeight() const { return phi.size; } size_t size() const { return size_t(phi.size) * theta.size; } };
struct PinholeModel { uint32_t width, height; Interval range; float f[2], c[2];
    uint32_t getWidth() const { return width; } uint32_t getHeight() const { return height; } size_t size() const { return size_t(width) * height; } };
struct O1DnModel { uint32_t width, height; Interval range; Vector3f orig; std::vector<Vector3f> dirs;
    uint32_t getWidth() const { return width; } uint32_t getHeight() const { return height; } size_t size() const { return size_t(width) * height; } };
struct OnDnModel { uint32_t width, height; Interval range; std::vector<Vector3f> origs, dirs;
    uint32_t getWidth() const { return width; } uint32_t getHeight() const { return height; } size_t size() const { return size_t(width) * height; } };

struct RAM {};
struct VRAM_CUDA {};
// non-owning view (rmagine::MemoryView): for VRAM_CUDA the pointer is a device address
template <typename T, typename MemT = RAM> struct MemoryView { T* ptr = nullptr; size_t n = 0; T* raw() const { return ptr; } size_t size() const { return n; }
    MemoryView operator()(size_t b, size_t e) const { return MemoryView{ptr + b, e - b}; } };
template <typename ModelT> class ModelSetter { public: virtual ~ModelSetter() = default; virtual void setModel(const ModelT&) = 0; };

}  // namespace rmagine
#endif
