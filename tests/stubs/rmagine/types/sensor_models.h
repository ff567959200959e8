// STUB of rmagine/types/sensor_models.h (tests/stubs/README.md): fields as filled at rmcl_ros/src/util/conversions.cpp:22-120, ModelSetter mix-in
#pragma once
#include <cmath>
#include <cstdint>

#include "../math/types.h"
#include "Memory.hpp"

namespace rmagine {

struct SphericalModel { DiscreteInterval phi, theta; Interval range;
    uint32_t getWidth() const { return theta.size; } uint32_t getHeight() const { return phi.size; } size_t size() const { return size_t(phi.size) * theta.size; }
    uint32_t getBufferId(uint32_t vid, uint32_t hid) const { return vid * theta.size + hid; }
    Vector getDirection(uint32_t vid, uint32_t hid) const { const float p = phi[vid], t = theta[hid]; return {std::cos(p) * std::cos(t), std::cos(p) * std::sin(t), std::sin(p)}; } };
struct PinholeModel { uint32_t width, height; Interval range; float f[2], c[2];
    uint32_t getWidth() const { return width; } uint32_t getHeight() const { return height; } size_t size() const { return size_t(width) * height; }
    uint32_t getBufferId(uint32_t vid, uint32_t hid) const { return vid * width + hid; } };
struct O1DnModel { uint32_t width, height; Interval range; Vector orig; Memory<Vector, RAM> dirs;
    uint32_t getWidth() const { return width; } uint32_t getHeight() const { return height; } size_t size() const { return size_t(width) * height; } };
struct OnDnModel { uint32_t width, height; Interval range; Memory<Vector, RAM> origs, dirs;
    uint32_t getWidth() const { return width; } uint32_t getHeight() const { return height; } size_t size() const { return size_t(width) * height; } };

template <typename ModelT> class ModelSetter { public: virtual ~ModelSetter() = default; virtual void setModel(const ModelT&) = 0; };

}  // namespace rmagine
