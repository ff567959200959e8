// STUB of rmagine/types/PointCloud.hpp (tests/stubs/README.md)
#pragma once
#include <cstdint>

#include "../math/types.h"
#include "Memory.hpp"

namespace rmagine {

template <typename MemT> struct PointCloud_ { Memory<Vector, MemT> points; Memory<uint8_t, MemT> mask; Memory<Vector, MemT> normals; Memory<unsigned int, MemT> ids; };
template <typename MemT> struct PointCloudView_ {
    MemoryView<Vector, MemT> points;
    MemoryView<uint8_t, MemT> mask = MemoryView<uint8_t, MemT>::Empty();
    MemoryView<Vector, MemT> normals = MemoryView<Vector, MemT>::Empty();
    MemoryView<unsigned int, MemT> ids = MemoryView<unsigned int, MemT>::Empty();
};
template <typename MemT> PointCloudView_<MemT> watch(PointCloud_<MemT>& c) { return PointCloudView_<MemT>{c.points, c.mask, c.normals, c.ids}; }
template <typename MemT> const PointCloudView_<MemT> watch(const PointCloud_<MemT>& c)
{
    auto& m = const_cast<PointCloud_<MemT>&>(c);
    return PointCloudView_<MemT>{m.points, m.mask, m.normals, m.ids};
}
using PointCloud = PointCloud_<RAM>;

}  // namespace rmagine
