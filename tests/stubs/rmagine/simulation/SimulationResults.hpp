// STUB of rmagine/simulation/SimulationResults.hpp (tests/stubs/README.md): the attribute structs of a simulation bundle
#pragma once
#include <cstdint>

#include "../math/types.h"
#include "../types/Bundle.hpp"
#include "../types/Memory.hpp"

namespace rmagine {

template <typename MemT> struct Hits { Memory<uint8_t, MemT> hits; };
template <typename MemT> struct Ranges { Memory<float, MemT> ranges; };
template <typename MemT> struct Points { Memory<Point, MemT> points; };
template <typename MemT> struct Normals { Memory<Vector, MemT> normals; };
template <typename MemT> struct FaceIds { Memory<unsigned int, MemT> face_ids; };

// grow every attribute of a bundle to W x H x N entries
template <typename MemT, typename BundleT> void resize_memory_bundle(BundleT& res, unsigned int W, unsigned int H, unsigned int N)
{
    const size_t n = size_t(W) * H * N;
    if constexpr (BundleT::template has<Hits<MemT>>()) res.Hits<MemT>::hits.resize(n);
    if constexpr (BundleT::template has<Ranges<MemT>>()) res.Ranges<MemT>::ranges.resize(n);
    if constexpr (BundleT::template has<Points<MemT>>()) res.Points<MemT>::points.resize(n);
    if constexpr (BundleT::template has<Normals<MemT>>()) res.Normals<MemT>::normals.resize(n);
    if constexpr (BundleT::template has<FaceIds<MemT>>()) res.FaceIds<MemT>::face_ids.resize(n);
}

}  // namespace rmagine
