// bvh8.h -- in-HBM layout of the map: 8-wide BVH with 8-bit quantised child boxes (80-byte nodes) + 48-byte leaf
// triangle records.  Shared by the host builder, the device builder and the traversal kernels.
//
// The layout idea (8 children, per-node quantisation frame, octant-ordered child slots, triangle ranges per node) is the
// published "compressed wide BVH" of Ylitie, Karras, Laine, HPG 2017; this is an independent implementation.
//
// Node (80 B = 5 x 16 B, loaded as 5 x LDG.128):
//   q0: float px, py, pz;  u8 ex, ey, ez (IEEE exponent bytes: scale_k = 2^(e_k-127)), u8 imask (bit s: slot s is an inner node)
//   q1: u32 child_base (index of first inner child), u32 tri_base (index of first leaf triangle record),
//       u8 meta[8]  (slot s: 0 = empty; inner: 0x20 | (24 + s); leaf: (unary tri count 1|3|7) << 5 | offset of first tri from tri_base)
//   q2: u8 qlo_x[8], u8 qlo_y[8]
//   q3: u8 qlo_z[8], u8 qhi_x[8]
//   q4: u8 qhi_y[8], u8 qhi_z[8]
//   child box (real numbers): lo_k = p_k + qlo_k[s] * scale_k, hi_k = p_k + qhi_k[s] * scale_k  -- always CONTAINS the float AABB of
//   every triangle below it (the builder checks this in double precision).
//   Slot s "points" along D_s = (s&1 ? + : -, s&2 ? + : -, s&4 ? + : -); a ray with octant code r (bit k set iff d_k >= 0) visits
//   inner children in order of descending (s ^ r).
//
// Leaf triangle record (48 B = 3 x 16 B): (v0.xyz, face_id as bits), (v1.xyz, 0), (v2.xyz, 0); records of one node are contiguous,
// in slot order; at most 3 triangles per leaf child, at most 24 per node.
#pragma once
#include <stdint.h>

struct alignas(16) B2Node8 {
    float    p[3];
    uint8_t  e[3];
    uint8_t  imask;
    uint32_t child_base;
    uint32_t tri_base;
    uint8_t  meta[8];
    uint8_t  qlo[3][8];
    uint8_t  qhi[3][8];
};
static_assert(sizeof(B2Node8) == 80, "node must be 80 bytes");

struct alignas(16) B2Tri {
    float    v0[3]; uint32_t face_id;
    float    v1[3]; uint32_t pad1;
    float    v2[3]; uint32_t pad2;
};
static_assert(sizeof(B2Tri) == 48, "triangle record must be 48 bytes");

#define B2_TRAVERSAL_STACK 40          // uint2 entries per ray; builder refuses trees deeper than B2_TRAVERSAL_STACK - 4
#define B2_MAX_LEAF_TRIS 3

// host-side result of a build (either builder)
struct B2BvhHost {
    B2Node8* nodes = nullptr; uint32_t n_nodes = 0;
    B2Tri*   tris = nullptr;  uint32_t n_tris = 0;
    uint32_t max_depth = 0;
    float    sah_cost = 0.f;
    float    scene_lo[3] = {0, 0, 0}, scene_hi[3] = {0, 0, 0};
};

// host SAH builder (bvh_build.cpp). Returns 0 on success, negative on failure (message via *err).
int b2_build_bvh8_host(const float* verts, uint32_t nv, const uint32_t* faces, uint32_t nf, B2BvhHost* out, const char** err);
void b2_free_bvh8_host(B2BvhHost* b);
