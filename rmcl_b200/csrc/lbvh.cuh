// lbvh.cuh -- device builder of the map (B2_BUILD_DEVICE_LBVH): everything from the triangle soup to the 8-wide float-box BVH of bvh8.h
// runs in CUDA kernels on the B200; the host only reads back two counters per tree level.
//
//   k_lbvh_prims      per triangle: AABB, centroid; scene bounds by ordered-int atomics
//   k_lbvh_morton     30-bit Morton code of the centroid
//   cub::DeviceRadixSort::SortPairs (CUDA toolkit)  codes + triangle ids
//   k_lbvh_hierarchy  binary radix tree over the sorted codes (T. Karras, "Maximizing Parallelism in the Construction of BVHs, Octrees
//                     and k-d Trees", HPG 2012: every internal node finds its key range and split with two binary searches)
//   k_lbvh_refit      bottom-up boxes, second arrival at a parent continues (atomic flags)
//   k_lbvh_collapse   one launch per level of the wide tree: a thread owns one 8-wide node, pulls up the largest-area grandchildren
//                     until it has 8 children (sub-trees with <= 3 triangles become leaf children), assigns octant slots, allocates its
//                     inner children / leaf records with two atomics and writes node + records in the bvh8.h layout.
// It replaces the Embree/OptiX scene commit (rm::import_embree_map, rmcl_ros/src/nodes/micp_localization.cpp:188) and is the DEFAULT build
// (B2_BUILD_DEVICE_LBVH): 7 ms per million triangles, refittable, and on the measured scans its shallower tree (depth 8) traces as fast as the
// host SAH builder's (B2_BUILD_HOST_SAH, 1 s per million triangles, kept as an option).  Results are identical either way (the hit definition
// does not depend on the tree, trace.cuh).
#pragma once
#include <cub/device/device_radix_sort.cuh>

#include <vector>
#include "bvh8.h"

struct LbvhBox { float lo[3], hi[3]; };

__device__ __forceinline__ unsigned int f2ord(float f) { const unsigned int u = __float_as_uint(f); return (u & 0x80000000u) ? ~u : (u | 0x80000000u); }
__device__ __forceinline__ float ord2f(unsigned int o) { return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o); }

// bounds[0..2] = ordered min, bounds[3..5] = ordered max (of centroids), bounds[6..8] = ordered max |coordinate|
__global__ void k_lbvh_prims(const float* __restrict__ verts, const uint32_t* __restrict__ faces, uint32_t nf, LbvhBox* __restrict__ tbox, float* __restrict__ cent,
                             unsigned int* __restrict__ bounds)
{
    const uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
    float c[3] = {0.f, 0.f, 0.f}, am[3] = {0.f, 0.f, 0.f};
    const bool in = f < nf;
    if (in) {
        LbvhBox b;
        const float* v0 = verts + 3 * (size_t)faces[3 * (size_t)f], *v1 = verts + 3 * (size_t)faces[3 * (size_t)f + 1], *v2 = verts + 3 * (size_t)faces[3 * (size_t)f + 2];
        for (int k = 0; k < 3; k++) {
            b.lo[k] = fminf(fminf(v0[k], v1[k]), v2[k]); b.hi[k] = fmaxf(fmaxf(v0[k], v1[k]), v2[k]);
            c[k] = 0.5f * (b.lo[k] + b.hi[k]); am[k] = fmaxf(fabsf(b.lo[k]), fabsf(b.hi[k]));
            cent[3 * (size_t)f + k] = c[k];
        }
        tbox[f] = b;
    }
    for (int k = 0; k < 3; k++) {
        unsigned int lo = in ? f2ord(c[k]) : 0xffffffffu, hi = in ? f2ord(c[k]) : 0u, a = in ? f2ord(am[k]) : 0u;
        lo = __reduce_min_sync(0xffffffffu, lo); hi = __reduce_max_sync(0xffffffffu, hi); a = __reduce_max_sync(0xffffffffu, a);
        if ((threadIdx.x & 31) == 0) { atomicMin(bounds + k, lo); atomicMax(bounds + 3 + k, hi); atomicMax(bounds + 6 + k, a); }
    }
}

__device__ __forceinline__ uint32_t expand10(uint32_t x) { x &= 0x3ffu; x = (x | (x << 16)) & 0x30000ffu; x = (x | (x << 8)) & 0x300f00fu; x = (x | (x << 4)) & 0x30c30c3u; x = (x | (x << 2)) & 0x9249249u; return x; }

__global__ void k_lbvh_morton(const float* __restrict__ cent, uint32_t nf, const unsigned int* __restrict__ bounds, uint32_t* __restrict__ codes, uint32_t* __restrict__ ids)
{
    const uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nf) return;
    uint32_t q[3];
    for (int k = 0; k < 3; k++) {
        const float lo = ord2f(bounds[k]), hi = ord2f(bounds[3 + k]);
        const float ext = hi - lo;
        const float u = ext > 0.f ? (cent[3 * (size_t)f + k] - lo) / ext : 0.f;
        q[k] = (uint32_t)fminf(fmaxf(u * 1024.0f, 0.0f), 1023.0f);
    }
    codes[f] = (expand10(q[0]) << 2) | (expand10(q[1]) << 1) | expand10(q[2]);
    ids[f] = f;
}

// length of the common prefix of keys i and j (ties broken by position), -1 outside the array
__device__ __forceinline__ int lbvh_delta(const uint32_t* __restrict__ codes, int n, int i, int j)
{
    if (j < 0 || j >= n) return -1;
    const uint32_t a = codes[i], b = codes[j];
    return a != b ? __clz(a ^ b) : 32 + __clz((uint32_t)i ^ (uint32_t)j);
}

// node ids: internal 0..n-2, leaf k (sorted position) = (n-1) + k.  parent[] covers all 2n-1 ids.
__global__ void k_lbvh_hierarchy(const uint32_t* __restrict__ codes, int n, uint32_t* __restrict__ left, uint32_t* __restrict__ right, uint32_t* __restrict__ parent,
                                 uint32_t* __restrict__ first, uint32_t* __restrict__ last)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    const int d = (lbvh_delta(codes, n, i, i + 1) - lbvh_delta(codes, n, i, i - 1)) >= 0 ? 1 : -1;
    const int dmin = lbvh_delta(codes, n, i, i - d);
    int lmax = 2;
    while (lbvh_delta(codes, n, i, i + lmax * d) > dmin) lmax *= 2;
    int l = 0;
    for (int t = lmax / 2; t >= 1; t /= 2) if (lbvh_delta(codes, n, i, i + (l + t) * d) > dmin) l += t;
    const int j = i + l * d;
    const int dnode = lbvh_delta(codes, n, i, j);
    int s = 0;
    for (int t = (l + 1) / 2;; t = (t + 1) / 2) {
        if (lbvh_delta(codes, n, i, i + (s + t) * d) > dnode) s += t;
        if (t <= 1) break;
    }
    const int g = i + s * d + min(d, 0);
    const int lo = min(i, j), hi = max(i, j);
    const uint32_t L = (lo == g) ? (uint32_t)(n - 1 + g) : (uint32_t)g;
    const uint32_t R = (hi == g + 1) ? (uint32_t)(n - 1 + g + 1) : (uint32_t)(g + 1);
    left[i] = L; right[i] = R; parent[L] = (uint32_t)i; parent[R] = (uint32_t)i;
    first[i] = (uint32_t)lo; last[i] = (uint32_t)hi;
    if (i == 0) parent[0] = 0xffffffffu;
}

__global__ void k_lbvh_refit(const LbvhBox* __restrict__ tbox, const uint32_t* __restrict__ ids, int n, const uint32_t* __restrict__ left, const uint32_t* __restrict__ right,
                             const uint32_t* __restrict__ parent, LbvhBox* __restrict__ nbox, unsigned int* __restrict__ flags)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const uint32_t leaf = (uint32_t)(n - 1 + k);
    nbox[leaf] = tbox[ids[k]];
    __threadfence();
    uint32_t p = parent[leaf];
    while (p != 0xffffffffu) {
        if (atomicAdd(flags + p, 1u) == 0u) return;          // first child to arrive: the sibling will finish this parent
        __threadfence();
        const LbvhBox a = nbox[left[p]], b = nbox[right[p]];
        LbvhBox u;
        for (int c = 0; c < 3; c++) { u.lo[c] = fminf(a.lo[c], b.lo[c]); u.hi[c] = fmaxf(a.hi[c], b.hi[c]); }
        nbox[p] = u;
        __threadfence();
        p = parent[p];
    }
}

__device__ __forceinline__ float lbvh_area(const LbvhBox& b) { const float dx = b.hi[0] - b.lo[0], dy = b.hi[1] - b.lo[1], dz = b.hi[2] - b.lo[2]; return dx * dy + dy * dz + dz * dx; }
__device__ __forceinline__ uint32_t lbvh_count(uint32_t id, int n, const uint32_t* first, const uint32_t* last) { return id >= (uint32_t)(n - 1) ? 1u : last[id] - first[id] + 1u; }

// one thread per wide node of the current level
__global__ void k_lbvh_collapse(uint32_t level_begin, uint32_t level_end, uint32_t* __restrict__ root_of, int n, const uint32_t* __restrict__ left, const uint32_t* __restrict__ right,
                                const uint32_t* __restrict__ first, const uint32_t* __restrict__ last, const LbvhBox* __restrict__ nbox, const uint32_t* __restrict__ ids,
                                const float* __restrict__ verts, const uint32_t* __restrict__ faces, B2Node8* __restrict__ nodes8, B2Tri* __restrict__ tris8,
                                uint32_t* __restrict__ counters /* [0] nodes, [1] tris */)
{
    const uint32_t t = level_begin + blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= level_end) return;
    const uint32_t root = root_of[t];
    uint32_t c[8]; int cnt;
    if (root >= (uint32_t)(n - 1)) { c[0] = root; cnt = 1; }
    else { c[0] = left[root]; c[1] = right[root]; cnt = 2; }
    // pull up grandchildren: first the sub-trees that cannot be leaves (> 3 triangles), largest area first, then (slots left) any inner
    for (int pass = 0; pass < 2; pass++) {
        while (cnt < 8) {
            int best = -1; float ba = -1.f;
            for (int k = 0; k < cnt; k++) {
                if (c[k] >= (uint32_t)(n - 1)) continue;
                const uint32_t m = lbvh_count(c[k], n, first, last);
                if (pass == 0 && m <= B2_MAX_LEAF_TRIS) continue;
                const float a = lbvh_area(nbox[c[k]]);
                if (a > ba) { ba = a; best = k; }
            }
            if (best < 0) break;
            const uint32_t b = c[best];
            c[best] = left[b]; c[cnt++] = right[b];
        }
    }
    // octant slot assignment (greedy on (centroid_c - centroid_node) . D_s), as in the host builder
    const LbvhBox nb = nbox[root];
    float cen[3]; for (int k = 0; k < 3; k++) cen[k] = 0.5f * (nb.lo[k] + nb.hi[k]);
    float dvec[8][3];
    for (int k = 0; k < cnt; k++) { const LbvhBox b = nbox[c[k]]; for (int a = 0; a < 3; a++) dvec[k][a] = 0.5f * (b.lo[a] + b.hi[a]) - cen[a]; }
    int child_in_slot[8]; for (int s = 0; s < 8; s++) child_in_slot[s] = -1;
    uint32_t assigned = 0;
    for (int it = 0; it < cnt; it++) {
        float bv = -3.0e38f; int bc = -1, bs = -1;
        for (int k = 0; k < cnt; k++) {
            if (assigned & (1u << k)) continue;
            for (int s = 0; s < 8; s++) {
                if (child_in_slot[s] >= 0) continue;
                const float v = ((s & 1) ? dvec[k][0] : -dvec[k][0]) + ((s & 2) ? dvec[k][1] : -dvec[k][1]) + ((s & 4) ? dvec[k][2] : -dvec[k][2]);
                if (v > bv) { bv = v; bc = k; bs = s; }
            }
        }
        child_in_slot[bs] = bc; assigned |= 1u << bc;
    }
    uint32_t n_inner = 0, n_leaf_tris = 0;
    for (int k = 0; k < cnt; k++) { const uint32_t m = lbvh_count(c[k], n, first, last); if (m > B2_MAX_LEAF_TRIS) n_inner++; else n_leaf_tris += m; }
    const uint32_t child_base = n_inner ? atomicAdd(counters + 0, n_inner) : 0u;
    const uint32_t tri_base = n_leaf_tris ? atomicAdd(counters + 1, n_leaf_tris) : 0u;
    B2Node8 nd;
    for (int a = 0; a < 3; a++) for (int s = 0; s < 8; s++) { nd.lo[a][s] = __int_as_float(0x7f800000); nd.hi[a][s] = __int_as_float(0xff800000); }
    for (int s = 0; s < 8; s++) nd.meta[s] = 0;
    nd.child_base = child_base; nd.tri_base = tri_base; nd.masks = 0; nd.pad0 = 0; nd.pad[0] = nd.pad[1] = 0;
    uint32_t ki = 0, toff = 0;
    for (int s = 0; s < 8; s++) {
        const int k = child_in_slot[s];
        if (k < 0) continue;
        const uint32_t id = c[k];
        const LbvhBox b = nbox[id];
        for (int a = 0; a < 3; a++) { nd.lo[a][s] = b.lo[a]; nd.hi[a][s] = b.hi[a]; }
        const uint32_t m = lbvh_count(id, n, first, last);
        if (m > B2_MAX_LEAF_TRIS) {
            nd.meta[s] = (uint8_t)(0x20 | (24 + s));
            root_of[child_base + ki] = id; ki++;
        } else {
            const uint32_t unary = m == 1 ? 1u : (m == 2 ? 3u : 7u);
            nd.meta[s] = (uint8_t)((unary << 5) | toff);
            const uint32_t p0 = id >= (uint32_t)(n - 1) ? id - (uint32_t)(n - 1) : first[id];
            for (uint32_t q = 0; q < m; q++) {
                const uint32_t f = ids[p0 + q];
                B2Tri tr;
                const float* v0 = verts + 3 * (size_t)faces[3 * (size_t)f], *v1 = verts + 3 * (size_t)faces[3 * (size_t)f + 1], *v2 = verts + 3 * (size_t)faces[3 * (size_t)f + 2];
                for (int a = 0; a < 3; a++) { tr.v0[a] = v0[a]; tr.v1[a] = v1[a]; tr.v2[a] = v2[a]; }
                tr.face_id = f; tr.pad1 = 0; tr.pad2 = 0;
                tris8[tri_base + toff + q] = tr;
            }
            toff += m;
        }
    }
    nd.masks = b2_masks_from_meta(nd.meta);
    nodes8[t] = nd;
}

// host driver.  verts/faces are DEVICE pointers.  On success *nodes_out / *tris_out are exact-size device buffers owned by the caller.
static int lbvh_build_device(const float* d_verts, uint32_t nv, const uint32_t* d_faces, uint32_t nf, B2Node8** nodes_out, uint32_t* n_nodes_out, B2Tri** tris_out,
                             uint32_t* n_tris_out, uint32_t* depth_out, float abs_max_out[3], const char** err, std::vector<uint32_t>* level_begin = nullptr)
{
    (void)nv;
    static const char* e_cuda = "CUDA error in the device BVH build";
    static const char* e_depth = "BVH too deep for the traversal stack";
    *err = e_cuda;
    const int n = (int)nf;
    LbvhBox *tbox = nullptr, *nbox = nullptr; float* cent = nullptr; unsigned int *bounds = nullptr, *flags = nullptr;
    uint32_t *codes = nullptr, *ids = nullptr, *codes_s = nullptr, *ids_s = nullptr, *left = nullptr, *right = nullptr, *parent = nullptr, *first = nullptr, *last = nullptr;
    uint32_t *root_of = nullptr, *counters = nullptr; void* cub_tmp = nullptr; B2Node8* nodes8 = nullptr; B2Tri* tris8 = nullptr;
    int rc = -2;
    auto freeall = [&]() {
        cudaFree(tbox); cudaFree(nbox); cudaFree(cent); cudaFree(bounds); cudaFree(flags); cudaFree(codes); cudaFree(ids); cudaFree(codes_s); cudaFree(ids_s);
        cudaFree(left); cudaFree(right); cudaFree(parent); cudaFree(first); cudaFree(last); cudaFree(root_of); cudaFree(counters); cudaFree(cub_tmp);
    };
    static thread_local char e_detail[256];
#define LB(call) do { const cudaError_t e_ = (call); if (e_ != cudaSuccess) { snprintf(e_detail, sizeof(e_detail), "CUDA error in the device BVH build: %s at lbvh.cuh:%d", cudaGetErrorString(e_), __LINE__); *err = e_detail; (void)cudaGetLastError(); freeall(); cudaFree(nodes8); cudaFree(tris8); return rc; } } while (0)
    const size_t N = (size_t)nf;
    LB(cudaMalloc(&tbox, sizeof(LbvhBox) * N)); LB(cudaMalloc(&nbox, sizeof(LbvhBox) * (2 * N)));
    LB(cudaMalloc(&cent, sizeof(float) * 3 * N)); LB(cudaMalloc(&bounds, sizeof(unsigned int) * 9)); LB(cudaMalloc(&flags, sizeof(unsigned int) * N));
    LB(cudaMalloc(&codes, 4 * N)); LB(cudaMalloc(&ids, 4 * N)); LB(cudaMalloc(&codes_s, 4 * N)); LB(cudaMalloc(&ids_s, 4 * N));
    LB(cudaMalloc(&left, 4 * N)); LB(cudaMalloc(&right, 4 * N)); LB(cudaMalloc(&parent, 4 * 2 * N)); LB(cudaMalloc(&first, 4 * N)); LB(cudaMalloc(&last, 4 * N));
    LB(cudaMalloc(&root_of, 4 * N)); LB(cudaMalloc(&counters, 8));
    LB(cudaMalloc(&nodes8, sizeof(B2Node8) * N)); LB(cudaMalloc(&tris8, sizeof(B2Tri) * N));
    const unsigned int binit[9] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0u, 0u, 0u, 0u, 0u, 0u};
    LB(cudaMemcpy(bounds, binit, sizeof(binit), cudaMemcpyHostToDevice));
    LB(cudaMemset(flags, 0, sizeof(unsigned int) * N));
    const uint32_t gb = (nf + 255) / 256;
    k_lbvh_prims<<<gb, 256>>>(d_verts, d_faces, nf, tbox, cent, bounds);
    k_lbvh_morton<<<gb, 256>>>(cent, nf, bounds, codes, ids);
    size_t tmp_bytes = 0;
    LB(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, codes, codes_s, ids, ids_s, n, 0, 30));
    LB(cudaMalloc(&cub_tmp, tmp_bytes ? tmp_bytes : 16));
    LB(cub::DeviceRadixSort::SortPairs(cub_tmp, tmp_bytes, codes, codes_s, ids, ids_s, n, 0, 30));
    if (n > 1) k_lbvh_hierarchy<<<(nf + 255) / 256, 256>>>(codes_s, n, left, right, parent, first, last);
    else { const uint32_t none = 0xffffffffu; LB(cudaMemcpy(parent, &none, 4, cudaMemcpyHostToDevice)); }
    k_lbvh_refit<<<gb, 256>>>(tbox, ids_s, n, left, right, parent, nbox, flags);
    // level-by-level collapse
    const uint32_t root_id = n > 1 ? 0u : 0u;                       // internal node 0, or leaf id (n-1)+0 = 0 when n == 1
    uint32_t h_counters[2] = {1u, 0u};
    LB(cudaMemcpy(root_of, &root_id, 4, cudaMemcpyHostToDevice));
    LB(cudaMemcpy(counters, h_counters, 8, cudaMemcpyHostToDevice));
    uint32_t begin = 0, end = 1, depth = 0;
    while (begin < end) {
        if (level_begin) level_begin->push_back(begin);       // nodes of one level are contiguous: [level_begin[l], level_begin[l+1])
        depth++;
        if (depth > B2_TRAVERSAL_STACK - 4) { *err = e_depth; rc = -5; freeall(); cudaFree(nodes8); cudaFree(tris8); return rc; }
        k_lbvh_collapse<<<(end - begin + 127) / 128, 128>>>(begin, end, root_of, n, left, right, first, last, nbox, ids_s, d_verts, d_faces, nodes8, tris8, counters);
        LB(cudaMemcpy(h_counters, counters, 8, cudaMemcpyDeviceToHost));
        begin = end; end = h_counters[0];
    }
    if (level_begin) level_begin->push_back(end);
    LB(cudaGetLastError());
    unsigned int hb[9];
    LB(cudaMemcpy(hb, bounds, sizeof(hb), cudaMemcpyDeviceToHost));
    for (int k = 0; k < 3; k++) { const unsigned int o = hb[6 + k]; unsigned int u = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o; memcpy(&abs_max_out[k], &u, 4); }
    // shrink to exact size
    B2Node8* nodes_exact = nullptr; B2Tri* tris_exact = nullptr;
    const uint32_t nn = h_counters[0], nt = h_counters[1];
    if (cudaMalloc(&nodes_exact, sizeof(B2Node8) * (size_t)nn) != cudaSuccess || cudaMalloc(&tris_exact, sizeof(B2Tri) * (size_t)(nt ? nt : 1)) != cudaSuccess ||
        cudaMemcpy(nodes_exact, nodes8, sizeof(B2Node8) * (size_t)nn, cudaMemcpyDeviceToDevice) != cudaSuccess ||
        cudaMemcpy(tris_exact, tris8, sizeof(B2Tri) * (size_t)nt, cudaMemcpyDeviceToDevice) != cudaSuccess) {
        cudaFree(nodes_exact); cudaFree(tris_exact); freeall(); cudaFree(nodes8); cudaFree(tris8); return rc;
    }
    freeall(); cudaFree(nodes8); cudaFree(tris8);
#undef LB
    *nodes_out = nodes_exact; *n_nodes_out = nn; *tris_out = tris_exact; *n_tris_out = nt; *depth_out = depth;
    return 0;
}


// ---------------------------------------------------------------------------------------------------------------------
// Refit for dynamic maps: the vertices moved, the topology (faces, tree) stays.  Bottom-up over the levels of the wide
// tree: a leaf child's box is the exact float AABB of its re-fetched triangles, an inner child's box the union of that child node's own
// child boxes (already refitted: deeper level).  Boxes stay exact AABBs of the triangles below, so the hit definition -- and with it
// every result -- is that of a freshly built map; only the tree's quality degrades as the mesh deforms.
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_bvh8_refit_level(uint32_t begin, uint32_t end, B2Node8* __restrict__ nodes, B2Tri* __restrict__ tris, const float* __restrict__ verts,
                                                          const uint32_t* __restrict__ faces)
{
    const uint32_t t = begin + blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= end) return;
    bvh8_refit_node(t, nodes, tris, verts, faces);
}
// max |coordinate| per axis (slack constant of the box test); bits[k] must be zeroed by the caller
__global__ void k_abs_max(const float* __restrict__ verts, uint32_t nv, unsigned int* __restrict__ bits)
{
    float m[3] = {0.f, 0.f, 0.f};
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += gridDim.x * blockDim.x)
        for (int k = 0; k < 3; k++) m[k] = fmaxf(m[k], fabsf(verts[3 * (size_t)i + k]));
    for (int k = 0; k < 3; k++) {
        for (int o = 16; o > 0; o >>= 1) m[k] = fmaxf(m[k], __shfl_xor_sync(0xffffffffu, m[k], o));
        if ((threadIdx.x & 31) == 0) atomicMax(bits + k, __float_as_uint(m[k]));          // non-negative floats order like their bit patterns
    }
}
