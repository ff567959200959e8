// trace.cuh -- software closest-hit traversal of the 8-wide quantised BVH (bvh8.h) for sm_100a.
//
// Stands in for Embree's rtcIntersect1 behind rm::*SimulatorEmbree::simulate (called at
// rmcl/src/rmcl/registration/RCCEmbree.cpp:35,67,98,130) and for the direct call at
// rmcl_ros/src/rmcl/PCDSensorUpdaterEmbree.cpp:44.  B200 has no RT cores: this is the whole ray tracer.
//
// HIT DEFINITION (identical to oracle/oracle.c, independent of the acceleration structure):
//   slab formula F for a box: idir_k = 1/(|d_k| < 1e-18 ? copysign(1e-18,d_k) : d_k); t = fl(fl(plane - o) * idir)
//   candidate t (Moeller-Trumbore, scaled form, explicit FMA chains) is a HIT iff t > 0 and, for the triangle's own AABB,
//       tn <= fl(t*C1)  and  tf >= fl(t*C2)                       C1 = 1+2^-14, C2 = 1-2^-14
//   result = argmin (t, face id) over HITs with t <= tfar.
// CULLING: a child box (exact float AABB of the triangles below it) is skipped only if NOT
//       max(tn'_x, tn'_y, tn'_z, 0) <= min(tf'_x, tf'_y, tf'_z, fl(tbest*C1))
//   with  tf'_k = fma(far_plane_k,  idir_k,      -cf_k),  cf_k = oi_k - dl_k                 oi_k = fl(o_k * idir_k)
//         tn'_k = fma(near_plane_k, idir_k*C3',  -cn_k),  cn_k = (oi_k + dl_k) * C3'         C3' = 1 - 2^-11
//         dl_k  = 2^-20 * (|oi_k| + B_k * |idir_k|)        B_k = max |vertex coordinate| on axis k
//   All constants are per RAY (RaySetup); a node visit is 48 FMAs + min/max, no per-node setup.  dl_k bounds the FMA-path error
//   (<= 2^-23 (|plane*idir| + |oi|)) plus F's own error, so tn' <= C3' * tn_F and tf' >= tf_F for the same float box; with the C3' scale
//   this test is at least as permissive as the oracle's visit rule V (tn <= tbest*C1 && tf >= max(0, tn*C3), C3 = 1-2^-12): if tn_F > 0,
//   tn' <= tn_F*C3' <= fl(tn_F*C3) <= tf_F <= tf'; if tn_F <= 0, tn' <= 0 <= tf_F.  Hence every HIT's leaf is visited and the result
//   equals the brute-force answer.
#pragma once
#include "b2_math.cuh"
#include "bvh8.h"

#define B2_C1 1.00006103515625f
#define B2_C2 0.99993896484375f
#define B2_C3 0.999755859375f
#define B2_NOFACE 0xFFFFFFFFu

struct BvhView {
    const float4* nodes;    // B2_NODE_QUADS (14) x 16 B per node
    const float4* tris;     // 3 x float4 per triangle record
    float bx, by, bz;       // max |vertex coordinate| per axis
};

struct HitRec {
    float    t;             // closest t (== tfar on a miss)
    uint32_t face;          // original face id, B2_NOFACE on a miss
    uint32_t tri;           // leaf-record index of the hit triangle
};

struct RaySetup {
    V3 o, d, idir;
    V3 idn, cn, cf;         // near-plane slope (idir*C3'), near / far constants of the box test (see header)
    uint32_t oct;           // bit k set iff idir_k >= 0
    uint32_t onx, ony, onz; // float4 offsets of the near planes inside a node (lo arrays for positive directions, hi arrays otherwise)
};

#define B2_C3P 0.99951171875f          // 1 - 2^-11

B2_DEV RaySetup ray_setup(V3 o, V3 d, const struct BvhView& bvh);

// formula F on an exact float box
B2_DEV void slab_F(const RaySetup& r, V3 lo, V3 hi, float& tn, float& tf)
{
    const float x0 = mul(sub(lo.x, r.o.x), r.idir.x), x1 = mul(sub(hi.x, r.o.x), r.idir.x);
    const float y0 = mul(sub(lo.y, r.o.y), r.idir.y), y1 = mul(sub(hi.y, r.o.y), r.idir.y);
    const float z0 = mul(sub(lo.z, r.o.z), r.idir.z), z1 = mul(sub(hi.z, r.o.z), r.idir.z);
    tn = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fminf(z0, z1));
    tf = fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fmaxf(z0, z1));
}

// Moeller-Trumbore + validation; updates (tbest, face, tri) under the (t, face) lexicographic rule
B2_DEV void tri_test(const BvhView& bvh, const RaySetup& r, uint32_t tri_idx, HitRec& best)
{
    const float4 a = ldg(bvh.tris + 3 * (size_t)tri_idx + 0);
    const float4 b = ldg(bvh.tris + 3 * (size_t)tri_idx + 1);
    const float4 c = ldg(bvh.tris + 3 * (size_t)tri_idx + 2);
    const V3 v0 = mk3(a.x, a.y, a.z), v1 = mk3(b.x, b.y, b.z), v2 = mk3(c.x, c.y, c.z);
    const V3 e1 = v_sub(v1, v0), e2 = v_sub(v2, v0);
    const V3 pv = cross_fma(r.d, e2);
    const float det = dot_fma(e1, pv);
    if (!(det != 0.0f)) return;                 // zero or NaN
    const V3 tv = v_sub(r.o, v0);
    const float U = dot_fma(tv, pv);
    const V3 qv = cross_fma(tv, e1);
    const float V = dot_fma(r.d, qv);
    const float T = dot_fma(e2, qv);
    const float UV = add(U, V);
    bool ok;
    if (det > 0.0f) ok = (U >= 0.0f) && (V >= 0.0f) && (UV <= det) && (T > 0.0f);
    else            ok = (U <= 0.0f) && (V <= 0.0f) && (UV >= det) && (T < 0.0f);
    if (!ok) return;
    const float t = dvd(T, det);
    const uint32_t face = f2u(a.w);
    if (!(t < best.t || (t == best.t && face < best.face))) return;
    const V3 lo = mk3(fminf(fminf(v0.x, v1.x), v2.x), fminf(fminf(v0.y, v1.y), v2.y), fminf(fminf(v0.z, v1.z), v2.z));
    const V3 hi = mk3(fmaxf(fmaxf(v0.x, v1.x), v2.x), fmaxf(fmaxf(v0.y, v1.y), v2.y), fmaxf(fmaxf(v0.z, v1.z), v2.z));
    float tn, tf; slab_F(r, lo, hi, tn, tf);
    if (!(tn <= mul(t, B2_C1) && tf >= mul(t, B2_C2))) return;
    best.t = t; best.face = face; best.tri = tri_idx;
}

// raw geometric normal (v1-v0) x (v2-v0) of a leaf record (Embree's Ng, |Ng| = 2*area)
B2_DEV V3 tri_ng(const BvhView& bvh, uint32_t tri_idx)
{
    const float4 a = ldg(bvh.tris + 3 * (size_t)tri_idx + 0);
    const float4 b = ldg(bvh.tris + 3 * (size_t)tri_idx + 1);
    const float4 c = ldg(bvh.tris + 3 * (size_t)tri_idx + 2);
    const V3 v0 = mk3(a.x, a.y, a.z);
    return cross_fma(v_sub(mk3(b.x, b.y, b.z), v0), v_sub(mk3(c.x, c.y, c.z), v0));
}

B2_DEV RaySetup ray_setup(V3 o, V3 d, const BvhView& bvh)
{
    RaySetup r; r.o = o; r.d = d;
    float dx = d.x, dy = d.y, dz = d.z;
    if (fabsf(dx) < 1e-18f) dx = copysignf(1e-18f, dx);
    if (fabsf(dy) < 1e-18f) dy = copysignf(1e-18f, dy);
    if (fabsf(dz) < 1e-18f) dz = copysignf(1e-18f, dz);
    r.idir = mk3(dvd(1.0f, dx), dvd(1.0f, dy), dvd(1.0f, dz));
    r.oct = (r.idir.x >= 0.f ? 1u : 0u) | (r.idir.y >= 0.f ? 2u : 0u) | (r.idir.z >= 0.f ? 4u : 0u);
    const float k20 = 9.5367431640625e-07f;   // 2^-20
    const float oix = o.x * r.idir.x, oiy = o.y * r.idir.y, oiz = o.z * r.idir.z;
    const float dlx = (fabsf(oix) + bvh.bx * fabsf(r.idir.x)) * k20, dly = (fabsf(oiy) + bvh.by * fabsf(r.idir.y)) * k20, dlz = (fabsf(oiz) + bvh.bz * fabsf(r.idir.z)) * k20;
    r.idn = mk3(r.idir.x * B2_C3P, r.idir.y * B2_C3P, r.idir.z * B2_C3P);
    r.cn = mk3((oix + dlx) * B2_C3P, (oiy + dly) * B2_C3P, (oiz + dlz) * B2_C3P);
    r.cf = mk3(oix - dlx, oiy - dly, oiz - dlz);
    // node layout in float4 units: lo_x 0..1, lo_y 2..3, lo_z 4..5, hi_x 6..7, hi_y 8..9, hi_z 10..11
    r.onx = (r.oct & 1u) ? 0u : 6u; r.ony = (r.oct & 2u) ? 2u : 8u; r.onz = (r.oct & 4u) ? 4u : 10u;
    return r;
}

// Intersect the 8 children of one node; returns the hit mask: bits 31..24 inner children by priority (slot ^ oct), bits 23..0 leaf triangles.
B2_DEV uint32_t node_test(const float4* __restrict__ np, const RaySetup& r, float tbest, uint32_t& child_base, uint32_t& tri_base, uint32_t& imask)
{
    // near planes: the "lo" arrays for positive directions, the "hi" arrays otherwise; far planes: the other one (offset 6 quads apart)
    const float4 nxa = ldg(np + r.onx), nxb = ldg(np + r.onx + 1), fxa = ldg(np + (6u - r.onx)), fxb = ldg(np + (7u - r.onx));
    const float4 nya = ldg(np + r.ony), nyb = ldg(np + r.ony + 1), fya = ldg(np + (10u - r.ony)), fyb = ldg(np + (11u - r.ony));
    const float4 nza = ldg(np + r.onz), nzb = ldg(np + r.onz + 1), fza = ldg(np + (14u - r.onz)), fzb = ldg(np + (15u - r.onz));
    const float4 h0 = ldg(np + 12), h1 = ldg(np + 13);
    child_base = f2u(h0.x); tri_base = f2u(h0.y); imask = f2u(h1.x);
    // octant permutation of the priority bits of all inner children at once: meta ^= oct where (meta & 0x18) == 0x18
    const uint32_t w0 = f2u(h0.z), w1 = f2u(h0.w);
    const uint32_t m0 = w0 ^ ((((w0 >> 3) & (w0 >> 4)) & 0x01010101u) * r.oct);
    const uint32_t m1 = w1 ^ ((((w1 >> 3) & (w1 >> 4)) & 0x01010101u) * r.oct);
    const float tlim = tbest * B2_C1;
    uint32_t hitmask = 0;
#define B2_CHILD(S, NX, NY, NZ, FX, FY, FZ, M)                                                                        \
    {                                                                                                                  \
        const float tn = fmaxf(fmaxf(fmaf(NX, r.idn.x, -r.cn.x), fmaf(NY, r.idn.y, -r.cn.y)), fmaxf(fmaf(NZ, r.idn.z, -r.cn.z), 0.0f)); \
        const float tf = fminf(fminf(fmaf(FX, r.idir.x, -r.cf.x), fmaf(FY, r.idir.y, -r.cf.y)), fminf(fmaf(FZ, r.idir.z, -r.cf.z), tlim)); \
        const uint32_t meta = ((M) >> (8 * ((S) & 3))) & 0xffu;                                                        \
        const uint32_t bits = (meta >> 5) << (meta & 0x1fu);                                                           \
        hitmask |= (tn <= tf) ? bits : 0u;                                                                             \
    }
    B2_CHILD(0, nxa.x, nya.x, nza.x, fxa.x, fya.x, fza.x, m0)
    B2_CHILD(1, nxa.y, nya.y, nza.y, fxa.y, fya.y, fza.y, m0)
    B2_CHILD(2, nxa.z, nya.z, nza.z, fxa.z, fya.z, fza.z, m0)
    B2_CHILD(3, nxa.w, nya.w, nza.w, fxa.w, fya.w, fza.w, m0)
    B2_CHILD(4, nxb.x, nyb.x, nzb.x, fxb.x, fyb.x, fzb.x, m1)
    B2_CHILD(5, nxb.y, nyb.y, nzb.y, fxb.y, fyb.y, fzb.y, m1)
    B2_CHILD(6, nxb.z, nyb.z, nzb.z, fxb.z, fyb.z, fzb.z, m1)
    B2_CHILD(7, nxb.w, nyb.w, nzb.w, fxb.w, fyb.w, fzb.w, m1)
#undef B2_CHILD
    return hitmask;
}

// Closest hit. best.t must be initialised to tfar, best.face to B2_NOFACE by the caller (see trace_init).
template <bool STATS>
B2_DEV void trace_closest(const BvhView& bvh, const RaySetup& r, HitRec& best, uint32_t& n_nodes, uint32_t& n_tris)
{
    uint2 stack[B2_TRAVERSAL_STACK];
    int sp = 0;
    // root group: slot 0 of a virtual parent -> priority bit 24 + (0 ^ oct), imask bit 0
    uint2 G = make_uint2(0u, (1u << (24 + r.oct)) | 1u);      // pending inner children: (child_base, hit bits 31..24 | imask 7..0)
    uint2 Gt = make_uint2(0u, 0u);                             // pending leaf triangles of the last visited node: (tri_base, bits 23..0)
    // Each trip does ONE unit of work per lane -- a triangle test if one is pending, otherwise a node visit -- instead of a node visit
    // followed by an inner loop over that node's triangles: a warp then never waits for the lane with the most triangles in a node
    // (profiles/r01: the max-over-lanes triangle loop dominated the slowest warps' instruction streams).  The closest hit under the
    // (t, face) rule does not depend on the order of the tests.
    while (true) {
        if (Gt.y) {
            const uint32_t i = 31u - (uint32_t)clz32(Gt.y);
            Gt.y &= ~(1u << i);
            tri_test(bvh, r, Gt.x + i, best);
            if (STATS) n_tris++;
        }
        if (!Gt.y) {                                           // last pending triangle done (or none): visit a node in the same trip
            if (!(G.y & 0xff000000u)) {
                if (sp == 0) break;
                G = stack[--sp];
            }
            const uint32_t bitpos = 31u - (uint32_t)clz32(G.y);
            G.y &= ~(1u << bitpos);
            const uint32_t slot = (bitpos - 24u) ^ r.oct;
            const uint32_t rel = popc32(G.y & 0xffu & ((1u << slot) - 1u));
            const uint32_t node_idx = G.x + rel;
            if (G.y & 0xff000000u) stack[sp++] = G;
            uint32_t child_base, tri_base, imask;
            const uint32_t hm = node_test(bvh.nodes + B2_NODE_QUADS * (size_t)node_idx, r, best.t, child_base, tri_base, imask);
            if (STATS) n_nodes++;
            G = make_uint2(child_base, (hm & 0xff000000u) | imask);
            Gt = make_uint2(tri_base, hm & 0x00ffffffu);
        }
    }
}

B2_DEV HitRec trace_init(float tfar) { HitRec h; h.t = tfar; h.face = B2_NOFACE; h.tri = 0u; return h; }
