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
// CULLING: a child box is skipped only if NOT (tn' <= fl(tbest*C1) && tf' >= max(0, fl(tn'*C3))), C3 = 1-2^-12, where tn'/tf' are
//   computed from the QUANTISED planes with FMAs and widened by a per-node, per-axis slack
//       delta_k = 2^-20 * (|(p_k - o_k) * idir_k| + 256 * |scale_k * idir_k|)
//   which dominates both the FMA-path rounding error and F's own relative error, so tn' <= tn_F(box) and tf' >= tf_F(box) for the
//   float box the quantised one contains.  Hence every HIT's leaf is visited and the result equals the brute-force answer.
#pragma once
#include "b2_math.cuh"
#include "bvh8.h"

#define B2_C1 1.00006103515625f
#define B2_C2 0.99993896484375f
#define B2_C3 0.999755859375f
#define B2_NOFACE 0xFFFFFFFFu

struct BvhView {
    const uint4*  nodes;    // 5 x uint4 per node
    const float4* tris;     // 3 x float4 per triangle record
};

struct HitRec {
    float    t;             // closest t (== tfar on a miss)
    uint32_t face;          // original face id, B2_NOFACE on a miss
    uint32_t tri;           // leaf-record index of the hit triangle
};

struct RaySetup {
    V3 o, d, idir;
    uint32_t oct;           // bit k set iff idir_k >= 0
};

B2_DEV RaySetup ray_setup(V3 o, V3 d)
{
    RaySetup r; r.o = o; r.d = d;
    float dx = d.x, dy = d.y, dz = d.z;
    if (fabsf(dx) < 1e-18f) dx = copysignf(1e-18f, dx);
    if (fabsf(dy) < 1e-18f) dy = copysignf(1e-18f, dy);
    if (fabsf(dz) < 1e-18f) dz = copysignf(1e-18f, dz);
    r.idir = mk3(dvd(1.0f, dx), dvd(1.0f, dy), dvd(1.0f, dz));
    r.oct = (r.idir.x >= 0.f ? 1u : 0u) | (r.idir.y >= 0.f ? 2u : 0u) | (r.idir.z >= 0.f ? 4u : 0u);
    return r;
}

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

// byte s of w as an exact float 2^23 + b: one PRMT, no I2F (the XU pipe was the busiest pipe with cvt, profiles/r01)
template <int S> B2_DEV float byte_magic(uint32_t w, uint32_t k4b) { return u2f(prmt_imm<0x7540 + S>(w, k4b)); }

#define B2_C3P 0.99951171875f          // 1 - 2^-11 : C3 folded into the near planes (see below)

// Intersect the 8 children of one node; returns the hit mask: bits 31..24 inner children by priority (slot ^ oct), bits 23..0 leaf triangles.
//
// Per axis k (ad = scale*idir, ao = (p - o)*idir) the real-valued child planes are t = q*ad + ao, q in 0..255.  Computed form:
//     far :  tf_k = fma(q', ad,       cf),  cf = (ao + dl) - 2^23*ad                q' = 2^23 + q (exact float, byte_magic)
//     near:  tn_k = fma(q', ad*C3',   cn),  cn = (ao - dl)*C3' - 2^23*(ad*C3')      C3' = 1 - 2^-11
//     dl = |ao|*2^-20 + |ad|*(1 + 2^-12)
// The 2^23 offset of q' cancels exactly against the constant (same rounded ad); rounding the constant costs at most |ad|/2 and is
// covered by the |ad| term of dl; the remaining dl >= 2^-20(|ao| + 256|ad|) dominates the FMA-path error and F's relative error
// (trace.cuh header).  Scaling the near side by C3' makes  max(tn_x,tn_y,tn_z,0) <= min(tf_x,tf_y,tf_z,tbest*C1)  at least as
// permissive as the oracle's visit rule V (tn <= tbest*C1 && tf >= max(0, tn*C3)) on the float box inside the quantised one.
B2_DEV uint32_t node_test(const uint4* __restrict__ np, const RaySetup& r, float tbest, uint32_t& child_base, uint32_t& tri_base, uint32_t& imask)
{
    const uint4 n0 = ldg(np + 0), n1 = ldg(np + 1), n2 = ldg(np + 2), n3 = ldg(np + 3), n4 = ldg(np + 4);
    const float px = u2f(n0.x), py = u2f(n0.y), pz = u2f(n0.z);
    const float sx = u2f((n0.w & 0xffu) << 23), sy = u2f(((n0.w >> 8) & 0xffu) << 23), sz = u2f(((n0.w >> 16) & 0xffu) << 23);
    imask = n0.w >> 24;
    child_base = n1.x; tri_base = n1.y;

    const float adx = sx * r.idir.x, ady = sy * r.idir.y, adz = sz * r.idir.z;
    const float aox = (px - r.o.x) * r.idir.x, aoy = (py - r.o.y) * r.idir.y, aoz = (pz - r.o.z) * r.idir.z;
    const float k20 = 9.5367431640625e-07f, k1 = 1.000244140625f, two23 = 8388608.0f;
    const float dlx = fmaf(fabsf(adx), k1, fabsf(aox) * k20), dly = fmaf(fabsf(ady), k1, fabsf(aoy) * k20), dlz = fmaf(fabsf(adz), k1, fabsf(aoz) * k20);
    const float anx = adx * B2_C3P, any_ = ady * B2_C3P, anz = adz * B2_C3P;
    const float cnx = fmaf(-two23, anx, (aox - dlx) * B2_C3P), cny = fmaf(-two23, any_, (aoy - dly) * B2_C3P), cnz = fmaf(-two23, anz, (aoz - dlz) * B2_C3P);
    const float cfx = fmaf(-two23, adx, aox + dlx), cfy = fmaf(-two23, ady, aoy + dly), cfz = fmaf(-two23, adz, aoz + dlz);

    // near/far byte planes by ray direction sign
    const bool nx = r.idir.x < 0.f, ny = r.idir.y < 0.f, nz = r.idir.z < 0.f;
    const uint32_t qnx0 = nx ? n3.z : n2.x, qnx1 = nx ? n3.w : n2.y, qfx0 = nx ? n2.x : n3.z, qfx1 = nx ? n2.y : n3.w;
    const uint32_t qny0 = ny ? n4.x : n2.z, qny1 = ny ? n4.y : n2.w, qfy0 = ny ? n2.z : n4.x, qfy1 = ny ? n2.w : n4.y;
    const uint32_t qnz0 = nz ? n4.z : n3.x, qnz1 = nz ? n4.w : n3.y, qfz0 = nz ? n3.x : n4.z, qfz1 = nz ? n3.y : n4.w;

    // octant permutation of the priority bits of all inner children at once: meta ^= oct where (meta & 0x18) == 0x18
    const uint32_t m0 = n1.z ^ ((((n1.z >> 3) & (n1.z >> 4)) & 0x01010101u) * r.oct);
    const uint32_t m1 = n1.w ^ ((((n1.w >> 3) & (n1.w >> 4)) & 0x01010101u) * r.oct);

    const float tlim = tbest * B2_C1;
    const uint32_t k4b = opaque_const(0x4B000000u);
    uint32_t hitmask = 0;
#define B2_CHILD(S, QNX, QNY, QNZ, QFX, QFY, QFZ, M)                                                    \
    {                                                                                                    \
        const float tnx = fmaf(byte_magic<(S) & 3>(QNX, k4b), anx, cnx);                                 \
        const float tny = fmaf(byte_magic<(S) & 3>(QNY, k4b), any_, cny);                                \
        const float tnz = fmaf(byte_magic<(S) & 3>(QNZ, k4b), anz, cnz);                                 \
        const float tfx = fmaf(byte_magic<(S) & 3>(QFX, k4b), adx, cfx);                                 \
        const float tfy = fmaf(byte_magic<(S) & 3>(QFY, k4b), ady, cfy);                                 \
        const float tfz = fmaf(byte_magic<(S) & 3>(QFZ, k4b), adz, cfz);                                 \
        const float tn = fmaxf(fmaxf(tnx, tny), fmaxf(tnz, 0.0f));                                       \
        const float tf = fminf(fminf(tfx, tfy), fminf(tfz, tlim));                                       \
        const uint32_t meta = ((M) >> (8 * ((S) & 3))) & 0xffu;                                          \
        const uint32_t bits = (meta >> 5) << (meta & 0x1fu);                                             \
        hitmask |= (tn <= tf) ? bits : 0u;                                                               \
    }
    B2_CHILD(0, qnx0, qny0, qnz0, qfx0, qfy0, qfz0, m0)
    B2_CHILD(1, qnx0, qny0, qnz0, qfx0, qfy0, qfz0, m0)
    B2_CHILD(2, qnx0, qny0, qnz0, qfx0, qfy0, qfz0, m0)
    B2_CHILD(3, qnx0, qny0, qnz0, qfx0, qfy0, qfz0, m0)
    B2_CHILD(4, qnx1, qny1, qnz1, qfx1, qfy1, qfz1, m1)
    B2_CHILD(5, qnx1, qny1, qnz1, qfx1, qfy1, qfz1, m1)
    B2_CHILD(6, qnx1, qny1, qnz1, qfx1, qfy1, qfz1, m1)
    B2_CHILD(7, qnx1, qny1, qnz1, qfx1, qfy1, qfz1, m1)
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
    uint2 G = make_uint2(0u, (1u << (24 + r.oct)) | 1u);
    uint2 Gt = make_uint2(0u, 0u);
    while (true) {
        if (G.y & 0xff000000u) {
            const uint32_t bitpos = 31u - (uint32_t)clz32(G.y);
            G.y &= ~(1u << bitpos);
            const uint32_t slot = (bitpos - 24u) ^ r.oct;
            const uint32_t rel = popc32(G.y & 0xffu & ((1u << slot) - 1u));
            const uint32_t node_idx = G.x + rel;
            if (G.y & 0xff000000u) stack[sp++] = G;
            uint32_t child_base, tri_base, imask;
            const uint32_t hm = node_test(bvh.nodes + 5 * (size_t)node_idx, r, best.t, child_base, tri_base, imask);
            if (STATS) n_nodes++;
            G = make_uint2(child_base, (hm & 0xff000000u) | imask);
            Gt = make_uint2(tri_base, hm & 0x00ffffffu);
        } else {
            Gt = G; G = make_uint2(0u, 0u);
        }
        while (Gt.y) {
            const uint32_t i = 31u - (uint32_t)clz32(Gt.y);
            Gt.y &= ~(1u << i);
            tri_test(bvh, r, Gt.x + i, best);
            if (STATS) n_tris++;
        }
        if (!(G.y & 0xff000000u)) {
            if (sp == 0) break;
            G = stack[--sp];
        }
    }
}

B2_DEV HitRec trace_init(float tfar) { HitRec h; h.t = tfar; h.face = B2_NOFACE; h.tri = 0u; return h; }
