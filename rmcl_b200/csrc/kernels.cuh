// kernels.cuh -- CUDA kernels of the ray-casting-correspondence path (sm_100a).  Compiled with -fmad=false: every FMA is explicit.
#pragma once
#include <cstddef>
#include "trace.cuh"
#if defined(__CUDACC__)
#include <cooperative_groups.h>
#endif

// ---------------------------------------------------------------------------------------------------------------------
// device-side CrossStatistics helpers (mirror oracle/oracle.c op for op; tolerance-level parity is all that is required,
// but identical order keeps whole ICP chains bit-identical in practice)
// ---------------------------------------------------------------------------------------------------------------------
struct CStats { V3 dm, mm; float C[9]; uint32_t n; };     // C column-major [c*3+r]

B2_DEV CStats cs_identity() { CStats s; s.dm = mk3(0, 0, 0); s.mm = mk3(0, 0, 0); for (int i = 0; i < 9; i++) s.C[i] = 0.f; s.n = 0; return s; }
B2_DEV CStats cs_load(const b2_cross_stats* p)
{
    CStats s; s.dm = mk3(p->dataset_mean.x, p->dataset_mean.y, p->dataset_mean.z); s.mm = mk3(p->model_mean.x, p->model_mean.y, p->model_mean.z);
    for (int i = 0; i < 9; i++) s.C[i] = p->covariance.m[i]; s.n = p->n_meas; return s;
}
B2_DEV void cs_store(b2_cross_stats* p, const CStats& s)
{
    p->dataset_mean.x = s.dm.x; p->dataset_mean.y = s.dm.y; p->dataset_mean.z = s.dm.z;
    p->model_mean.x = s.mm.x; p->model_mean.y = s.mm.y; p->model_mean.z = s.mm.z;
    for (int i = 0; i < 9; i++) p->covariance.m[i] = s.C[i]; p->n_meas = s.n;
}
// rm::CrossStatistics::operator+= (oracle: orc_cross_stats_merge)
B2_DEV CStats cs_merge(const CStats& a, const CStats& b)
{
    CStats r;
    const uint32_t n = a.n + b.n;
    if (n == 0) return cs_identity();
    const float w1 = dvd((float)a.n, (float)n), w2 = dvd((float)b.n, (float)n);
    r.n = n;
    r.dm = v_add(v_scale(a.dm, w1), v_scale(b.dm, w2));
    r.mm = v_add(v_scale(a.mm, w1), v_scale(b.mm, w2));
    const V3 ma = v_sub(a.mm, r.mm), da = v_sub(a.dm, r.dm), mb = v_sub(b.mm, r.mm), db = v_sub(b.dm, r.dm);
    const float mav[3] = {ma.x, ma.y, ma.z}, dav[3] = {da.x, da.y, da.z}, mbv[3] = {mb.x, mb.y, mb.z}, dbv[3] = {db.x, db.y, db.z};
    #pragma unroll
    for (int c = 0; c < 3; c++)
        #pragma unroll
        for (int rr = 0; rr < 3; rr++) {
            const float p1 = add(mul(a.C[c * 3 + rr], w1), mul(b.C[c * 3 + rr], w2));
            const float p2 = add(mul(mul(mav[rr], dav[c]), w1), mul(mul(mbv[rr], dbv[c]), w2));
            r.C[c * 3 + rr] = add(p1, p2);
        }
    return r;
}
// Transform * CrossStatistics (oracle: orc_cross_stats_transform)
B2_DEV CStats cs_transform(Tf T, const CStats& s)
{
    CStats r; r.n = s.n; r.dm = tf_apply(T, s.dm); r.mm = tf_apply(T, s.mm);
    const float x = T.R.x, y = T.R.y, z = T.R.z, w = T.R.w;
    float R[3][3];
    R[0][0] = sub(1.0f, mul(2.0f, add(mul(y, y), mul(z, z)))); R[0][1] = mul(2.0f, sub(mul(x, y), mul(z, w))); R[0][2] = mul(2.0f, add(mul(x, z), mul(y, w)));
    R[1][0] = mul(2.0f, add(mul(x, y), mul(z, w))); R[1][1] = sub(1.0f, mul(2.0f, add(mul(x, x), mul(z, z)))); R[1][2] = mul(2.0f, sub(mul(y, z), mul(x, w)));
    R[2][0] = mul(2.0f, sub(mul(x, z), mul(y, w))); R[2][1] = mul(2.0f, add(mul(y, z), mul(x, w))); R[2][2] = sub(1.0f, mul(2.0f, add(mul(x, x), mul(y, y))));
    float RC[3][3];
    #pragma unroll
    for (int i = 0; i < 3; i++)
        #pragma unroll
        for (int j = 0; j < 3; j++) { float acc = 0.0f; for (int k = 0; k < 3; k++) acc = add(acc, mul(R[i][k], s.C[j * 3 + k])); RC[i][j] = acc; }
    #pragma unroll
    for (int i = 0; i < 3; i++)
        #pragma unroll
        for (int j = 0; j < 3; j++) { float acc = 0.0f; for (int k = 0; k < 3; k++) acc = add(acc, mul(RC[i][k], R[j][k])); r.C[j * 3 + i] = acc; }
    return r;
}

// ---------------------------------------------------------------------------------------------------------------------
// rm::umeyama_transform (micp_localization.cpp:952-953): 3x3 SVD in double, same algorithm / order as oracle svd3()
// ---------------------------------------------------------------------------------------------------------------------
B2_DEV double det3d(const double A[3][3])
{
    return A[0][0] * (A[1][1] * A[2][2] - A[1][2] * A[2][1]) - A[0][1] * (A[1][0] * A[2][2] - A[1][2] * A[2][0])
         + A[0][2] * (A[1][0] * A[2][1] - A[1][1] * A[2][0]);
}

__host__ __device__ __noinline__ void svd3_dev(const double A[3][3], double U[3][3], double w[3], double V[3][3])
{
    double B[3][3];
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { B[i][j] = A[i][j]; V[i][j] = (i == j) ? 1.0 : 0.0; }
    for (int sweep = 0; sweep < 60; sweep++) {
        double off = 0.0;
        for (int p = 0; p < 2; p++) for (int q = p + 1; q < 3; q++) {
            double alpha = 0, beta = 0, gamma = 0;
            for (int i = 0; i < 3; i++) { alpha += B[i][p] * B[i][p]; beta += B[i][q] * B[i][q]; gamma += B[i][p] * B[i][q]; }
            if (gamma == 0.0) continue;
            const double lim = 1e-30 + 1e-16 * sqrt(alpha * beta);
            if (fabs(gamma) <= lim) continue;
            off += fabs(gamma);
            const double zeta = (beta - alpha) / (2.0 * gamma);
            const double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
            const double c = 1.0 / sqrt(1.0 + t * t), sn = c * t;
            for (int i = 0; i < 3; i++) {
                const double bp = B[i][p], bq = B[i][q]; B[i][p] = c * bp - sn * bq; B[i][q] = sn * bp + c * bq;
                const double vp = V[i][p], vq = V[i][q]; V[i][p] = c * vp - sn * vq; V[i][q] = sn * vp + c * vq;
            }
        }
        if (off == 0.0) break;
    }
    int idx[3] = {0, 1, 2}; double nrm[3];
    for (int j = 0; j < 3; j++) nrm[j] = sqrt(B[0][j] * B[0][j] + B[1][j] * B[1][j] + B[2][j] * B[2][j]);
    for (int a = 0; a < 2; a++) for (int b = a + 1; b < 3; b++) if (nrm[idx[b]] > nrm[idx[a]]) { int t = idx[a]; idx[a] = idx[b]; idx[b] = t; }
    double Vs[3][3], Bs[3][3];
    for (int j = 0; j < 3; j++) { w[j] = nrm[idx[j]]; for (int i = 0; i < 3; i++) { Vs[i][j] = V[i][idx[j]]; Bs[i][j] = B[i][idx[j]]; } }
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) V[i][j] = Vs[i][j];
    const double tiny = 1e-12 * (w[0] > 0 ? w[0] : 1.0);
    bool good[3];
    for (int j = 0; j < 3; j++) {
        good[j] = w[j] > tiny;
        if (good[j]) for (int i = 0; i < 3; i++) U[i][j] = Bs[i][j] / w[j];
    }
    if (!good[0]) { for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) U[i][j] = (i == j) ? 1.0 : 0.0; return; }
    if (!good[1]) {
        const double a[3] = {U[0][0], U[1][0], U[2][0]};
        const int k = fabs(a[0]) < fabs(a[1]) ? (fabs(a[0]) < fabs(a[2]) ? 0 : 2) : (fabs(a[1]) < fabs(a[2]) ? 1 : 2);
        double e[3] = {0, 0, 0}; e[k] = 1.0;
        const double b[3] = {a[1] * e[2] - a[2] * e[1], a[2] * e[0] - a[0] * e[2], a[0] * e[1] - a[1] * e[0]};
        const double nb = sqrt(b[0] * b[0] + b[1] * b[1] + b[2] * b[2]);
        for (int i = 0; i < 3; i++) U[i][1] = b[i] / nb;
    }
    if (!good[2] || !good[1]) {
        const double c[3] = {U[1][0] * U[2][1] - U[2][0] * U[1][1], U[2][0] * U[0][1] - U[0][0] * U[2][1], U[0][0] * U[1][1] - U[1][0] * U[0][1]};
        const double sgn = det3d(V) < 0 ? -1.0 : 1.0;
        for (int i = 0; i < 3; i++) U[i][2] = sgn * c[i];
    }
}

#if B2_ON_DEVICE
B2_DEV float b2_rcp_approx(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
B2_DEV float b2_rsqrt_approx(float x) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
#else
B2_DEV float b2_rcp_approx(float x) { return 1.0f / x; }
B2_DEV float b2_rsqrt_approx(float x) { return 1.0f / std::sqrt(x); }
#endif

// orthogonal polar factor of a 3x3 matrix with positive determinant; false if not applicable / not converged.
// Frobenius-scaled Newton iteration X <- (g X + X^-T / g) / 2 in FP32 (4-cycle ops, MUFU-based div/sqrt: ~3x shorter serial chain than
// FP64 on the one thread that executes it), then one unscaled FP64 step that makes the result orthogonal to double precision.
// Accuracy: FP32 rounding of the iterates perturbs the rotation by ~1e-7 rad (the reference's own rmagine SVD is FP32 as well); the
// translation inherits ~1e-7 * |mean| -- two orders below the 1e-5 tolerance on dT.
__host__ __device__ __noinline__ bool polar_newton3(const float A[3][3], double Q[3][3])
{
    // scale to unit Frobenius norm so that the iteration starts in its well-behaved range; the polar factor does not depend on the scale,
    // so the approximate reciprocal square root is as good as the exact one here
    float fro = 0.0f;
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) fro += A[i][j] * A[i][j];
    if (!(fro > 1e-30f)) return false;
    const float inv_n = b2_rsqrt_approx(fro);
    float X[3][3];
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) X[i][j] = A[i][j] * inv_n;
    bool conv = false;
    for (int it = 0; it < 40 && !conv; it++) {
        float Cf[3][3];                                           // cofactor matrix = det * X^-T
        Cf[0][0] = X[1][1] * X[2][2] - X[1][2] * X[2][1]; Cf[0][1] = X[1][2] * X[2][0] - X[1][0] * X[2][2]; Cf[0][2] = X[1][0] * X[2][1] - X[1][1] * X[2][0];
        Cf[1][0] = X[0][2] * X[2][1] - X[0][1] * X[2][2]; Cf[1][1] = X[0][0] * X[2][2] - X[0][2] * X[2][0]; Cf[1][2] = X[0][1] * X[2][0] - X[0][0] * X[2][1];
        Cf[2][0] = X[0][1] * X[1][2] - X[0][2] * X[1][1]; Cf[2][1] = X[0][2] * X[1][0] - X[0][0] * X[1][2]; Cf[2][2] = X[0][0] * X[1][1] - X[0][1] * X[1][0];
        const float det = X[0][0] * Cf[0][0] + X[0][1] * Cf[0][1] + X[0][2] * Cf[0][2];
        if (!(det > 1e-12f)) return false;                        // reflection, singular or NaN: let the SVD decide
        float a, b;
        if (it < 3) {
            // Frobenius scaling g^2 = |X^-T|_F / |X|_F only matters while X is far from orthogonal; approximate reciprocal / rsqrt are
            // enough (any g > 0 keeps the iteration convergent, and its errors are corrected by the following steps)
            float nx = 0.f, nc = 0.f;
            for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { nx += X[i][j] * X[i][j]; nc += Cf[i][j] * Cf[i][j]; }
            const float rdet = b2_rcp_approx(det);
            const float q = nc * b2_rcp_approx(nx);                                  // (|Cf|/|X|)^2
            const float g2 = q * b2_rsqrt_approx(q) * rdet;                          // sqrt(q) / det
            const float rg = b2_rsqrt_approx(g2);                                    // 1 / g
            a = 0.5f * g2 * rg; b = 0.5f * rg * rdet;
        } else {
            a = 0.5f; b = 0.5f * b2_rcp_approx(det);
        }
        float diff = 0.f;
        for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
            const float y = a * X[i][j] + b * Cf[i][j];
            const float d = y - X[i][j]; diff += d * d;
            X[i][j] = y;
        }
        conv = (it >= 3) && diff < 1e-10f;                        // |dX|_F < 1e-5 on an unscaled step: the FP64 step finishes the job
    }
    if (!conv) return false;
    double Y[3][3];
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) Y[i][j] = (double)X[i][j];
    for (int it = 0; it < 1; it++) {                              // unscaled FP64 step: orthogonality 1e-6 -> 1e-12
        double Cf[3][3];
        Cf[0][0] = Y[1][1] * Y[2][2] - Y[1][2] * Y[2][1]; Cf[0][1] = Y[1][2] * Y[2][0] - Y[1][0] * Y[2][2]; Cf[0][2] = Y[1][0] * Y[2][1] - Y[1][1] * Y[2][0];
        Cf[1][0] = Y[0][2] * Y[2][1] - Y[0][1] * Y[2][2]; Cf[1][1] = Y[0][0] * Y[2][2] - Y[0][2] * Y[2][0]; Cf[1][2] = Y[0][1] * Y[2][0] - Y[0][0] * Y[2][1];
        Cf[2][0] = Y[0][1] * Y[1][2] - Y[0][2] * Y[1][1]; Cf[2][1] = Y[0][2] * Y[1][0] - Y[0][0] * Y[1][2]; Cf[2][2] = Y[0][0] * Y[1][1] - Y[0][1] * Y[1][0];
        const double det = Y[0][0] * Cf[0][0] + Y[0][1] * Cf[0][1] + Y[0][2] * Cf[0][2];
        const double b = 0.5 / det;
        for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) Y[i][j] = 0.5 * Y[i][j] + b * Cf[i][j];
    }
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) Q[i][j] = Y[i][j];
    return true;
}

__host__ __device__ __noinline__ Tf umeyama_dev(const CStats& s)
{
    Tf out = tf_identity();
    if (s.n == 0) return out;
    float Cf[3][3]; double R[3][3];
    for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) Cf[r][c] = s.C[c * 3 + r];
    // Fast path: for det(C) > 0 the Umeyama rotation U S V^T (S = I) is the orthogonal polar factor of C, obtained with the scaled
    // Newton iteration X <- (g X + X^-T / g) / 2 (quadratically convergent, ~100 serial FP64 instructions per step instead of the
    // ~4000 of a Jacobi SVD -- this code runs on ONE thread between two reductions of the ICP loop, so its latency is the step's).
    // det(C) <= 0 (reflection case), a singular C or no convergence fall back to the SVD.
    if (!polar_newton3(Cf, R)) {
        double C[3][3], U[3][3], V[3][3], w[3];
        for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) C[r][c] = (double)Cf[r][c];
        svd3_dev(C, U, w, V);
        const double sgn = (det3d(U) * det3d(V) < 0.0) ? -1.0 : 1.0;
        for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) R[i][j] = U[i][0] * V[j][0] + U[i][1] * V[j][1] + sgn * U[i][2] * V[j][2];
    }
    // rotation matrix -> quaternion in FP32 (the reference's rm::Quaternion::set(Matrix3x3) is FP32 as well); R is orthogonal to 1e-12
    float Rf[3][3];
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) Rf[i][j] = (float)R[i][j];
    float q[4];
    const float tr = Rf[0][0] + Rf[1][1] + Rf[2][2];
    if (tr > 0.0f) {
        const float sc = sqrt_rn(tr + 1.0f) * 2.0f; q[3] = 0.25f * sc;
        q[0] = dvd(Rf[2][1] - Rf[1][2], sc); q[1] = dvd(Rf[0][2] - Rf[2][0], sc); q[2] = dvd(Rf[1][0] - Rf[0][1], sc);
    } else if (Rf[0][0] > Rf[1][1] && Rf[0][0] > Rf[2][2]) {
        const float sc = sqrt_rn(1.0f + Rf[0][0] - Rf[1][1] - Rf[2][2]) * 2.0f; q[3] = dvd(Rf[2][1] - Rf[1][2], sc);
        q[0] = 0.25f * sc; q[1] = dvd(Rf[0][1] + Rf[1][0], sc); q[2] = dvd(Rf[0][2] + Rf[2][0], sc);
    } else if (Rf[1][1] > Rf[2][2]) {
        const float sc = sqrt_rn(1.0f + Rf[1][1] - Rf[0][0] - Rf[2][2]) * 2.0f; q[3] = dvd(Rf[0][2] - Rf[2][0], sc);
        q[0] = dvd(Rf[0][1] + Rf[1][0], sc); q[1] = 0.25f * sc; q[2] = dvd(Rf[1][2] + Rf[2][1], sc);
    } else {
        const float sc = sqrt_rn(1.0f + Rf[2][2] - Rf[0][0] - Rf[1][1]) * 2.0f; q[3] = dvd(Rf[1][0] - Rf[0][1], sc);
        q[0] = dvd(Rf[0][2] + Rf[2][0], sc); q[1] = dvd(Rf[1][2] + Rf[2][1], sc); q[2] = 0.25f * sc;
    }
    Q4 qq; qq.x = q[0]; qq.y = q[1]; qq.z = q[2]; qq.w = q[3];
    out.R = q_normalize(qq);
    out.t = v_sub(s.mm, q_rot(out.R, s.dm));
    return out;
}

// ---------------------------------------------------------------------------------------------------------------------
// P2L accumulation (rm::statistics_p2l, called at CorrespondencesCPU.cpp:26-30): FP32 per-element math identical to the oracle
// (=> identical gating, identical n_meas), sums in FP64 "sum form": n, S_d, S_m, S_{m d^T}
// ---------------------------------------------------------------------------------------------------------------------
#define B2_NACC 15
struct P2LAcc { double v[B2_NACC]; uint32_t n; };

B2_DEV void acc_zero(P2LAcc& a) { for (int i = 0; i < B2_NACC; i++) a.v[i] = 0.0; a.n = 0; }
B2_DEV void acc_add_pair(P2LAcc& a, V3 D, V3 M)
{
    const double d[3] = {(double)D.x, (double)D.y, (double)D.z}, m[3] = {(double)M.x, (double)M.y, (double)M.z};
    a.v[0] += d[0]; a.v[1] += d[1]; a.v[2] += d[2];
    a.v[3] += m[0]; a.v[4] += m[1]; a.v[5] += m[2];
    #pragma unroll
    for (int c = 0; c < 3; c++)
        #pragma unroll
        for (int r = 0; r < 3; r++) a.v[6 + c * 3 + r] += m[r] * d[c];
    a.n++;
}
// P2L gate for one pair; returns true and (D, M) if accepted
B2_DEV bool p2l_pair(Tf Tpre, V3 d, V3 I, V3 N, float max_dist, V3& D, V3& M)
{
    D = tf_apply(Tpre, d);
    const float sd = v_dot(v_sub(I, D), N);
    if (!(fabsf(sd) < max_dist)) return false;
    M = v_add(D, v_scale(N, sd));
    return true;
}
B2_DEV CStats acc_finalize(const double* v, uint32_t n)
{
    CStats s = cs_identity();
    if (n == 0) return s;
    const double inv = 1.0 / (double)n;
    double dm[3], mm[3];
    for (int k = 0; k < 3; k++) { dm[k] = v[k] * inv; mm[k] = v[3 + k] * inv; }
    s.dm = mk3((float)dm[0], (float)dm[1], (float)dm[2]);
    s.mm = mk3((float)mm[0], (float)mm[1], (float)mm[2]);
    for (int c = 0; c < 3; c++) for (int r = 0; r < 3; r++) s.C[c * 3 + r] = (float)(v[6 + c * 3 + r] * inv - mm[r] * dm[c]);
    s.n = n;
    return s;
}

// block-level sum of the accumulators; result valid in thread 0.  smem: double[(B2_NACC+1) * (BLOCK/32)]
//
// Warp stage = reduce-scatter: at each of four steps a lane hands half of its remaining values to its partner and adds the partner's half
// of the values it keeps, so 16 values cost 8+4+2+1 (+1 final) = 16 64-bit shuffles instead of 16 x 5 = 80 for 16 independent butterflies
// (the FP64 shuffles of the plain version were the longest part of the reduction pass: profiles/r01, clock stamps in k_icp_loop).
// After the four steps lane l holds value index ((l>>1) & 15) summed over the 16 lanes that share its bit 0; one xor-1 step completes it.
#if defined(__CUDACC__)
template <int H>
__device__ __forceinline__ void rs_step(double (&v)[16], int offset, bool upper)
{
    #pragma unroll
    for (int i = 0; i < H; i++) {
        const double send = upper ? v[i] : v[i + H];
        const double keep = upper ? v[i + H] : v[i];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, offset);
    }
}
template <int BLOCK>
__device__ __forceinline__ void block_reduce_acc(P2LAcc& a, double* smem)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int NW = BLOCK / 32;
    double v[16];
    #pragma unroll
    for (int i = 0; i < B2_NACC; i++) v[i] = a.v[i];
    v[15] = 0.0;
    rs_step<8>(v, 16, (lane & 16) != 0);      // keeps values [8,16) if bit 4 set else [0,8)
    rs_step<4>(v, 8, (lane & 8) != 0);
    rs_step<2>(v, 4, (lane & 4) != 0);
    rs_step<1>(v, 2, (lane & 2) != 0);
    const double tot = v[0] + __shfl_xor_sync(0xffffffffu, v[0], 1);
    const int vidx = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
    a.n = warp_sum_u32(a.n);
    if ((lane & 1) == 0 && vidx < B2_NACC) smem[vidx * NW + warp] = tot;
    if (lane == 0) smem[B2_NACC * NW + warp] = (double)a.n;
    __syncthreads();
    if (warp == 0) {
        // lane i < 16 sums value i over the NW warps (fixed order)
        if (lane <= B2_NACC) {
            double x = 0.0;
            #pragma unroll
            for (int w = 0; w < NW; w++) x += smem[lane * NW + w];
            smem[lane * NW] = x;
        }
        __syncwarp();
        if (lane == 0) {
            #pragma unroll
            for (int i = 0; i < B2_NACC; i++) a.v[i] = smem[i * NW];
            a.n = (uint32_t)(smem[B2_NACC * NW] + 0.5);
        }
    }
    __syncthreads();
}
#endif

// ---------------------------------------------------------------------------------------------------------------------
// ICP state kept on the device between the kernels of one correctOnce (micp_localization.cpp:899-984)
// ---------------------------------------------------------------------------------------------------------------------
struct alignas(16) IcpState {
    b2_transform Tom, Tbo, Tsb;          // inputs
    b2_transform T_onew_oold;            // :910, :963
    b2_transform T_snew_sold;            // pre-transform of the NEXT reduction (MICPSensor.hpp:178)
    b2_transform Tom_new;                // :972-984
    b2_transform Tos, Tso;               // Tos = Tbo * Tsb (sensor -> odom), Tso = ~Tos: composed once per correctOnce on the host
    b2_cross_stats Cmerged_o;            // :918-937 (single sensor)
    b2_cross_stats stats_s;              // last sensor-frame statistics
    float max_dist;
    uint32_t iter;
    float Ros[9];                        // rotation matrix of Tos, row-major
    uint32_t pad_[1];
    unsigned long long dbg[8];           // SM-clock stamps of the last reduction's tail (profiling aid, b2_rcc_debug_clocks)
};

// rotation matrix (row-major) of a unit quaternion
B2_DEV void quat_to_mat(Q4 q, float R[9])
{
    const float x = q.x, y = q.y, z = q.z, w = q.w;
    R[0] = sub(1.0f, mul(2.0f, add(mul(y, y), mul(z, z)))); R[1] = mul(2.0f, sub(mul(x, y), mul(z, w))); R[2] = mul(2.0f, add(mul(x, z), mul(y, w)));
    R[3] = mul(2.0f, add(mul(x, y), mul(z, w))); R[4] = sub(1.0f, mul(2.0f, add(mul(x, x), mul(z, z)))); R[5] = mul(2.0f, sub(mul(y, z), mul(x, w)));
    R[6] = mul(2.0f, sub(mul(x, z), mul(y, w))); R[7] = mul(2.0f, add(mul(y, z), mul(x, w))); R[8] = sub(1.0f, mul(2.0f, add(mul(x, x), mul(y, y))));
}

static_assert(offsetof(IcpState, Tos) % 16 == 0 && offsetof(IcpState, Tso) % 16 == 0 && offsetof(IcpState, Cmerged_o) % 16 == 0 && sizeof(IcpState) % 16 == 0, "IcpState alignment");

B2_DEV Tf icp_pretransform(Tf Tbo, Tf Tsb, Tf T_onew_oold)
{
    const Tf T_bnew_bold = tf_mul(tf_mul(tf_inv(Tbo), T_onew_oold), Tbo);      // micp_localization.cpp:926
    return tf_mul(tf_mul(tf_inv(Tsb), T_bnew_bold), Tsb);                      // MICPSensor.hpp:178
}

// one inner iteration after the reduction delivered stats_s (thread 0 only)
B2_DEV void icp_step(IcpState* st, const CStats& stats_s)
{
    const Tf Tbo = tf_load(&st->Tbo), Tsb = tf_load(&st->Tsb), Tom = tf_load(&st->Tom);
    Tf T_onew_oold = tf_load(&st->T_onew_oold);
    const CStats stats_b = cs_transform(Tsb, stats_s);                        // MICPSensor.hpp:182
    const CStats Cs_o = cs_transform(Tbo, stats_b);                           // micp_localization.cpp:931
    const CStats Cmerged = cs_merge(cs_identity(), Cs_o);                     // :918,:936
    const Tf T_inner = umeyama_dev(Cmerged);                                  // :952-953
    T_onew_oold = tf_mul(T_onew_oold, T_inner);                               // :963
    tf_store(&st->T_onew_oold, T_onew_oold);
    tf_store(&st->T_snew_sold, icp_pretransform(Tbo, Tsb, T_onew_oold));
    Tf Tn = tf_mul(Tom, T_onew_oold);                                         // :972
    if (Cmerged.n > 0) Tn.R = q_normalize(Tn.R); else Tn = Tom;               // :974-984
    tf_store(&st->Tom_new, Tn);
    cs_store(&st->Cmerged_o, Cmerged);
    cs_store(&st->stats_s, stats_s);
    st->iter++;
}

// ---------------------------------------------------------------------------------------------------------------------
// generic closest hit for arbitrary rays (b2_mesh_intersect)
// ---------------------------------------------------------------------------------------------------------------------
template <bool STATS>
__global__ void __launch_bounds__(128) k_intersect(BvhView bvh, const float* __restrict__ origs, const float* __restrict__ dirs, uint32_t n, float tfar,
                                                   float* __restrict__ t_out, uint32_t* __restrict__ face_out, float* __restrict__ ng_out, uint8_t* __restrict__ hit_out,
                                                   unsigned long long* __restrict__ counters)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t nn = 0, nt = 0;
    if (i < n) {
        const RaySetup r = ray_setup(mk3(origs[3 * i], origs[3 * i + 1], origs[3 * i + 2]), mk3(dirs[3 * i], dirs[3 * i + 1], dirs[3 * i + 2]), bvh);
        HitRec h = trace_init(tfar);
        trace_closest<STATS>(bvh, r, h, nn, nt);
        const bool hit = h.face != B2_NOFACE;
        if (t_out) t_out[i] = hit ? h.t : u2f(0x7f800000u);
        if (face_out) face_out[i] = h.face;
        if (hit_out) hit_out[i] = hit ? 1 : 0;
        if (ng_out) {
            V3 ng = mk3(0.f, 0.f, 0.f);
            if (hit) ng = tri_ng(bvh, h.tri);
            ng_out[3 * i] = ng.x; ng_out[3 * i + 1] = ng.y; ng_out[3 * i + 2] = ng.z;
        }
    }
    if (STATS) {
        const uint32_t sn = warp_sum_u32(nn), stt = warp_sum_u32(nt);
        if ((threadIdx.x & 31) == 0) { atomicAdd(counters, (unsigned long long)sn); atomicAdd(counters + 1, (unsigned long long)stt); }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// MICP*Sensor*::unpackMessage on the device (MICPSphericalSensorCPU.cpp:181-233)
// ---------------------------------------------------------------------------------------------------------------------
__global__ void k_dataset_from_ranges(const float* __restrict__ ranges, const float* __restrict__ dirs, const float* __restrict__ origs, uint32_t n_origs,
                                      uint32_t n, float range_min, float range_max, float* __restrict__ pts, uint8_t* __restrict__ mask)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float r = ranges[i];
    const uint32_t oi = n_origs == 1 ? 0 : i;
    pts[3 * i + 0] = add(mul(dirs[3 * i + 0], r), origs[3 * oi + 0]);
    pts[3 * i + 1] = add(mul(dirs[3 * i + 1], r), origs[3 * oi + 1]);
    pts[3 * i + 2] = add(mul(dirs[3 * i + 2], r), origs[3 * oi + 2]);
    mask[i] = (r < range_min || r > range_max) ? 0 : 1;
}

// ---------------------------------------------------------------------------------------------------------------------
// RCC*::find (RCCEmbree.cpp:26-36 ...): one thread per (pose, ray); writes the model buffers in the SENSOR frame
// ---------------------------------------------------------------------------------------------------------------------
struct RayModel {
    const float* dirs;      // n x 3 (sensor frame)
    const float* origs;     // n_origs x 3
    uint32_t n_origs;       // 1 or n
    uint32_t n;             // rays per pose
    float range_min, range_max;
    uint32_t width, height; // scan raster (buffer id = vid*width + hid); used for the coherent tile order
    uint32_t sim_opts;      // b2_rcc_set_sim_options: bit 0 tfar = +inf instead of range.max; bit 1 a closest hit below range.min is a miss; bit 2 misses filled with zeros instead of NaN
};
B2_DEV float sim_tfar(const RayModel& m) { return (m.sim_opts & 1u) ? u2f(0x7f800000u) : m.range_max; }
B2_DEV bool sim_is_hit(const RayModel& m, const HitRec& h) { return h.face != B2_NOFACE && !((m.sim_opts & 2u) && h.t < m.range_min); }

// Rays are traced in 8x4 raster tiles (one tile per warp) instead of 32-long row segments: neighbouring rows/columns of a
// LiDAR / depth raster stay inside the same BVH subtrees, which shortens the warp's union of traversal paths.  Results are
// written at the original buffer id, so the order is invisible outside.
B2_DEV uint32_t tile_order(uint32_t k, uint32_t width, uint32_t height)
{
    if ((width & 7u) || (height & 3u)) return k;
    const uint32_t tile = k >> 5, within = k & 31u, tiles_per_row = width >> 3;
    const uint32_t hid = (tile % tiles_per_row) * 8u + (within & 7u);
    const uint32_t vid = (tile / tiles_per_row) * 4u + (within >> 3);
    return vid * width + hid;
}

struct ModelBuffers {
    float* pts; float* nrm; uint8_t* hits; uint32_t* faces; float* ranges;
};

// shared epilogue: hit record -> sensor-frame point / normal exactly like the oracle's orc_simulate
B2_DEV void hit_to_sensor(const BvhView& bvh, const HitRec& h, Q4 Rms, V3 dir_s, V3 orig_s, V3& p, V3& ns)
{
    const V3 ng = tri_ng(bvh, h.tri);
    p = v_add(v_scale(dir_s, h.t), orig_s);
    const V3 nm = v_normalize(ng);
    ns = q_rot(Rms, nm);
    if (v_dot(dir_s, ns) > 0.0f) ns = v_neg(ns);
    ns = v_normalize(ns);
}

// one ray of find(): trace + write the model buffers at index o
B2_DEV void find_one(const BvhView& bvh, Tf Tsm, const RayModel& model, uint32_t i, uint64_t o, const ModelBuffers& out)
{
    const Q4 Rms = q_conj(Tsm.R);
    const uint32_t oi = model.n_origs == 1 ? 0 : i;
    const V3 orig_s = mk3(model.origs[3 * oi], model.origs[3 * oi + 1], model.origs[3 * oi + 2]);
    const V3 dir_s = mk3(model.dirs[3 * i], model.dirs[3 * i + 1], model.dirs[3 * i + 2]);
    const RaySetup r = ray_setup(tf_apply(Tsm, orig_s), q_rot(Tsm.R, dir_s), bvh);
    HitRec h = trace_init(sim_tfar(model));
    uint32_t nn = 0, nt = 0;
    trace_closest<false>(bvh, r, h, nn, nt);
    if (sim_is_hit(model, h)) {
        V3 p, ns; hit_to_sensor(bvh, h, Rms, dir_s, orig_s, p, ns);
        out.pts[3 * o] = p.x; out.pts[3 * o + 1] = p.y; out.pts[3 * o + 2] = p.z;
        out.nrm[3 * o] = ns.x; out.nrm[3 * o + 1] = ns.y; out.nrm[3 * o + 2] = ns.z;
        out.hits[o] = 1; out.faces[o] = h.face; out.ranges[o] = h.t;
    } else {
        const float qnan = (model.sim_opts & 4u) ? 0.0f : u2f(0x7fc00000u);
        out.pts[3 * o] = qnan; out.pts[3 * o + 1] = qnan; out.pts[3 * o + 2] = qnan;
        out.nrm[3 * o] = qnan; out.nrm[3 * o + 1] = qnan; out.nrm[3 * o + 2] = qnan;
        out.hits[o] = 0; out.faces[o] = B2_NOFACE; out.ranges[o] = add(model.range_max, 1.0f);
    }
}

// bulk prefetch of this block's slice of the node array into L2 (UBLKPF): after a cold start the top of the tree then comes from
// L2 instead of one DRAM round trip per level.  16-byte granularity, fire and forget.
__device__ __forceinline__ void bulk_prefetch_slice(const void* base, uint64_t total)
{
    const uint64_t per = ((total + gridDim.x - 1) / gridDim.x + 15ull) & ~15ull;
    const uint64_t beg = (uint64_t)blockIdx.x * per;
    if (beg < total) {
        const uint32_t bytes = (uint32_t)min(per, total - beg) & ~15u;
        const char* ptr = reinterpret_cast<const char*>(base) + beg;
        if (bytes) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(ptr), "r"(bytes) : "memory");
    }
}
// mode 1: node array; mode 2: node array + leaf triangle records (the whole map, 56 MB for 1M triangles, fits the 126 MB L2)
__device__ __forceinline__ void prefetch_map_l2(const BvhView& bvh, uint32_t n_nodes, uint32_t n_tris, int mode)
{
    if (threadIdx.x == 0 && mode >= 1) bulk_prefetch_slice(bvh.nodes, (uint64_t)n_nodes * (uint64_t)B2_NODE_BYTES);
    if (threadIdx.x == 32 % blockDim.x && mode >= 2) bulk_prefetch_slice(bvh.tris, (uint64_t)n_tris * 48ull);
}

#define B2_FIND_BLOCK 64
// profiling aid: when set (b2_rcc_debug_find_warp_times), lane 0 of every warp stores {start, end} in %globaltimer nanoseconds
__device__ unsigned long long* g_find_warp_times = nullptr;
__device__ __forceinline__ unsigned long long globaltimer_ns() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }

// 14 blocks per SM (72 registers): the 2048 blocks of the C2 scan are resident in ONE wave on 148 SMs (13.84 per SM); with the 76
// registers ptxas picks on its own only 13 fit and the last blocks wait for a slot (measured 54.0 -> 51.7 us).  A variant with 16 rays
// per warp (half-empty warps, shorter max-over-lanes trip count) was measured slower (71 us: the idle lanes still own registers).
// Tile schedule (single-pose launches): the block scheduler hands out blocks in index order over a few microseconds, and the kernel ends
// with its slowest warp -- so the tiles that took longest in the PREVIOUS launch of this handle (same model, nearly the same pose: MICP-L
// corrects continuously) go first.  Each warp leaves its duration in `tile_cost`; the ICP loop that follows the find turns the durations
// into the order `tile_perm` of the next launch (icp_loop.cuh: tile_perm_group, idle warps of one block).  Purely a schedule: every tile is traced exactly once, results do not depend on it.
__global__ void __launch_bounds__(B2_FIND_BLOCK, 14) k_rcc_find(BvhView bvh, uint32_t n_nodes, uint32_t n_tris, int prefetch_mode, const b2_transform* __restrict__ Tbm_dev,
                                                                const IcpState* __restrict__ icp, b2_transform Tbm_val, b2_transform Tsb_val, RayModel model, uint32_t n_poses,
                                                                ModelBuffers out, int early_dependents, const uint16_t* __restrict__ tile_perm, uint32_t* __restrict__ tile_cost)
{
    __shared__ uint32_t s_t0[B2_FIND_BLOCK / 32];
    uint64_t gid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tile_perm) gid = ((uint64_t)tile_perm[gid >> 5] << 5) | (gid & 31u);                      // n_poses == 1; a permutation of the tiles
    if (tile_cost && (threadIdx.x & 31u) == 0u) { uint32_t c; asm volatile("mov.u32 %0, %%clock;" : "=r"(c)); s_t0[threadIdx.x >> 5] = c; }
    // let a dependent kernel launched with programmatic stream serialization (k_icp_loop) become resident as SMs drain; it still waits
    // (griddepcontrol.wait) for this grid to complete and flush before it reads the model buffers
    if (early_dependents) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (g_find_warp_times && (threadIdx.x & 31u) == 0u) g_find_warp_times[2 * (((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5)] = globaltimer_ns();      // nothing stays live across the trace
    prefetch_map_l2(bvh, n_nodes, n_tris, prefetch_mode);
    const uint64_t total = (uint64_t)model.n * n_poses;
    if (gid < total) {
        const uint32_t pose = (uint32_t)(gid / model.n), i = tile_order((uint32_t)(gid % model.n), model.width, model.height);
        Tf Tbm;
        if (icp) Tbm = tf_mul(tf_load(&icp->Tom), tf_load(&icp->Tbo));           // MICPSensor.hpp:148
        else if (Tbm_dev) Tbm = tf_load(Tbm_dev + pose);
        else Tbm = tf_from_pod(Tbm_val);
        find_one(bvh, tf_mul(Tbm, tf_from_pod(Tsb_val)), model, i, (uint64_t)pose * model.n + i, out);
    }
    if (tile_cost) {
        __syncwarp();
        if ((threadIdx.x & 31u) == 0u) { uint32_t c; asm volatile("mov.u32 %0, %%clock;" : "=r"(c)); tile_cost[gid >> 5] = c - s_t0[threadIdx.x >> 5]; }
    }
    if (g_find_warp_times) {
        __syncwarp();
        if ((threadIdx.x & 31u) == 0u) g_find_warp_times[2 * (((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5) + 1] = globaltimer_ns();
    }
}

#ifdef __CUDACC__
// the tile order before any durations are known; also what a launch falls back to when the loop that should have sorted never ran
__global__ void k_perm_identity(uint16_t* __restrict__ perm, uint32_t n)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) perm[i] = (uint16_t)i;
}
#endif

// ---------------------------------------------------------------------------------------------------------------------
// CPCEmbree::find (CPCEmbree.cpp:17-43): one thread per dataset point (the mask is NOT consulted, as in the reference);
//   Pm = Tsm * d_i;  cp = closestPoint(Pm);  hits = cp.d <= max_dist;  points = Tms * cp.p;  normals = Tms.R * cp.n
// The model buffers get one entry per dataset point, so computeCrossStatistics / the ICP loop run unchanged on them.
// ---------------------------------------------------------------------------------------------------------------------
B2_DEV void cpc_find_one(const BvhView& bvh, Tf Tsm, Tf Tms, const float* __restrict__ dpts, float max_dist, uint32_t i, const ModelBuffers& out)
{
    const V3 q = tf_apply(Tsm, mk3(dpts[3 * i], dpts[3 * i + 1], dpts[3 * i + 2]));
    CpBest best; uint32_t nn = 0, nt = 0;
    closest_point<false>(bvh, q, best, nn, nt);
    if (best.face != B2_NOFACE) {
        const float d = sqrtf(best.d2);
        const V3 p = tf_apply(Tms, best.p);
        const V3 ns = q_rot(Tms.R, v_normalize(tri_ng(bvh, best.tri)));
        out.pts[3 * i] = p.x; out.pts[3 * i + 1] = p.y; out.pts[3 * i + 2] = p.z;
        out.nrm[3 * i] = ns.x; out.nrm[3 * i + 1] = ns.y; out.nrm[3 * i + 2] = ns.z;
        out.hits[i] = d <= max_dist ? 1 : 0; out.faces[i] = best.face; out.ranges[i] = d;
    } else {
        const float qnan = u2f(0x7fc00000u);
        out.pts[3 * i] = qnan; out.pts[3 * i + 1] = qnan; out.pts[3 * i + 2] = qnan;
        out.nrm[3 * i] = qnan; out.nrm[3 * i + 1] = qnan; out.nrm[3 * i + 2] = qnan;
        out.hits[i] = 0; out.faces[i] = B2_NOFACE; out.ranges[i] = u2f(0x7f800000u);
    }
}

#ifdef __CUDACC__
__global__ void __launch_bounds__(B2_FIND_BLOCK) k_cpc_find(BvhView bvh, uint32_t n_nodes, uint32_t n_tris, int prefetch_mode, const IcpState* __restrict__ icp,
                                                            b2_transform Tbm_val, b2_transform Tsb_val, const float* __restrict__ dpts, uint32_t n, float max_dist,
                                                            ModelBuffers out, const uint8_t* __restrict__ skip_mask)
{
    prefetch_map_l2(bvh, n_nodes, n_tris, prefetch_mode);
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (skip_mask && !(skip_mask[i] > 0)) {
        // b2_rcc_set_cpc_options(skip_masked): a masked-out dataset point (e.g. a dropped beam unpacked to range.max + 1, far outside the map) never
        // enters the statistics; its query is the most expensive of the scan (2 % of the points cost a third of the kernel: one such lane keeps
        // its whole warp) and is skipped -- the entry reads hits = 0, NaN, like a miss.  Off by default: the reference queries every point.
        const float qnan = u2f(0x7fc00000u);
        out.pts[3 * i] = qnan; out.pts[3 * i + 1] = qnan; out.pts[3 * i + 2] = qnan;
        out.nrm[3 * i] = qnan; out.nrm[3 * i + 1] = qnan; out.nrm[3 * i + 2] = qnan;
        out.hits[i] = 0; out.faces[i] = B2_NOFACE; out.ranges[i] = u2f(0x7f800000u);
        return;
    }
    const Tf Tbm = icp ? tf_mul(tf_load(&icp->Tom), tf_load(&icp->Tbo)) : tf_from_pod(Tbm_val);
    const Tf Tsm = tf_mul(Tbm, tf_from_pod(Tsb_val));
    cpc_find_one(bvh, Tsm, tf_inv(Tsm), dpts, max_dist, i, out);
}
#endif

// ---------------------------------------------------------------------------------------------------------------------
// refit of ONE node of the wide tree (b2_mesh_refit, lbvh.cuh:k_bvh8_refit_level): leaf children re-fetch their triangles, inner children
// take the union of the (already refitted) child node's boxes
// ---------------------------------------------------------------------------------------------------------------------
B2_DEV void bvh8_refit_node(uint32_t t, B2Node8* nodes, B2Tri* tris, const float* verts, const uint32_t* faces)
{
    B2Node8& nd = nodes[t];
    const float inf = u2f(0x7f800000u);
    for (int s = 0; s < 8; s++) {
        const uint32_t meta = nd.meta[s];
        float lo[3] = {inf, inf, inf}, hi[3] = {-inf, -inf, -inf};
        if ((nd.imask() >> s) & 1u) {
            const B2Node8& ch = nodes[nd.child_base + popc32(nd.imask() & ((1u << s) - 1u))];
            for (int c = 0; c < 8; c++) {
                if (!ch.meta[c]) continue;
                for (int k = 0; k < 3; k++) { lo[k] = fminf(lo[k], ch.lo[k][c]); hi[k] = fmaxf(hi[k], ch.hi[k][c]); }
            }
        } else if (meta) {
            const uint32_t cnt = (meta >> 5) == 7u ? 3u : ((meta >> 5) == 3u ? 2u : 1u), first = nd.tri_base + (meta & 0x1fu);
            for (uint32_t j = 0; j < cnt; j++) {
                B2Tri& tr = tris[first + j];
                const uint32_t f = tr.face_id;
                const float* a = verts + 3 * (size_t)faces[3 * (size_t)f + 0];
                const float* b = verts + 3 * (size_t)faces[3 * (size_t)f + 1];
                const float* c = verts + 3 * (size_t)faces[3 * (size_t)f + 2];
                for (int k = 0; k < 3; k++) {
                    tr.v0[k] = a[k]; tr.v1[k] = b[k]; tr.v2[k] = c[k];
                    lo[k] = fminf(lo[k], fminf(fminf(a[k], b[k]), c[k])); hi[k] = fmaxf(hi[k], fmaxf(fmaxf(a[k], b[k]), c[k]));
                }
            }
        }
        for (int k = 0; k < 3; k++) { nd.lo[k][s] = lo[k]; nd.hi[k][s] = hi[k]; }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// scan-vs-map segmentation: classification loop of ScanMapSegmentationEmbreeNode
// (rmcl_ros/src/nodes/filter/scan_map_segmentation_embree.cpp:110-187) on the model buffers left by find() and the real ranges.
// label 0: neither cloud, 1: outlier_scan (point = preal), 2: outlier_map (point = pint)
// ---------------------------------------------------------------------------------------------------------------------
B2_DEV uint32_t segment_classify(const RayModel& model, uint32_t i, float rr, float rs, V3 nsim, float min_scan, float min_map, V3& p)
{
    const uint32_t oi = model.n_origs == 1 ? 0 : i;
    const V3 dir = mk3(model.dirs[3 * i], model.dirs[3 * i + 1], model.dirs[3 * i + 2]);
    const V3 orig = mk3(model.origs[3 * oi], model.origs[3 * oi + 1], model.origs[3 * oi + 2]);
    const bool real_valid = (model.range_min <= rr) && (rr <= model.range_max);           // model.range.inside, :121-122
    const bool sim_valid = (model.range_min <= rs) && (rs <= model.range_max);
    p = mk3(0.f, 0.f, 0.f);
    if (real_valid) {
        const V3 preal = v_add(v_scale(dir, rr), orig);                                   // :126
        if (!sim_valid) { p = preal; return 1u; }                                         // :164-171
        const V3 pint = v_scale(dir, rs);                                                 // :130 (no origin, as in the reference)
        const V3 nint = v_normalize(nsim);                                                // :131-132
        const float spd = v_dot(v_sub(preal, pint), nint);                                // :134
        const V3 pmesh = v_add(preal, v_scale(nint, spd));                                // :135
        const float plane_distance = v_l2norm(v_sub(pmesh, preal));                       // :136
        if (rr < rs) { if (plane_distance > min_scan) { p = preal; return 1u; } }         // :138-149
        else         { if (plane_distance > min_map)  { p = pint;  return 2u; } }         // :150-161
        return 0u;
    }
    if (sim_valid) { p = v_add(v_scale(dir, rs), orig); return 2u; }                      // :173-182
    return 0u;
}

#ifdef __CUDACC__
#define B2_SEG_BLOCK 256
// pass 1 (counts != nullptr, out_* == nullptr): per-block counts of both classes.  pass 3: recompute and write each outlier at
// block offset + its rank inside the block, so both clouds come out in raster order like the reference's push_back loop.
__global__ void __launch_bounds__(B2_SEG_BLOCK) k_segment(RayModel model, const float* __restrict__ ranges_real, const float* __restrict__ ranges_sim,
                                                          const float* __restrict__ normals_sim, float min_scan, float min_map, uint32_t* __restrict__ counts,
                                                          const uint32_t* __restrict__ offsets, float* __restrict__ out_scan, float* __restrict__ out_map, uint8_t* __restrict__ labels)
{
    __shared__ uint32_t s_cnt[2][B2_SEG_BLOCK / 32];
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x, lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint32_t label = 0u; V3 p = mk3(0.f, 0.f, 0.f);
    if (i < model.n) label = segment_classify(model, i, ranges_real[i], ranges_sim[i], mk3(normals_sim[3 * i], normals_sim[3 * i + 1], normals_sim[3 * i + 2]), min_scan, min_map, p);
    const uint32_t b1 = __ballot_sync(0xffffffffu, label == 1u), b2 = __ballot_sync(0xffffffffu, label == 2u);
    if (lane == 0) { s_cnt[0][warp] = __popc(b1); s_cnt[1][warp] = __popc(b2); }
    __syncthreads();
    uint32_t before1 = 0u, before2 = 0u, tot1 = 0u, tot2 = 0u;
    #pragma unroll
    for (uint32_t w = 0; w < B2_SEG_BLOCK / 32; w++) { if (w < warp) { before1 += s_cnt[0][w]; before2 += s_cnt[1][w]; } tot1 += s_cnt[0][w]; tot2 += s_cnt[1][w]; }
    if (!out_scan) {
        if (threadIdx.x == 0) { counts[2 * blockIdx.x] = tot1; counts[2 * blockIdx.x + 1] = tot2; }
        return;
    }
    const uint32_t lt = (1u << lane) - 1u;
    if (label == 1u) { const size_t o = (size_t)offsets[2 * blockIdx.x] + before1 + __popc(b1 & lt); out_scan[3 * o] = p.x; out_scan[3 * o + 1] = p.y; out_scan[3 * o + 2] = p.z; }
    if (label == 2u) { const size_t o = (size_t)offsets[2 * blockIdx.x + 1] + before2 + __popc(b2 & lt); out_map[3 * o] = p.x; out_map[3 * o + 1] = p.y; out_map[3 * o + 2] = p.z; }
    if (labels && i < model.n) labels[i] = (uint8_t)label;
}
// pass 2: exclusive scan of the per-block counts (one block; a few thousand entries at most), totals to totals[0..1]
__global__ void __launch_bounds__(1024) k_segment_scan(const uint32_t* __restrict__ counts, uint32_t n_blocks, uint32_t* __restrict__ offsets, uint32_t* __restrict__ totals)
{
    __shared__ uint32_t s[2][1024];
    __shared__ uint32_t carry[2];
    if (threadIdx.x == 0) { carry[0] = 0u; carry[1] = 0u; }
    __syncthreads();
    for (uint32_t base = 0; base < n_blocks; base += 1024u) {
        const uint32_t b = base + threadIdx.x;
        const uint32_t v0 = b < n_blocks ? counts[2 * b] : 0u, v1 = b < n_blocks ? counts[2 * b + 1] : 0u;
        s[0][threadIdx.x] = v0; s[1][threadIdx.x] = v1;
        __syncthreads();
        for (uint32_t d = 1; d < 1024u; d <<= 1) {                      // Hillis-Steele inclusive scan
            const uint32_t a0 = threadIdx.x >= d ? s[0][threadIdx.x - d] : 0u, a1 = threadIdx.x >= d ? s[1][threadIdx.x - d] : 0u;
            __syncthreads();
            s[0][threadIdx.x] += a0; s[1][threadIdx.x] += a1;
            __syncthreads();
        }
        if (b < n_blocks) { offsets[2 * b] = carry[0] + s[0][threadIdx.x] - v0; offsets[2 * b + 1] = carry[1] + s[1][threadIdx.x] - v1; }
        __syncthreads();
        if (threadIdx.x == 1023) { carry[0] += s[0][1023]; carry[1] += s[1][1023]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { totals[0] = carry[0]; totals[1] = carry[1]; }
}
#endif

// ---------------------------------------------------------------------------------------------------------------------
// Correspondences*::computeCrossStatistics (CorrespondencesCPU.cpp:10-39): N-element masked reduction -> CrossStatistics.
// Deterministic: per-block partials, the last block to finish sums them in block order.  With `icp` set, the last block also
// performs the rest of the inner iteration (frame changes, Umeyama, compose) so one correctOnce needs 1 + iterations launches.
// ---------------------------------------------------------------------------------------------------------------------
#define B2_RED_BLOCK 256
__global__ void __launch_bounds__(B2_RED_BLOCK) k_p2l_reduce(const float* __restrict__ dpts, const uint8_t* __restrict__ dmask,
                                                             const float* __restrict__ mpts, const float* __restrict__ mnrm, const uint8_t* __restrict__ mmask,
                                                             uint32_t n, b2_transform Tpre_val, float max_dist_val, IcpState* icp,
                                                             double* __restrict__ partials, unsigned int* __restrict__ ticket, b2_cross_stats* __restrict__ out)
{
    __shared__ double smem[(B2_NACC + 1) * (B2_RED_BLOCK / 32)];
    __shared__ bool is_last;
    const long long c0 = clock64();
    const Tf Tpre = icp ? tf_load(&icp->T_snew_sold) : tf_from_pod(Tpre_val);
    const float max_dist = icp ? icp->max_dist : max_dist_val;
    P2LAcc acc; acc_zero(acc);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        // all eleven loads are issued before the first use (no mask-dependent early-out): one memory round trip per element
        const uint8_t dm = dmask[i], mm = mmask[i];
        const V3 d = mk3(dpts[3 * i], dpts[3 * i + 1], dpts[3 * i + 2]);
        const V3 I = mk3(mpts[3 * i], mpts[3 * i + 1], mpts[3 * i + 2]);
        const V3 N = mk3(mnrm[3 * i], mnrm[3 * i + 1], mnrm[3 * i + 2]);
        V3 D, M;
        if ((dm > 0) && (mm > 0) && p2l_pair(Tpre, d, I, N, max_dist, D, M)) acc_add_pair(acc, D, M);
    }
    block_reduce_acc<B2_RED_BLOCK>(acc, smem);
    if (threadIdx.x == 0) {
        double* p = partials + (size_t)blockIdx.x * (B2_NACC + 1);
        for (int i = 0; i < B2_NACC; i++) p[i] = acc.v[i];
        p[B2_NACC] = (double)acc.n;
        __threadfence();
        const unsigned int t = atomicAdd(ticket, 1u);
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    const long long c1 = clock64();
    // deterministic parallel sum of the block partials: thread (g, i) adds value i of blocks g, g+16, ...; then the 16 group sums in order
    __shared__ double s_part[16][B2_NACC + 1];
    {
        const uint32_t i = threadIdx.x & 15u, g = threadIdx.x >> 4;           // 256 threads = 16 groups x 16 values
        double a = 0.0;
        for (uint32_t b = g; b < gridDim.x; b += 16u) a += __ldcg(partials + (size_t)b * (B2_NACC + 1) + i);
        s_part[g][i] = a;
    }
    __syncthreads();
    if (threadIdx.x < B2_NACC + 1) {
        double a = 0.0;
        #pragma unroll
        for (int g = 0; g < 16; g++) a += s_part[g][threadIdx.x];
        s_part[0][threadIdx.x] = a;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const long long c2 = clock64();
        const CStats st = acc_finalize(&s_part[0][0], (uint32_t)(s_part[0][B2_NACC] + 0.5));
        if (out) cs_store(out, st);
        const long long c3 = clock64();
        if (icp) {
            icp_step(icp, st);
            const long long c4 = clock64();
            icp->dbg[0] = (unsigned long long)(c1 - c0); icp->dbg[1] = (unsigned long long)(c2 - c1); icp->dbg[2] = (unsigned long long)(c3 - c2);
            icp->dbg[3] = (unsigned long long)(c4 - c3);
        }
        *ticket = 0u;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// v1 batched correct(): fused trace -> P2L gate -> per-block partial statistics (no model buffers written)
// grid = n_poses * blocks_per_pose; each block handles `rays_per_block` consecutive rays of one pose
// ---------------------------------------------------------------------------------------------------------------------
#define B2_FUSED_BLOCK 128
// 6 blocks per SM (80 registers, a few spilled values): 2.5 % faster than the uncapped 128 registers / 4 blocks (3.06 -> 2.98 ms per 1000-pose call)
__global__ void __launch_bounds__(B2_FUSED_BLOCK, 6) k_rcc_fused_batch(BvhView bvh, const b2_transform* __restrict__ Tbm_dev, b2_transform Tsb_val, RayModel model,
                                                                    const float* __restrict__ dpts, const uint8_t* __restrict__ dmask, float max_dist,
                                                                    uint32_t blocks_per_pose, uint32_t rays_per_block, double* __restrict__ partials)
{
    __shared__ double smem[(B2_NACC + 1) * (B2_FUSED_BLOCK / 32)];
    const uint32_t pose = blockIdx.x / blocks_per_pose, chunk = blockIdx.x % blocks_per_pose;
    const Tf Tsm = tf_mul(tf_load(Tbm_dev + pose), tf_from_pod(Tsb_val));
    const Q4 Rms = q_conj(Tsm.R);
    const Tf I = tf_identity();
    P2LAcc acc; acc_zero(acc);
    const uint32_t begin = chunk * rays_per_block;
    const uint32_t end = min(begin + rays_per_block, model.n);
    for (uint32_t k = begin + threadIdx.x; k < end; k += blockDim.x) {
        const uint32_t i = tile_order(k, model.width, model.height);
        if (!(dmask[i] > 0)) continue;
        const uint32_t oi = model.n_origs == 1 ? 0 : i;
        const V3 orig_s = mk3(model.origs[3 * oi], model.origs[3 * oi + 1], model.origs[3 * oi + 2]);
        const V3 dir_s = mk3(model.dirs[3 * i], model.dirs[3 * i + 1], model.dirs[3 * i + 2]);
        const RaySetup r = ray_setup(tf_apply(Tsm, orig_s), q_rot(Tsm.R, dir_s), bvh);
        HitRec h = trace_init(sim_tfar(model));
        uint32_t nn = 0, nt = 0;
        trace_closest<false>(bvh, r, h, nn, nt);
        if (!sim_is_hit(model, h)) continue;
        V3 p, ns; hit_to_sensor(bvh, h, Rms, dir_s, orig_s, p, ns);
        V3 D, M;
        if (p2l_pair(I, mk3(dpts[3 * i], dpts[3 * i + 1], dpts[3 * i + 2]), p, ns, max_dist, D, M)) acc_add_pair(acc, D, M);
    }
    block_reduce_acc<B2_FUSED_BLOCK>(acc, smem);
    if (threadIdx.x == 0) {
        double* p = partials + (size_t)blockIdx.x * (B2_NACC + 1);
        for (int i = 0; i < B2_NACC; i++) p[i] = acc.v[i];
        p[B2_NACC] = (double)acc.n;
    }
}

// stage-split twin of k_rcc_fused_batch for the v1 benchmark() call ({sim, red, svd} seconds, lidar_corrector_optix_benchmark.cpp:143-155):
// the reduction alone over model buffers that k_rcc_find wrote for ALL poses (pose-major), same partial layout as the fused kernel
__global__ void __launch_bounds__(B2_FUSED_BLOCK) k_p2l_batch(const float* __restrict__ mpts, const float* __restrict__ mnrm, const uint8_t* __restrict__ mhits, uint32_t n,
                                                             const float* __restrict__ dpts, const uint8_t* __restrict__ dmask, float max_dist,
                                                             uint32_t blocks_per_pose, uint32_t rays_per_block, double* __restrict__ partials)
{
    __shared__ double smem[(B2_NACC + 1) * (B2_FUSED_BLOCK / 32)];
    const uint32_t pose = blockIdx.x / blocks_per_pose, chunk = blockIdx.x % blocks_per_pose;
    const Tf I = tf_identity();
    P2LAcc acc; acc_zero(acc);
    const uint32_t begin = chunk * rays_per_block, end = min(begin + rays_per_block, n);
    for (uint32_t i = begin + threadIdx.x; i < end; i += blockDim.x) {
        const size_t o = (size_t)pose * n + i;
        if (!(dmask[i] > 0) || !(mhits[o] > 0)) continue;
        V3 D, M;
        if (p2l_pair(I, mk3(dpts[3 * i], dpts[3 * i + 1], dpts[3 * i + 2]), mk3(mpts[3 * o], mpts[3 * o + 1], mpts[3 * o + 2]), mk3(mnrm[3 * o], mnrm[3 * o + 1], mnrm[3 * o + 2]), max_dist, D, M))
            acc_add_pair(acc, D, M);
    }
    block_reduce_acc<B2_FUSED_BLOCK>(acc, smem);
    if (threadIdx.x == 0) {
        double* p = partials + (size_t)blockIdx.x * (B2_NACC + 1);
        for (int i = 0; i < B2_NACC; i++) p[i] = acc.v[i];
        p[B2_NACC] = (double)acc.n;
    }
}

// per pose: sum the block partials in order -> stats_s -> stats_b = Tsb * stats_s -> Umeyama
__global__ void k_umeyama_from_partials(const double* __restrict__ partials, uint32_t blocks_per_pose, uint32_t n_poses, b2_transform Tsb_val,
                                        b2_transform* __restrict__ Tdelta, uint32_t* __restrict__ ncorr, b2_cross_stats* __restrict__ stats_b_out)
{
    const uint32_t pose = blockIdx.x * blockDim.x + threadIdx.x;
    if (pose >= n_poses) return;
    double v[B2_NACC]; for (int i = 0; i < B2_NACC; i++) v[i] = 0.0;
    double cnt = 0.0;
    for (uint32_t b = 0; b < blocks_per_pose; b++) {
        const double* p = partials + ((size_t)pose * blocks_per_pose + b) * (B2_NACC + 1);
        for (int i = 0; i < B2_NACC; i++) v[i] += p[i];
        cnt += p[B2_NACC];
    }
    const CStats ss = acc_finalize(v, (uint32_t)(cnt + 0.5));
    const CStats sb = cs_transform(tf_from_pod(Tsb_val), ss);
    const Tf T = umeyama_dev(sb);
    if (Tdelta) tf_store(Tdelta + pose, T);
    if (ncorr) ncorr[pose] = sb.n;
    if (stats_b_out) cs_store(stats_b_out + pose, sb);
}

__global__ void k_umeyama_batch(const b2_cross_stats* __restrict__ stats, uint32_t n, b2_transform* __restrict__ out)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    tf_store(out + i, umeyama_dev(cs_load(stats + i)));
}

// ---------------------------------------------------------------------------------------------------------------------
// particle filter: PCDSensorUpdaterEmbree::update hot loop (PCDSensorUpdaterEmbree.cpp:290-342) in ONE launch.
// A block owns PPB particles x all beams.  Phase 1: every (particle, beam) ray is traced by some thread, its Gaussian
// evaluation stored in shared memory.  Phase 2: one thread per particle merges the evaluations IN BEAM ORDER
// (the reference's sequential FP32 Gaussian1D += order, quirk D6) and read-modify-writes the 36-byte attrs once.
// ---------------------------------------------------------------------------------------------------------------------
struct PfBeam { float ox, oy, oz, dx, dy, dz, range; uint32_t slot; };   // slot = original beam index (merge order)

B2_DEV void gaussian1d_add(b2_gaussian1d& a, float b_mean, float b_sigma, uint32_t b_n)
{
    const uint32_t n = a.n_meas + b_n;
    if (n == 0) return;
    const float w1 = dvd((float)a.n_meas, (float)n), w2 = dvd((float)b_n, (float)n);
    const float mean = add(mul(a.mean, w1), mul(b_mean, w2));
    const float d1 = sub(a.mean, mean), d2 = sub(b_mean, mean);
    const float sigma = add(add(mul(a.sigma, w1), mul(b_sigma, w2)), add(mul(mul(d1, d1), w1), mul(mul(d2, d2), w2)));
    a.mean = mean; a.sigma = sigma; a.n_meas = n;
}

// Gaussian evaluation of one (particle, beam) pair: sensorUpdate() up to `eval` (PCDSensorUpdaterEmbree.cpp:197-224)
template <int CORR>
B2_DEV float pf_eval_one(const BvhView& bvh, Tf Tsm, const PfBeam& b, const b2_pf_params& prm, float sigma_quad, double denom)
{
    const V3 orig_m = tf_apply(Tsm, mk3(b.ox, b.oy, b.oz));                                // RangeMeasurement.hpp:29-42
    const V3 dir_m = q_rot(Tsm.R, mk3(b.dx, b.dy, b.dz));
    if (CORR == 1) {
        // evaluate_cpc (:88-95): error = distance of meas_m.mean() = orig + dir*range (RangeMeasurement.hpp:17-20) to the surface
        CpBest best; uint32_t nn = 0, nt = 0;
        closest_point<false>(bvh, v_add(orig_m, v_scale(dir_m, b.range)), best, nn, nt);
        const float error = best.face != B2_NOFACE ? sqrtf(best.d2) : u2f(0x7f800000u);
        const float arg = dvd(dvd(-mul(error, error), sigma_quad), 2.0f);                  // :224
        return (float)(exp((double)arg) / denom);
    }
    const bool real_hit = (prm.range_min <= b.range) && (b.range <= prm.range_max);        // :27
    const RaySetup r = ray_setup(orig_m, dir_m, bvh);
    HitRec h = trace_init(u2f(0x7f800000u));                                               // tfar = +inf (:38)
    uint32_t nn = 0, nt = 0;
    trace_closest<false>(bvh, r, h, nn, nt);
    const bool sim_hit = (h.face != B2_NOFACE) && (h.t > prm.range_min);                   // :47
    float error;
    if (sim_hit) {
        if (real_hit) {
            V3 n = tri_ng(bvh, h.tri);
            if (prm.ng_mode == 1) n = v_normalize(n);
            const V3 preal = v_add(orig_m, v_scale(dir_m, b.range));
            const V3 pint = v_add(orig_m, v_scale(dir_m, h.t));
            error = fabsf(v_dot(v_sub(pint, preal), n));                                   // :54-68
        } else error = prm.real_miss_sim_hit_error;
    } else error = real_hit ? prm.real_hit_sim_miss_error : prm.real_miss_sim_miss_error;
    const float arg = dvd(dvd(-mul(error, error), sigma_quad), 2.0f);                      // :224
    return (float)(exp((double)arg) / denom);
}
B2_DEV void pf_constants(const b2_pf_params& prm, float& sigma_quad, double& denom)
{
    sigma_quad = mul(prm.dist_sigma, prm.dist_sigma);
    denom = sqrt((double)mul(2.0f, sigma_quad) * 3.14159265358979323846);                  // sqrt(2*sq*M_PI)
}
// merge the evaluations of one particle in beam order (:232-238)
B2_DEV void pf_merge(b2_gaussian1d& lk, const float* e, uint32_t n_beams)
{
    for (uint32_t b = 0; b < n_beams; b++) {
        gaussian1d_add(lk, e[b], 0.0f, 1u);
        lk.n_meas = lk.n_meas < 10000u ? lk.n_meas : 10000u;                               // MAX_N_MEAS
    }
}

#define B2_PF_BLOCK 128
// 72 registers (7 blocks/SM) measured 8 % faster than the uncapped 84 (6 blocks/SM); a persistent-lane variant with dynamic ray fetch
// (idle lanes claim new rays) was measured SLOWER (2.3 vs 3.3 G rays/s: the extra live state and warp votes cost more than the refill gains)
// MAP 0: consecutive lanes = consecutive beams of one particle (rays of a warp share their origin: the mapping for particle sets spread over
// the map, where no two particles are close).  MAP 1: consecutive lanes = the particles of the block, same beam -- coherent when neighbouring
// particles have nearly the same pose, i.e. for a converged cloud after `order` sorted it by (heading, cell): 1.7x faster there, slower on
// spread-out sets (scripts/exp_pf_mapping.py); the host picks by timing both (api.cu).  `order` (may be nullptr) = particle processed at
// position p.  The per-particle merge runs over the beams in their original order either way: results are bit-identical.
template <int CORR, int MAP>
__global__ void __launch_bounds__(B2_PF_BLOCK, 7) k_pf_update(BvhView bvh, const b2_transform* __restrict__ poses, b2_particle_attr* __restrict__ attrs, uint32_t n_particles,
                                                           b2_transform Tsb_val, const PfBeam* __restrict__ beams, uint32_t n_beams, b2_pf_params prm, uint32_t ppb,
                                                           const uint32_t* __restrict__ order)
{
    extern __shared__ float s_eval[];            // [ppb][n_beams]
    const uint32_t p0 = blockIdx.x * ppb;
    const uint32_t np = min(ppb, n_particles - p0);
    float sigma_quad; double denom; pf_constants(prm, sigma_quad, denom);
    const Tf Tsb = tf_from_pod(Tsb_val);
    const uint32_t total = np * n_beams;
    for (uint32_t w = threadIdx.x; w < total; w += blockDim.x) {
        const uint32_t pl = MAP ? w % np : w / n_beams, bi = MAP ? w / np : w % n_beams;
        const uint32_t pi = order ? order[p0 + pl] : p0 + pl;
        const Tf Tsm = tf_mul(tf_load(poses + pi), Tsb);                                   // :337-338
        const PfBeam b = beams[bi];
        s_eval[pl * n_beams + b.slot] = pf_eval_one<CORR>(bvh, Tsm, b, prm, sigma_quad, denom);
    }
    __syncthreads();
    if (threadIdx.x < np) {
        b2_particle_attr* ap = attrs + (order ? order[p0 + threadIdx.x] : p0 + threadIdx.x);
        b2_gaussian1d lk = ap->likelihood;
        pf_merge(lk, s_eval + threadIdx.x * n_beams, n_beams);
        ap->likelihood = lk;
    }
}

#ifdef __CUDACC__
// sort key of a particle for MAP 1: 16 heading bins (yaw), then the Morton code of its (x, y) cell on a 1024 x 1024 lattice over the map
__global__ void k_pf_sort_keys(const b2_transform* __restrict__ poses, uint32_t n, float bx, float by, uint32_t* __restrict__ keys, uint32_t* __restrict__ idx)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Tf T = tf_load(poses + i);
    const float yaw = atan2f(2.0f * (T.R.w * T.R.z + T.R.x * T.R.y), 1.0f - 2.0f * (T.R.y * T.R.y + T.R.z * T.R.z));
    const uint32_t yb = min(15u, (uint32_t)((yaw + 3.14159265f) * (16.0f / 6.2831853f)));
    auto cell = [](float v, float b) { const float u = (v / (b + 1e-6f)) * 0.5f + 0.5f; return (uint32_t)min(1023.0f, max(0.0f, u * 1023.0f)); };
    auto spread = [](uint32_t x) { x &= 0x3ffu; x = (x | (x << 8)) & 0x00ff00ffu; x = (x | (x << 4)) & 0x0f0f0f0fu; x = (x | (x << 2)) & 0x33333333u; x = (x | (x << 1)) & 0x55555555u; return x; };
    keys[i] = (yb << 20) | spread(cell(T.t.x, bx)) | (spread(cell(T.t.y, by)) << 1);
    idx[i] = i;
}
#endif

// ---------------------------------------------------------------------------------------------------------------------
// rest of the PF cycle
// ---------------------------------------------------------------------------------------------------------------------
// particle_move_and_forget_kernel (rmcl_ros/src/rmcl/particle_motion.cu:11-34): HBM stream, 2 x (32 + 36) B per particle
// With COLLIDE the wall check of the CPU updater (TFMotionUpdaterCPU.cpp:17-50,205-216) rides along: one ray from the old to the new
// position, tfar = their distance; a hit pins the likelihood to {0, 0, MAX_N_MEAS}.
B2_DEV void pf_motion_one(const BvhView* bvh, b2_transform* pose_io, b2_particle_attr* attr_io, Tf T, double forget_rate)
{
    const Tf pose_old = tf_load(pose_io);
    const Tf pose_new = tf_mul(pose_old, T);
    b2_gaussian1d lk = attr_io->likelihood;
    lk.n_meas = (uint32_t)((double)lk.n_meas - forget_rate * (double)lk.n_meas);
    if (bvh) {
        const V3 vec = v_sub(pose_new.t, pose_old.t);
        const float length = v_l2norm(vec);
        if (!(length < 0.00001f)) {
            const RaySetup r = ray_setup(pose_old.t, mk3(dvd(vec.x, length), dvd(vec.y, length), dvd(vec.z, length)), *bvh);
            HitRec h = trace_init(length);
            uint32_t nn = 0, nt = 0;
            trace_closest<false>(*bvh, r, h, nn, nt);
            if (h.face != B2_NOFACE) { lk.mean = 0.0f; lk.sigma = 0.0f; lk.n_meas = 10000u; }
        }
    }
    tf_store(pose_io, pose_new);
    attr_io->likelihood = lk;
}
#ifdef __CUDACC__
template <bool COLLIDE>
__global__ void __launch_bounds__(128) k_pf_motion(BvhView bvh, b2_transform* __restrict__ poses, b2_particle_attr* __restrict__ attrs, uint32_t n, b2_transform T_val, double forget_rate)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    pf_motion_one(COLLIDE ? &bvh : nullptr, poses + i, attrs + i, tf_from_pod(T_val), forget_rate);
}
#endif

// ---------------------------------------------------------------------------------------------------------------------
// Gladiator resampling (rmcl_ros/src/rmcl/resampling.cu:108-199).  Draws: Philox4x32-10 (Salmon et al., SC'11), counter =
// (global particle index, 0, step, block), key = seed -- one raw u32 (opponent) + three Box-Muller pairs per particle; a function
// of the GLOBAL index only, so the result does not depend on how particles are sharded over GPUs.  See DESIGN.md for why the
// reference's cuRAND XORWOW stream is not reproduced.
// ---------------------------------------------------------------------------------------------------------------------
B2_DEV void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t out[4])
{
    #pragma unroll
    for (int r = 0; r < 10; r++) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
B2_DEV float u01_open(uint32_t r) { return mul(add((float)(r >> 8), 0.5f), 5.9604644775390625e-08f); }
B2_DEV void gladiator_draws(uint64_t seed, uint32_t step, uint32_t gidx, uint32_t& raw, float N[6])
{
    uint32_t r[8];
    philox4x32_10(gidx, 0u, step, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), r);
    philox4x32_10(gidx, 0u, step, 1u, (uint32_t)seed, (uint32_t)(seed >> 32), r + 4);
    raw = r[0];
    #pragma unroll
    for (int p = 0; p < 3; p++) {
        const float u1 = u01_open(r[1 + 2 * p]), u2 = u01_open(r[2 + 2 * p]);
        const float rad = sqrtf(mul(-2.0f, logf(u1))), ang = mul(6.283185307179586f, u2);
        N[2 * p] = mul(rad, cosf(ang)); N[2 * p + 1] = mul(rad, sinf(ang));
    }
}
// rm::EulerAngles <-> Quaternion [RM-recalled: standard ZYX conversions], op order as in oracle/oracle.c
B2_DEV void quat_to_euler(Q4 q, float& roll, float& pitch, float& yaw)
{
    roll = atan2f(mul(2.0f, add(mul(q.w, q.x), mul(q.y, q.z))), sub(1.0f, mul(2.0f, add(mul(q.x, q.x), mul(q.y, q.y)))));
    const float sinp = mul(2.0f, sub(mul(q.w, q.y), mul(q.z, q.x)));
    pitch = fabsf(sinp) >= 1.0f ? copysignf(1.5707963267948966f, sinp) : asinf(sinp);
    yaw = atan2f(mul(2.0f, add(mul(q.w, q.z), mul(q.x, q.y))), sub(1.0f, mul(2.0f, add(mul(q.y, q.y), mul(q.z, q.z)))));
}
B2_DEV Q4 euler_to_quat(float roll, float pitch, float yaw)
{
    const float cr = cosf(mul(roll, 0.5f)), sr = sinf(mul(roll, 0.5f)), cp = cosf(mul(pitch, 0.5f)), sp = sinf(mul(pitch, 0.5f)), cy = cosf(mul(yaw, 0.5f)), sy = sinf(mul(yaw, 0.5f));
    Q4 q;
    q.w = add(mul(mul(cr, cp), cy), mul(mul(sr, sp), sy));
    q.x = sub(mul(mul(sr, cp), cy), mul(mul(cr, sp), sy));
    q.y = add(mul(mul(cr, sp), cy), mul(mul(sr, cp), sy));
    q.z = sub(mul(mul(cr, cp), sy), mul(mul(sr, sp), cy));
    return q;
}
// the opponent won (resampling.cu:150-192): the champion's slot receives a perturbed copy of the opponent whose n_meas is reduced by the forget rate
B2_DEV void gladiator_take(b2_transform pn, b2_particle_attr an, const float N[6], const b2_gladiator_config& cfg, b2_transform* out_p, b2_particle_attr* out_a)
{
    const Tf pose = tf_from_pod(pn);
    Tf pnew = pose;
    pnew.t = mk3(add(pose.t.x, mul(N[0], cfg.min_noise_tx)), add(pose.t.y, mul(N[1], cfg.min_noise_ty)), add(pose.t.z, mul(N[2], cfg.min_noise_tz)));   // :166-168
    float roll, pitch, yaw; quat_to_euler(pose.R, roll, pitch, yaw);
    pnew.R = euler_to_quat(add(roll, mul(N[3], cfg.min_noise_roll)), add(pitch, mul(N[4], cfg.min_noise_pitch)), add(yaw, mul(N[5], cfg.min_noise_yaw)));   // :169-173
    const Tf diff = tf_mul(tf_inv(pose), pnew);                                            // :175
    const float trans_dist = v_l2norm(diff.t);                                             // :178
    const float rot_dist = sqrtf(add(add(add(mul(diff.R.x, diff.R.x), mul(diff.R.y, diff.R.y)), mul(diff.R.z, diff.R.z)), mul(diff.R.w, diff.R.w)));   // :179
    const float frs = (float)(1.0 - pow(1.0 - (double)cfg.likelihood_forget_per_meter, (double)trans_dist));    // :182
    const float frr = (float)(1.0 - pow(1.0 - (double)cfg.likelihood_forget_per_radian, (double)rot_dist));     // :183
    const float forget = frs > frr ? frs : frr;
    const float remember = (float)(1.0 - (double)forget);                                  // :185
    an.likelihood.n_meas = (uint32_t)mul((float)an.likelihood.n_meas, remember);           // :187
    pn.R.x = pnew.R.x; pn.R.y = pnew.R.y; pn.R.z = pnew.R.z; pn.R.w = pnew.R.w; pn.t.x = pnew.t.x; pn.t.y = pnew.t.y; pn.t.z = pnew.t.z;   // stamp: the enemy's
    *out_p = pn; *out_a = an;
}
// one champion: poses/attrs are the n_all particles, champion = global index, outputs are written at out_p / out_a
B2_DEV void gladiator_one(const b2_transform* poses, const b2_particle_attr* attrs, uint32_t n_all, uint32_t champion, uint32_t raw, const float N[6],
                          const b2_gladiator_config& cfg, b2_transform* out_p, b2_particle_attr* out_a)
{
    const uint32_t enemy = raw % n_all;                                                    // :137
    const float Lc = attrs[champion].likelihood.mean, Le = attrs[enemy].likelihood.mean;
    if (!(Le > Lc)) { *out_p = poses[champion]; *out_a = attrs[champion]; return; }       // :150, :193-196
    gladiator_take(poses[enemy], attrs[enemy], N, cfg, out_p, out_a);
}

#ifdef __CUDACC__
// HBM stream: 68 B in (+68 B gathered from the opponent when it wins) and 68 B out per particle
__global__ void __launch_bounds__(256) k_pf_gladiator(const b2_transform* __restrict__ poses, const b2_particle_attr* __restrict__ attrs, uint32_t n_all, uint32_t first,
                                                      uint32_t n_local, b2_transform* __restrict__ poses_new, b2_particle_attr* __restrict__ attrs_new,
                                                      b2_gladiator_config cfg, uint64_t seed, uint32_t step, const uint32_t* __restrict__ raw_in, const float* __restrict__ normals_in)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_local) return;
    uint32_t raw; float N[6];
    if (raw_in) { raw = raw_in[i]; for (int k = 0; k < 6; k++) N[k] = normals_in[6 * (size_t)i + k]; }
    else gladiator_draws(seed, step, first + i, raw, N);
    gladiator_one(poses, attrs, n_all, first + i, raw, N, cfg, poses_new + i, attrs_new + i);
}
__global__ void k_pf_gladiator_randoms(uint64_t seed, uint32_t step, uint32_t first, uint32_t n, uint32_t* __restrict__ raw_out, float* __restrict__ normals_out)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t raw; float N[6];
    gladiator_draws(seed, step, first + i, raw, N);
    raw_out[i] = raw;
    for (int k = 0; k < 6; k++) normals_out[6 * (size_t)i + k] = N[k];
}
#endif

#ifdef __CUDACC__
// Gladiator resampling with the particles SHARDED over GPUs and no all-gather: every rank publishes its particles in a buffer its peers map over
// NVLink (CUDA IPC); a champion reads its opponent's 4-byte likelihood straight from the owner's HBM and fetches the 68-byte record only when
// the opponent wins -- the reference draws opponents from ALL particles (resampling.cu:137), so this is the one stage of the cycle whose data
// crosses GPUs.  Shards have equal size n_per_rank; global index = rank * n_per_rank + local index; draws are keyed by the global index, so
// the concatenation over ranks equals the single-GPU result bit for bit.  `traffic` counts the bytes actually read from remote ranks.
#define B2_MAX_PEERS 16
struct PfPeers { const b2_transform* poses[B2_MAX_PEERS]; const b2_particle_attr* attrs[B2_MAX_PEERS]; uint32_t world, rank, n_per_rank, pad; };
__global__ void __launch_bounds__(256) k_pf_gladiator_p2p(PfPeers peers, b2_transform* __restrict__ poses_new, b2_particle_attr* __restrict__ attrs_new, b2_gladiator_config cfg,
                                                          uint64_t seed, uint32_t step, unsigned long long* __restrict__ traffic)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long remote = 0ull;
    if (i < peers.n_per_rank) {
        const uint32_t n_all = peers.n_per_rank * peers.world, champion = peers.rank * peers.n_per_rank + i;
        uint32_t raw; float N[6];
        gladiator_draws(seed, step, champion, raw, N);
        const uint32_t enemy = raw % n_all, er = enemy / peers.n_per_rank, ei = enemy % peers.n_per_rank;      // resampling.cu:137
        const b2_particle_attr* ea = peers.attrs[er] + ei;
        const float Lc = peers.attrs[peers.rank][i].likelihood.mean;
        const float Le = __ldcv(&ea->likelihood.mean);                                                       // peer memory: never from a stale cache line
        if (er != peers.rank) remote += 4ull;
        if (!(Le > Lc)) { poses_new[i] = peers.poses[peers.rank][i]; attrs_new[i] = peers.attrs[peers.rank][i]; }
        else {
            const b2_transform* ep = peers.poses[er] + ei;
            b2_transform pn; b2_particle_attr an;
            const uint4 p0 = __ldcv(reinterpret_cast<const uint4*>(ep)), p1 = __ldcv(reinterpret_cast<const uint4*>(ep) + 1);
            memcpy(&pn, &p0, 16); memcpy(reinterpret_cast<char*>(&pn) + 16, &p1, 16);
            const uint32_t* aw = reinterpret_cast<const uint32_t*>(ea);
            uint32_t w[9];
            #pragma unroll
            for (int k = 0; k < 9; k++) w[k] = __ldcv(aw + k);
            memcpy(&an, w, 36);
            if (er != peers.rank) remote += 68ull;
            gladiator_take(pn, an, N, cfg, poses_new + i, attrs_new + i);
        }
    }
    remote = warp_sum_u64(remote);
    if (traffic && (threadIdx.x & 31u) == 0u && remote) atomicAdd(traffic, remote);
}
#endif

// compute_stats (rmcl_ros/src/rmcl/resampling.cu:41-92) over all SMs: FP64 block sums + max, the last block combines in block order
__global__ void __launch_bounds__(256) k_pf_stats(const b2_particle_attr* __restrict__ attrs, uint32_t n, double* __restrict__ partials, unsigned int* __restrict__ ticket,
                                                  float* __restrict__ out /* [sum, max] */)
{
    __shared__ double s_sum[8]; __shared__ float s_max[8]; __shared__ bool is_last;
    double sum = 0.0; float mx = 0.0f;                                                     // max starts at 0 like the reference (:54)
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) { const float L = attrs[i].likelihood.mean; sum += (double)L; mx = fmaxf(mx, L); }
    sum = warp_sum(sum);
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) { s_sum[threadIdx.x >> 5] = sum; s_max[threadIdx.x >> 5] = mx; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0; float m = 0.0f;
        for (int w = 0; w < 8; w++) { a += s_sum[w]; m = fmaxf(m, s_max[w]); }
        partials[2 * blockIdx.x] = a; partials[2 * blockIdx.x + 1] = (double)m;
        __threadfence();
        is_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    // the last block combines the block partials with all its threads in a fixed order (a serial loop over ~600 partials on one thread
    // was 20 of this kernel's 29 us: one L2 round trip per iteration)
    double a = 0.0; float m = 0.0f;
    for (uint32_t b = threadIdx.x; b < gridDim.x; b += blockDim.x) { a += __ldcg(partials + 2 * b); m = fmaxf(m, (float)__ldcg(partials + 2 * b + 1)); }
    a = warp_sum(a);
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) { s_sum[threadIdx.x >> 5] = a; s_max[threadIdx.x >> 5] = m; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0; float tm = 0.0f;
        for (int w = 0; w < 8; w++) { t += s_sum[w]; tm = fmaxf(tm, s_max[w]); }
        out[0] = (float)t; out[1] = tm;
        *ticket = 0u;
    }
}

#include "icp_loop.cuh"
