// STUB of rmagine/math/types.h (tests/stubs/README.md): the math PODs that cross the boundary, with the operations rmcl's interface code uses.
#pragma once
#include <cmath>
#include <cstddef>
#include <cstdint>

#ifndef RMAGINE_INLINE_FUNCTION
#define RMAGINE_INLINE_FUNCTION inline
#endif

namespace rmagine {

template <typename DataT, unsigned int Rows, unsigned int Cols> struct Matrix_;
using Matrix3x3 = Matrix_<float, 3, 3>;

struct Vector3f { float x, y, z;
    Vector3f operator*(float s) const { return {x * s, y * s, z * s}; } Vector3f operator+(const Vector3f& o) const { return {x + o.x, y + o.y, z + o.z}; }
    Vector3f operator-(const Vector3f& o) const { return {x - o.x, y - o.y, z - o.z}; } float dot(const Vector3f& o) const { return x * o.x + y * o.y + z * o.z; }
    float l2norm() const { return std::sqrt(x * x + y * y + z * z); } };
using Vector = Vector3f; using Point = Vector3f; using Vector3 = Vector3f;

struct Quaternion { float x, y, z, w;
    static Quaternion Identity() { return {0.f, 0.f, 0.f, 1.f}; }
    Quaternion inv() const { return {-x, -y, -z, w}; }
    Quaternion operator*(const Quaternion& b) const
    { return {w * b.x + x * b.w + y * b.z - z * b.y, w * b.y - x * b.z + y * b.w + z * b.x, w * b.z + x * b.y - y * b.x + z * b.w, w * b.w - x * b.x - y * b.y - z * b.z}; }
    Vector3f operator*(const Vector3f& v) const { const Quaternion p{v.x, v.y, v.z, 0.f}; const Quaternion r = (*this) * p * inv(); return {r.x, r.y, r.z}; }
    float dot(const Quaternion& o) const { return x * o.x + y * o.y + z * o.z + w * o.w; }
    operator Matrix3x3() const;
    void normalizeInplace() { const float n = std::sqrt(x * x + y * y + z * z + w * w); x /= n; y /= n; z /= n; w /= n; } };

struct Transform { Quaternion R; Vector3f t; uint32_t stamp;
    static Transform Identity() { return Transform{Quaternion::Identity(), {0.f, 0.f, 0.f}, 0u}; }
    Transform operator*(const Transform& b) const { return Transform{R * b.R, R * b.t + t, stamp}; }
    Vector3f operator*(const Vector3f& p) const { return R * p + t; }
    Transform operator~() const { const Quaternion Ri = R.inv(); const Vector3f ti = Ri * t; return Transform{Ri, {-ti.x, -ti.y, -ti.z}, stamp}; } };

template <typename DataT, unsigned int Rows, unsigned int Cols> struct Matrix_ { DataT data[Rows * Cols];      // column-major
    DataT& operator()(unsigned r, unsigned c) { return data[c * Rows + r]; } const DataT& operator()(unsigned r, unsigned c) const { return data[c * Rows + r]; }
    DataT trace() const { DataT s = DataT(0); for (unsigned i = 0; i < (Rows < Cols ? Rows : Cols); i++) s += (*this)(i, i); return s; }
    Matrix_<DataT, Cols, Rows> T() const { Matrix_<DataT, Cols, Rows> r{}; for (unsigned i = 0; i < Rows; i++) for (unsigned j = 0; j < Cols; j++) r(j, i) = (*this)(i, j); return r; }
    template <unsigned int C2> Matrix_<DataT, Rows, C2> operator*(const Matrix_<DataT, Cols, C2>& o) const
    { Matrix_<DataT, Rows, C2> r{}; for (unsigned i = 0; i < Rows; i++) for (unsigned j = 0; j < C2; j++) { DataT a = DataT(0); for (unsigned k = 0; k < Cols; k++) a += (*this)(i, k) * o(k, j); r(i, j) = a; } return r; } };
inline Quaternion::operator Matrix3x3() const
{
    Matrix3x3 M{};
    M(0, 0) = 1.f - 2.f * (y * y + z * z); M(0, 1) = 2.f * (x * y - z * w); M(0, 2) = 2.f * (x * z + y * w);
    M(1, 0) = 2.f * (x * y + z * w); M(1, 1) = 1.f - 2.f * (x * x + z * z); M(1, 2) = 2.f * (y * z - x * w);
    M(2, 0) = 2.f * (x * z - y * w); M(2, 1) = 2.f * (y * z + x * w); M(2, 2) = 1.f - 2.f * (x * x + y * y);
    return M;
}

struct CrossStatistics { Vector3f dataset_mean, model_mean; Matrix3x3 covariance; unsigned int n_meas;
    static CrossStatistics Identity() { CrossStatistics s{}; return s; } };
struct Gaussian1D { float mean, sigma; unsigned int n_meas; static Gaussian1D Identity() { return Gaussian1D{0.f, 0.f, 0u}; } };
struct Interval { float min, max; bool inside(float v) const { return min <= v && v <= max; } };
struct DiscreteInterval { float min, inc; uint32_t size; float operator[](uint32_t i) const { return min + static_cast<float>(i) * inc; } };
static_assert(sizeof(Transform) == 32 && sizeof(CrossStatistics) == 64 && sizeof(Gaussian1D) == 12, "rmagine layouts (SURVEY.md Appendix B)");

}  // namespace rmagine
