// eigen_shim.hpp -- minimal fixed-size stand-in for the handful of Eigen3 types
// the reference's hot-path sources use.  TEST INFRASTRUCTURE: it exists only so
// that /root/reference/src/{ndt_model,scan,scan_matcher_ndt,particle_filter,
// motion_model}.cpp can be compiled unmodified, in place, into oracle/_ref/
// (Eigen3 itself is an un-vendored, version-unpinned dependency of the
// reference -- package.xml:12, CMakeLists.txt:17 -- and is absent here).
// This is our own code, not a copy of Eigen.  Evaluation order follows what
// Eigen's lazy fixed-size products do (left-to-right dot products); the only
// intentionally approximate piece is EigenSolver<Matrix2d> (closed form instead
// of a Schur iteration: ulps apart, used only to pick the clamp branch).
#ifndef NDT2D_ORACLE_EIGEN_SHIM_HPP_
#define NDT2D_ORACLE_EIGEN_SHIM_HPP_

#include <cmath>
#include <cstddef>

namespace Eigen
{

template<int R, int C>
struct Mat
{
  double d[R * C];

  Mat() {}
  // Vector2d(x, y) / RowVector2d
  Mat(double a, double b)
  {
    static_assert(R * C == 2, "two-coefficient constructor");
    d[0] = a; d[1] = b;
  }
  Mat(double a, double b, double c)
  {
    static_assert(R * C == 3, "three-coefficient constructor");
    d[0] = a; d[1] = b; d[2] = c;
  }

  static Mat Zero()
  {
    Mat m;
    for (int i = 0; i < R * C; ++i) {m.d[i] = 0.0;}
    return m;
  }
  static Mat UnitZ()
  {
    static_assert(R == 3 && C == 1, "UnitZ on Vector3d");
    return Mat(0.0, 0.0, 1.0);
  }

  double & operator()(std::size_t i, std::size_t j) {return d[i * C + j];}
  const double & operator()(std::size_t i, std::size_t j) const {return d[i * C + j];}
  double & operator()(std::size_t i)
  {
    static_assert(R == 1 || C == 1, "vector access");
    return d[i];
  }
  const double & operator()(std::size_t i) const
  {
    static_assert(R == 1 || C == 1, "vector access");
    return d[i];
  }

  Mat<C, R> transpose() const
  {
    Mat<C, R> t;
    for (int i = 0; i < R; ++i) {
      for (int j = 0; j < C; ++j) {t.d[j * R + i] = d[i * C + j];}
    }
    return t;
  }

  Mat & operator+=(const Mat & o)
  {
    for (int i = 0; i < R * C; ++i) {d[i] += o.d[i];}
    return *this;
  }

  // Only meaningful for 2x2 / 3x3 (see below)
  Mat inverse() const;
};

template<int R, int C>
inline Mat<R, C> operator+(const Mat<R, C> & a, const Mat<R, C> & b)
{
  Mat<R, C> r;
  for (int i = 0; i < R * C; ++i) {r.d[i] = a.d[i] + b.d[i];}
  return r;
}
template<int R, int C>
inline Mat<R, C> operator-(const Mat<R, C> & a, const Mat<R, C> & b)
{
  Mat<R, C> r;
  for (int i = 0; i < R * C; ++i) {r.d[i] = a.d[i] - b.d[i];}
  return r;
}
template<int R, int C>
inline Mat<R, C> operator*(const Mat<R, C> & a, double s)
{
  Mat<R, C> r;
  for (int i = 0; i < R * C; ++i) {r.d[i] = a.d[i] * s;}
  return r;
}
template<int R, int C>
inline Mat<R, C> operator*(double s, const Mat<R, C> & a)
{
  Mat<R, C> r;
  for (int i = 0; i < R * C; ++i) {r.d[i] = s * a.d[i];}
  return r;
}
template<int R, int C>
inline Mat<R, C> operator/(const Mat<R, C> & a, double s)
{
  Mat<R, C> r;
  for (int i = 0; i < R * C; ++i) {r.d[i] = a.d[i] / s;}
  return r;
}

// General product, each coefficient a left-to-right dot product.
template<int R, int K, int C>
struct ProductResult {typedef Mat<R, C> type;};

template<int R, int K, int C>
inline Mat<R, C> matmul(const Mat<R, K> & a, const Mat<K, C> & b)
{
  Mat<R, C> r;
  for (int i = 0; i < R; ++i) {
    for (int j = 0; j < C; ++j) {
      double acc = a.d[i * K + 0] * b.d[0 * C + j];
      for (int k = 1; k < K; ++k) {acc += a.d[i * K + k] * b.d[k * C + j];}
      r.d[i * C + j] = acc;
    }
  }
  return r;
}

// 1x1 results (row * column) convert to a scalar, as Eigen's do.
struct Scalar1
{
  double v;
  operator double() const {return v;}
};

template<int K>
inline Scalar1 operator*(const Mat<1, K> & a, const Mat<K, 1> & b)
{
  double acc = a.d[0] * b.d[0];
  for (int k = 1; k < K; ++k) {acc += a.d[k] * b.d[k];}
  return Scalar1{acc};
}
// column * row (outer product), row * matrix, matrix * column, matrix * matrix
template<int R, int C>
inline Mat<R, C> operator*(const Mat<R, 1> & a, const Mat<1, C> & b) {return matmul<R, 1, C>(a, b);}
template<int K>
inline Mat<1, K> operator*(const Mat<1, K> & a, const Mat<K, K> & b) {return matmul<1, K, K>(a, b);}
template<int K>
inline Mat<K, 1> operator*(const Mat<K, K> & a, const Mat<K, 1> & b) {return matmul<K, K, 1>(a, b);}

typedef Mat<2, 1> Vector2d;
typedef Mat<3, 1> Vector3d;
typedef Mat<1, 2> RowVector2d;
typedef Mat<1, 3> RowVector3d;
typedef Mat<2, 2> Matrix2d;
typedef Mat<3, 3> Matrix3d;

// Eigen's closed-form 2x2 inverse: invdet = 1 / det, then multiply.
template<>
inline Matrix2d Matrix2d::inverse() const
{
  const double det = d[0] * d[3] - d[2] * d[1];
  const double invdet = 1.0 / det;
  Matrix2d r;
  r.d[0] = d[3] * invdet;
  r.d[2] = -d[2] * invdet;
  r.d[1] = -d[1] * invdet;
  r.d[3] = d[0] * invdet;
  return r;
}

// Cofactor 3x3 inverse (used only outside the hot path, e.g. constraint.cpp).
template<>
inline Matrix3d Matrix3d::inverse() const
{
  const double * m = d;
  const double c00 = m[4] * m[8] - m[5] * m[7];
  const double c01 = m[5] * m[6] - m[3] * m[8];
  const double c02 = m[3] * m[7] - m[4] * m[6];
  const double det = m[0] * c00 + m[1] * c01 + m[2] * c02;
  const double invdet = 1.0 / det;
  Matrix3d r;
  r.d[0] = c00 * invdet;
  r.d[1] = (m[2] * m[7] - m[1] * m[8]) * invdet;
  r.d[2] = (m[1] * m[5] - m[2] * m[4]) * invdet;
  r.d[3] = c01 * invdet;
  r.d[4] = (m[0] * m[8] - m[2] * m[6]) * invdet;
  r.d[5] = (m[2] * m[3] - m[0] * m[5]) * invdet;
  r.d[6] = c02 * invdet;
  r.d[7] = (m[1] * m[6] - m[0] * m[7]) * invdet;
  r.d[8] = (m[0] * m[4] - m[1] * m[3]) * invdet;
  return r;
}

// EigenSolver<Matrix2d>(m).eigenvalues().real()
template<typename M>
class EigenSolver;

template<>
class EigenSolver<Matrix2d>
{
public:
  struct Values
  {
    Vector2d re;
    Vector2d real() const {return re;}
  };
  explicit EigenSolver(const Matrix2d & m)
  {
    const double a = m(0, 0), b = m(0, 1), c = m(1, 0), dd = m(1, 1);
    const double mid = 0.5 * (a + dd);
    const double half = 0.5 * (a - dd);
    const double disc = half * half + b * c;
    const double rad = disc > 0.0 ? std::sqrt(disc) : 0.0;
    vals_.re = Vector2d(mid - rad, mid + rad);
  }
  const Values & eigenvalues() const {return vals_;}

private:
  Values vals_;
};

// --- Geometry: just enough for conversions.hpp:64-68 -------------------

struct Translation3d
{
  double x, y, z;
  Translation3d(double x_, double y_, double z_) : x(x_), y(y_), z(z_) {}
};

struct AngleAxisd
{
  double angle;
  Vector3d axis;
  AngleAxisd(double a, const Vector3d & ax) : angle(a), axis(ax) {}
  // Rodrigues form, same operation order as Eigen's AngleAxis::toRotationMatrix
  Matrix3d toRotationMatrix() const
  {
    Matrix3d res;
    const Vector3d sin_axis = std::sin(angle) * axis;
    const double c = std::cos(angle);
    const Vector3d cos1_axis = (1.0 - c) * axis;
    double tmp;
    tmp = cos1_axis(0) * axis(1);
    res(0, 1) = tmp - sin_axis(2);
    res(1, 0) = tmp + sin_axis(2);
    tmp = cos1_axis(0) * axis(2);
    res(0, 2) = tmp + sin_axis(1);
    res(2, 0) = tmp - sin_axis(1);
    tmp = cos1_axis(1) * axis(2);
    res(1, 2) = tmp - sin_axis(0);
    res(2, 1) = tmp + sin_axis(0);
    res(0, 0) = cos1_axis(0) * axis(0) + c;
    res(1, 1) = cos1_axis(1) * axis(1) + c;
    res(2, 2) = cos1_axis(2) * axis(2) + c;
    return res;
  }
};

struct Isometry3d
{
  Matrix3d linear;
  Vector3d translation;
  Isometry3d() {}
  // res = translation; res += linear * p   (Eigen transform * vector)
  Vector3d operator*(const Vector3d & p) const
  {
    Vector3d r = translation;
    r += linear * p;
    return r;
  }
};

inline Isometry3d operator*(const Translation3d & t, const AngleAxisd & aa)
{
  Isometry3d iso;
  iso.linear = aa.toRotationMatrix();
  iso.translation = Vector3d(t.x, t.y, t.z);
  return iso;
}

}  // namespace Eigen

#endif  // NDT2D_ORACLE_EIGEN_SHIM_HPP_
