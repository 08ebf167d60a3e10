// TEST INFRASTRUCTURE.  A stand-in for the single-header "taichi.h" the reference's 88-line program
// includes (mls-mpm88.cpp:3; the real header ships in a release zip that is not in the reference tree,
// README.md:27).  It exists so that the reference's OWN translation unit /root/reference/mls-mpm88.cpp
// — the algorithm lines 16-69, unmodified, where they lie — can be compiled and run here as a pin for
// the oracle's 2-D restatement (oracle/mpm88_ref.cpp, `make -C oracle ref`).
//
// What is restated here is only the vocabulary that program uses: 2-vectors, 3-vectors, 2x2 matrices
// (column-major, M[i] = column i, README.md:314), element-wise arithmetic with scalar broadcast,
// Matrix(real) = real * identity (so `Mat(1)` is I and `matrix + scalar` adds to the diagonal, the
// only reading under which lines 27-28 are the fixed-corotated stress), determinant / transposed /
// outer_product, and the two factorizations: polar_decomp(A,R,S): A = R S with R a rotation (closed
// form in 2-D, unique for det A > 0) and svd(A,U,Sig,V): A = U Sig V^T (via the polar factor; sign
// and order conventions of the real header are unknown, the program's use — clamp of Sig's diagonal,
// U Sig V^T — does not depend on them for det A > 0).  The GUI is a no-op.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace taichi {
using real = float;  // mls-mpm88.cpp:131-132 ("real = float")
constexpr real operator"" _f(long double v) { return (real)v; }

struct Vector2i;
struct Vector3;
struct Vector2 {
  real x, y;
  Vector2() : x(0), y(0) {}
  Vector2(real s) : x(s), y(s) {}
  Vector2(real x_, real y_) : x(x_), y(y_) {}
  explicit Vector2(const Vector3 &v);
  real &operator[](int i) { return i ? y : x; }
  real operator[](int i) const { return i ? y : x; }
  template <class T> auto cast() const;
  Vector2 &operator+=(const Vector2 &o) { x += o.x; y += o.y; return *this; }
  static Vector2 rand() { return Vector2((real)std::rand() / (real)RAND_MAX, (real)std::rand() / (real)RAND_MAX); }
};
struct Vector2i {
  int x, y;
  template <class T> auto cast() const;
};
template <> inline auto Vector2::cast<int>() const { return Vector2i{(int)x, (int)y}; }      // truncation, as the C cast
template <> inline auto Vector2i::cast<real>() const { return Vector2((real)x, (real)y); }
inline Vector2 operator+(const Vector2 &a, const Vector2 &b) { return Vector2(a.x + b.x, a.y + b.y); }
inline Vector2 operator-(const Vector2 &a, const Vector2 &b) { return Vector2(a.x - b.x, a.y - b.y); }
inline Vector2 operator*(const Vector2 &a, const Vector2 &b) { return Vector2(a.x * b.x, a.y * b.y); }
inline Vector2 operator*(const Vector2 &a, real s) { return Vector2(a.x * s, a.y * s); }
inline Vector2 operator*(real s, const Vector2 &a) { return Vector2(s * a.x, s * a.y); }
inline Vector2 sqr(const Vector2 &a) { return a * a; }
inline real sqr(real a) { return a * a; }
inline real clamp(real v, real lo, real hi) { return std::min(std::max(v, lo), hi); }

struct Vector3 {
  real d[3];
  Vector3() : d{0, 0, 0} {}
  Vector3(real s) : d{s, s, s} {}
  Vector3(real a, real b, real c) : d{a, b, c} {}
  Vector3(const Vector2 &v, real c) : d{v.x, v.y, c} {}
  real &operator[](int i) { return d[i]; }
  real operator[](int i) const { return d[i]; }
  Vector3 &operator+=(const Vector3 &o) { d[0] += o.d[0]; d[1] += o.d[1]; d[2] += o.d[2]; return *this; }
  Vector3 &operator/=(real s) { d[0] /= s; d[1] /= s; d[2] /= s; return *this; }
};
inline Vector2::Vector2(const Vector3 &v) : x(v[0]), y(v[1]) {}
inline Vector3 operator+(const Vector3 &a, const Vector3 &b) { return Vector3(a[0] + b[0], a[1] + b[1], a[2] + b[2]); }
inline Vector3 operator*(real s, const Vector3 &a) { return Vector3(s * a[0], s * a[1], s * a[2]); }

struct Matrix2 {
  Vector2 c[2];  // columns
  Matrix2() {}
  Matrix2(real s) { c[0] = Vector2(s, 0); c[1] = Vector2(0, s); }  // s * identity
  Vector2 &operator[](int i) { return c[i]; }
  const Vector2 &operator[](int i) const { return c[i]; }
  Matrix2 &operator+=(const Matrix2 &o) { c[0] += o.c[0]; c[1] += o.c[1]; return *this; }
  static Matrix2 outer_product(const Vector2 &a, const Vector2 &b) {  // a b^T
    Matrix2 m;
    m.c[0] = a * b.x;
    m.c[1] = a * b.y;
    return m;
  }
};
inline Matrix2 operator+(const Matrix2 &a, const Matrix2 &b) { Matrix2 m; m.c[0] = a.c[0] + b.c[0]; m.c[1] = a.c[1] + b.c[1]; return m; }
inline Matrix2 operator-(const Matrix2 &a, const Matrix2 &b) { Matrix2 m; m.c[0] = a.c[0] - b.c[0]; m.c[1] = a.c[1] - b.c[1]; return m; }
inline Matrix2 operator*(real s, const Matrix2 &a) { Matrix2 m; m.c[0] = s * a.c[0]; m.c[1] = s * a.c[1]; return m; }
inline Vector2 operator*(const Matrix2 &a, const Vector2 &v) { return a.c[0] * v.x + a.c[1] * v.y; }
inline Matrix2 operator*(const Matrix2 &a, const Matrix2 &b) { Matrix2 m; m.c[0] = a * b.c[0]; m.c[1] = a * b.c[1]; return m; }
inline Matrix2 transposed(const Matrix2 &a) { Matrix2 m; m.c[0] = Vector2(a.c[0].x, a.c[1].x); m.c[1] = Vector2(a.c[0].y, a.c[1].y); return m; }
inline real determinant(const Matrix2 &a) { return a.c[0].x * a.c[1].y - a.c[1].x * a.c[0].y; }

// A = R S, R rotation, S symmetric (closed form: the rotation that symmetrises R^T A)
inline void polar_decomp(const Matrix2 &A, Matrix2 &R, Matrix2 &S) {
  real x = A.c[0].x + A.c[1].y, y = A.c[0].y - A.c[1].x;  // a00 + a11, a10 - a01
  real scale = 1.0f / std::sqrt(x * x + y * y);
  real c = x * scale, s = y * scale;
  R.c[0] = Vector2(c, s);
  R.c[1] = Vector2(-s, c);
  S = transposed(R) * A;
}
// A = U Sig V^T through the polar factor: S = V Sig V^T (Jacobi angle of the symmetric 2x2), U = R V
inline void svd(const Matrix2 &A, Matrix2 &U, Matrix2 &Sig, Matrix2 &V) {
  Matrix2 R, S;
  polar_decomp(A, R, S);
  real c, s;
  real s00 = S.c[0].x, s01 = S.c[1].x, s11 = S.c[1].y;
  if (std::abs(s01) < 1e-6f) {
    c = 1; s = 0;
    Sig = Matrix2(0); Sig.c[0].x = s00; Sig.c[1].y = s11;
  } else {
    real tao = 0.5f * (s00 - s11);
    real w = std::sqrt(tao * tao + s01 * s01);
    real t = tao > 0 ? s01 / (tao + w) : s01 / (tao - w);
    c = 1.0f / std::sqrt(t * t + 1);
    s = -t * c;
    Sig = Matrix2(0);
    Sig.c[0].x = c * c * s00 - 2 * c * s * s01 + s * s * s11;
    Sig.c[1].y = s * s * s00 + 2 * c * s * s01 + c * c * s11;
  }
  if (Sig.c[0].x < Sig.c[1].y) {  // descending order: swap and rotate V by 90 degrees
    std::swap(Sig.c[0].x, Sig.c[1].y);
    V.c[0] = Vector2(-s, -c);
    V.c[1] = Vector2(c, -s);
  } else {
    V.c[0] = Vector2(c, -s);
    V.c[1] = Vector2(s, c);
  }
  U = R * V;
}

// ---- no-op GUI (main() is compiled but never run by the harness)
struct Canvas {
  struct Shape {
    Shape &radius(real) { return *this; }
    Shape &color(int) { return *this; }
    Shape &close() { return *this; }
  };
  void clear(int) {}
  Shape rect(const Vector2 &, const Vector2 &) { return Shape(); }
  Shape circle(const Vector2 &) { return Shape(); }
};
struct GUI {
  Canvas canvas;
  GUI(const std::string &, int, int) {}
  Canvas &get_canvas() { return canvas; }
  void update() {}
};
}  // namespace taichi
