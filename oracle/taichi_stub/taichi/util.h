// TEST INFRASTRUCTURE.  Stand-in for the headers of the (un-vendored) taichi-legacy core that the
// reference's hot-path sources include (src/particles.h:8-12, src/mpm_fwd.h:8-12, src/mpm.h:15-22,
// src/transfer.cpp:6-12, src/mpm.cpp:10-16).  With it the reference's OWN translation units compile
// where they lie, unmodified (oracle/particles_ref.cpp, kernel_ref.cpp, transfer_ref.cpp; `make -C
// oracle ref`), so that their lines — calculate_force() / plasticity(), friction_project(), the
// interpolation kernels, the P2G / G2P loops, the grid update, ordering, deletion, substep() — run
// here and pin the oracle's restatement of them.
//
// Restated here is the VOCABULARY only, with the meaning the call sites require:
//   VectorND<n,T>   n scalars, element-wise + - * /, scalar broadcast, dot/sum/length/abs/map/min/max/
//                   clamp/cast; 3- and 4-vectors of float are 16 bytes and expose `.v` (they hold an
//                   __m128 in the core: src/particles.h:80-82, src/transfer.cpp:490,503,929,951) —
//                   GridState's power-of-two size assert and the SSE loops need it
//   MatrixND<n,T>   n columns, M[i] = column i (README.md:314); Matrix(s) = s*I, Matrix(v) = diag(v),
//                   Matrix(c0,c1[,c2]) from columns; products, transpose(d), determinant, inverse,
//                   diag/trace/frobenius_norm(2)/elementwise_product/sum
//   svd(A,U,S,V)    (math/svd.h) A = U S V^T, U and V rotations, singular values descending in
//                   magnitude, the last one negative when det A < 0 (the convention of the implicit-QR
//                   3x3 SVDs graphics codes use; the real core's convention cannot be checked, so tests
//                   compare only what does not depend on it); polar_decomp(A,R,S): A = R S, R = U V^T;
//                   computed in double and rounded, so that differences seen by the tests come from the
//                   reference's formulas, not from this header's numerics
//   IndexND/RegionND  integer index boxes iterated last-axis-fastest (the stencil order of
//                   src/transfer.cpp:353-359)
//   Config          string -> number map with get(key, default) / has_key; other value types are
//                   accepted and dropped (the harness sets the solver's fields directly)
//   Unit, TC_IMPLEMENTATION / create_instance_placement  a name -> placement-constructor registry for the
//                   particle types (src/particle_allocator.h:62,71); the solver classes are not registered
//   ThreadedTaskManager::run, tbb::parallel_for/sort (stub_more.h)  serial loops / std::sort by default
//                   (a fixed order: every pin runs this way); OpenMP loops only when a harness raises
//                   stub_num_threads() to TIME the reference's loops
//   TC_STATIC_IF    (common/meta.h) `if constexpr`
//   logging / serialization / profiling macros collapse to nothing; textures, meshes, images, assets,
//   rigid bodies (dynamics/rigid_body.h) and the level set (dynamics/simulation.h: half-spaces in grid
//   units) exist so that the code naming them compiles — only the level set has behaviour.
#pragma once
#include <immintrin.h>

#include <algorithm>
#include <array>
#include <functional>
#include <memory>
#include <type_traits>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstdio>
#include <map>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#define TC_NAMESPACE_BEGIN namespace taichi {
#define TC_NAMESPACE_END }
#define TC_FORCE_INLINE inline __attribute__((always_inline))
#define TC_NOT_IMPLEMENTED throw std::runtime_error("not implemented");
#define TC_ERROR(...) throw std::runtime_error("TC_ERROR")
#define TC_WARN(...) ((void)0)
#define TC_INFO(...) ((void)0)
#define TC_STOP do { std::fprintf(stderr, "TC_STOP at %s:%d\n", __FILE__, __LINE__); std::abort(); } while (0)
#define TC_ALIGNED(x) alignas(x)
#define TC_IO_DEF_VIRT(...)
#define TC_IO_DEF_WITH_BASE(...)
#define TC_IO_DEF(...)
#define TC_IO(...)
#define TC_IO_DECL template <class S> void io(S &serializer) const
#define TC_IO_DECL_VIRT template <class S> void io(S &serializer) const
#define TC_SERIALIZER_IS(T) (false)
#define TC_P(x) ((void)0)
#define TC_TRACE(...) ((void)0)
#define TC_DEBUG(...) ((void)0)
#define TC_ASSERT(x) assert(x)
#define TC_ASSERT_INFO(x, ...) assert(x)
#define TC_STATIC_ASSERT(x) static_assert((x), "")
#define TC_STUB_CAT_(a, b) a##b
#define TC_STUB_CAT(a, b) TC_STUB_CAT_(a, b)
#define TC_INTERFACE(T) template <> struct stub_registers<T> { static constexpr bool value = true; }
#define TC_INTERFACE_DEF(T, name) static_assert(sizeof(T *) > 0, "")
// registration by name for create_instance_placement (src/particle_allocator.h:62,71)
#define TC_IMPLEMENTATION(base, derived, name) \
  static int TC_STUB_CAT(tc_stub_registrar_, __COUNTER__) = taichi::RegisterIf<taichi::stub_registers<base>::value, base, derived>::run(name)
#define CHECK(x) ((void)(x))

namespace taichi {
template <class T> struct stub_registers { static constexpr bool value = false; };
using real = float;
using float32 = float;
using float64 = double;
using int32 = int32_t;
using uint8 = uint8_t;
using uint16 = uint16_t;
using uint32 = uint32_t;
using uint64 = uint64_t;
using int64 = int64_t;
constexpr real operator"" _f(long double v) { return (real)v; }
constexpr real operator"" _f(unsigned long long v) { return (real)v; }
constexpr float64 operator"" _f64(long double v) { return (float64)v; }

using std::abs;
using std::max;
using std::min;
using std::pow;
using std::sqrt;
inline real clamp(real v, real lo, real hi) { return v < lo ? lo : (v > hi ? hi : v); }
inline real sqr(real a) { return a * a; }
template <int n, class T>
constexpr T pow(T a) {
  T r = 1;
  for (int i = 0; i < n; i++) r *= a;
  return r;
}
constexpr real eps = 1e-6f;
namespace math {
inline real radians(real deg) { return deg * real(3.14159265358979323846 / 180.0); }
inline real degrees(real rad) { return rad * real(180.0 / 3.14159265358979323846); }
}  // namespace math

template <class T>
struct InterfaceRegistry {
  static std::map<std::string, std::function<T *(void *)>> &map() {
    static std::map<std::string, std::function<T *(void *)>> m;
    return m;
  }
};
// only the particle interfaces are really registered (create_instance_placement needs them); registering the
// solver itself would instantiate every virtual member of MPM<2> and MPM<3>
template <bool enable, class Base, class Derived> struct RegisterIf { static int run(const std::string &) { return 0; } };
template <class Base, class Derived> struct RegisterIf<true, Base, Derived> {
  static int run(const std::string &name) {
    InterfaceRegistry<Base>::map()[name] = [](void *place) -> Base * { return new (place) Derived(); };
    return 0;
  }
};
template <class T>
inline T *create_instance_placement(const std::string &alias, void *place) {
  auto &m = InterfaceRegistry<T>::map();
  auto it = m.find(alias);
  if (it == m.end()) throw std::runtime_error("unknown implementation: " + alias);
  return it->second(place);
}

class Config {
  std::map<std::string, double> num;
  std::map<std::string, std::string> str;

 public:
  Config &set(const std::string &k, const std::string &v) { str[k] = v; return *this; }
  Config &set(const std::string &k, const char *v) { str[k] = v; return *this; }
  Config &set(const std::string &k, double v) { num[k] = v; return *this; }
  template <class T, class = typename std::enable_if<!std::is_arithmetic<T>::value && !std::is_convertible<T, std::string>::value>::type>
  Config &set(const std::string &, const T &) { return *this; }   // pointers, vectors: accepted, not stored
  bool has_key(const std::string &k) const { return num.count(k) != 0; }
  template <class T>
  T get(const std::string &k, const T &def) const {
    auto it = num.find(k);
    return it == num.end() ? def : (T)it->second;
  }
  template <class T> T get(const std::string &k) const { return get_impl(k, (T *)nullptr); }
  template <class T> T get_impl(const std::string &k, T *) const { auto it = num.find(k); return it == num.end() ? T() : stub_cast<T>(it->second); }
  std::string get_impl(const std::string &k, std::string *) const { auto it = str.find(k); return it == str.end() ? std::string() : it->second; }
  template <class T> static typename std::enable_if<std::is_arithmetic<T>::value, T>::type stub_cast(double v) { return (T)v; }
  template <class T> static typename std::enable_if<!std::is_arithmetic<T>::value, T>::type stub_cast(double) { return T(); }
  template <class T> T *get_ptr(const std::string &) const { return nullptr; }
  std::string get_string(const std::string &k) const { auto it = str.find(k); return it == str.end() ? std::string() : it->second; }
  bool has_key_str(const std::string &k) const { return str.count(k) != 0; }
};

// number of threads the stand-in's parallel loops use: 1 (the default) = plain serial loops in index order, which
// is what the pin tests run; > 1 only when a harness asks for it to TIME the reference's loops
inline int &stub_num_threads() { static int n = 1; return n; }
struct ThreadedTaskManager {  // parallel for i in [0, n) (src/mpm.h:218-219)
  template <class F> static void run(int n, int num_threads, const F &f) {
    const int nt = std::min(num_threads, stub_num_threads());
    if (nt <= 1) { for (int i = 0; i < n; i++) f(i); return; }
#if defined(_OPENMP)
#pragma omp parallel for num_threads(nt) schedule(dynamic, 8)
#endif
    for (int i = 0; i < n; i++) f(i);
  }
};

class Unit {
 public:
  virtual void initialize(const Config &) {}
  virtual std::string get_name() const { return "unit"; }
  template <class S> void binary_io(S &) const {}
  virtual ~Unit() {}
};

template <int n, class T, bool simd = (sizeof(T) == 4 && std::is_floating_point<T>::value && (n == 3 || n == 4))>
struct VecStorage {
  static constexpr int storage = n;
  T d[n];
};
template <class T>
struct VecStorage<2, T, false> {
  static constexpr int storage = 2;
  union {
    T d[2];
    struct { T x, y; };
  };
};
template <int n, class T>
struct VecStorage<n, T, true> {  // float 3-/4-vectors ARE an __m128 in the core (`.v`, src/transfer.cpp:490,503,929,951)
  static constexpr int storage = 4;
  union {
    T d[4];
    __m128 v;
    struct { T x, y, z, w; };
  };
};

template <int n, class T>
struct VectorND : public VecStorage<n, T> {
  using VecStorage<n, T>::d;
  using VecStorage<n, T>::storage;
  VectorND() { for (int i = 0; i < storage; i++) d[i] = 0; }
  VectorND(T s) { for (int i = 0; i < storage; i++) d[i] = i < n ? s : 0; }
  VectorND(T a, T b) : VectorND() { static_assert(n == 2, ""); d[0] = a; d[1] = b; }
  VectorND(T a, T b, T c) : VectorND() { static_assert(n == 3, ""); d[0] = a; d[1] = b; d[2] = c; }
  VectorND(T a, T b, T c, T e) : VectorND() { static_assert(n == 4, ""); d[0] = a; d[1] = b; d[2] = c; d[3] = e; }
  VectorND(const VectorND<n - 1, T> &v, T last) : VectorND() { for (int i = 0; i < n - 1; i++) d[i] = v[i]; d[n - 1] = last; }
  explicit VectorND(const VectorND<n + 1, T> &v) : VectorND() { for (int i = 0; i < n; i++) d[i] = v[i]; }
  explicit VectorND(const VectorND<n - 1, T> &v) : VectorND() { for (int i = 0; i < n - 1; i++) d[i] = v[i]; }   // pad with 0
  template <class U, class = typename std::enable_if<!std::is_same<U, T>::value>::type>
  explicit VectorND(const VectorND<n, U> &v) : VectorND() { for (int i = 0; i < n; i++) d[i] = (T)v[i]; }
  explicit VectorND(const std::array<T, n> &a) : VectorND() { for (int i = 0; i < n; i++) d[i] = a[i]; }
  template <class U> VectorND<n, U> cast() const { VectorND<n, U> r; for (int i = 0; i < n; i++) r[i] = (U)d[i]; return r; }
  T min() const { T m = d[0]; for (int i = 1; i < n; i++) m = std::min(m, d[i]); return m; }
  VectorND clamp(const VectorND &lo, const VectorND &hi) const { VectorND r; for (int i = 0; i < n; i++) r.d[i] = std::min(std::max(d[i], lo.d[i]), hi.d[i]); return r; }
  bool abnormal() const { for (int i = 0; i < n; i++) if (!(d[i] == d[i]) || std::abs((double)d[i]) > 1e30) return true; return false; }
  operator std::array<T, n>() const { std::array<T, n> a; for (int i = 0; i < n; i++) a[i] = d[i]; return a; }
  bool operator==(const VectorND &o) const { for (int i = 0; i < n; i++) if (!(d[i] == o.d[i])) return false; return true; }
  bool operator!=(const VectorND &o) const { return !(*this == o); }
  bool operator<(const VectorND &o) const { for (int i = 0; i < n; i++) if (!(d[i] < o.d[i])) return false; return true; }
  bool operator<=(const VectorND &o) const { for (int i = 0; i < n; i++) if (!(d[i] <= o.d[i])) return false; return true; }
  static VectorND axis(int k) { VectorND r; r.d[k] = 1; return r; }
  static VectorND rand() { VectorND r; for (int i = 0; i < n; i++) r.d[i] = (T)std::rand() / (T)RAND_MAX; return r; }
  VectorND cross(const VectorND &o) const { static_assert(n == 3, ""); VectorND r; r.d[0] = d[1] * o.d[2] - d[2] * o.d[1]; r.d[1] = d[2] * o.d[0] - d[0] * o.d[2]; r.d[2] = d[0] * o.d[1] - d[1] * o.d[0]; return r; }
  template <class F, class = decltype(std::declval<F>()(0))>
  explicit VectorND(const F &f) : VectorND() { for (int i = 0; i < n; i++) d[i] = f(i); }  // VectorND([&](int i) { ... })
  VectorND(__m128 v) : VectorND() { alignas(16) float t[4]; _mm_store_ps(t, v); for (int i = 0; i < n && i < 4; i++) d[i] = (T)t[i]; }
  operator __m128() const { alignas(16) float t[4] = {0, 0, 0, 0}; for (int i = 0; i < n && i < 4; i++) t[i] = (float)d[i]; return _mm_load_ps(t); }
  T &operator[](int i) { return d[i]; }
  const T &operator[](int i) const { return d[i]; }
  T dot(const VectorND &o) const { T s = 0; for (int i = 0; i < n; i++) s += d[i] * o.d[i]; return s; }
  T sum() const { T s = 0; for (int i = 0; i < n; i++) s += d[i]; return s; }
  T length() const { return std::sqrt(dot(*this)); }
  T length2() const { return dot(*this); }
  T max() const { T m = d[0]; for (int i = 1; i < n; i++) m = std::max(m, d[i]); return m; }
  VectorND abs() const { VectorND r; for (int i = 0; i < n; i++) r.d[i] = std::abs(d[i]); return r; }
  template <class F>
  VectorND map(F f) const { VectorND r; for (int i = 0; i < n; i++) r.d[i] = f(d[i]); return r; }
  VectorND &operator+=(const VectorND &o) { for (int i = 0; i < n; i++) d[i] += o.d[i]; return *this; }
  VectorND &operator-=(const VectorND &o) { for (int i = 0; i < n; i++) d[i] -= o.d[i]; return *this; }
  VectorND &operator*=(T s) { for (int i = 0; i < n; i++) d[i] *= s; return *this; }
  VectorND &operator*=(const VectorND &o) { for (int i = 0; i < n; i++) d[i] *= o.d[i]; return *this; }
  VectorND operator-() const { VectorND r; for (int i = 0; i < n; i++) r.d[i] = -d[i]; return r; }
};
#define TCSTUB_VEC_OP(op)                                                                                                         \
  template <int n, class T> inline VectorND<n, T> operator op(const VectorND<n, T> &a, const VectorND<n, T> &b) {                 \
    VectorND<n, T> r; for (int i = 0; i < n; i++) r[i] = a[i] op b[i]; return r; }                                               \
  template <int n, class T> inline VectorND<n, T> operator op(const VectorND<n, T> &a, T s) {                                     \
    VectorND<n, T> r; for (int i = 0; i < n; i++) r[i] = a[i] op s; return r; }                                                  \
  template <int n, class T> inline VectorND<n, T> operator op(T s, const VectorND<n, T> &a) {                                     \
    VectorND<n, T> r; for (int i = 0; i < n; i++) r[i] = s op a[i]; return r; }
TCSTUB_VEC_OP(+)
TCSTUB_VEC_OP(-)
TCSTUB_VEC_OP(*)
TCSTUB_VEC_OP(/)
#undef TCSTUB_VEC_OP
template <int dim> struct IndexND;
// integer 3-vectors also answer to .x .y .z (src/kernel.h:190-192)
template <>
struct VectorND<3, int> {
  union {
    int d[3];
    struct { int x, y, z; };
  };
  VectorND() : d{0, 0, 0} {}
  VectorND(int s) : d{s, s, s} {}
  VectorND(int a, int b, int c) : d{a, b, c} {}
  explicit VectorND(const std::array<int, 3> &a) : d{a[0], a[1], a[2]} {}
  template <class F, class = decltype(std::declval<F>()(0))>
  explicit VectorND(const F &f) : d{0, 0, 0} { for (int i = 0; i < 3; i++) d[i] = f(i); }
  int &operator[](int i) { return d[i]; }
  const int &operator[](int i) const { return d[i]; }
  template <class U> VectorND<3, U> cast() const { VectorND<3, U> r; for (int i = 0; i < 3; i++) r[i] = (U)d[i]; return r; }
  explicit VectorND(const IndexND<3> &idx);
  operator std::array<int, 3>() const { return std::array<int, 3>{d[0], d[1], d[2]}; }
  bool operator==(const VectorND &o) const { return d[0] == o.d[0] && d[1] == o.d[1] && d[2] == o.d[2]; }
  bool operator!=(const VectorND &o) const { return !(*this == o); }
  bool operator<(const VectorND &o) const { return d[0] < o.d[0] && d[1] < o.d[1] && d[2] < o.d[2]; }
  bool operator<=(const VectorND &o) const { return d[0] <= o.d[0] && d[1] <= o.d[1] && d[2] <= o.d[2]; }
  int min() const { return std::min(d[0], std::min(d[1], d[2])); }
  int max() const { return std::max(d[0], std::max(d[1], d[2])); }
  int prod() const { return d[0] * d[1] * d[2]; }
};
template <int n, class T> inline std::array<T, n> to_std_array(const VectorND<n, T> &v) { std::array<T, n> a; for (int i = 0; i < n; i++) a[i] = v[i]; return a; }
template <int n, class T> inline VectorND<n, T> fract(const VectorND<n, T> &a) { VectorND<n, T> r; for (int i = 0; i < n; i++) r[i] = a[i] - std::floor(a[i]); return r; }
template <int n, class T> inline VectorND<n, T> fused_mul_add(const VectorND<n, T> &a, const VectorND<n, T> &b, const VectorND<n, T> &c) {
  VectorND<n, T> r; for (int i = 0; i < n; i++) r[i] = std::fma(a[i], b[i], c[i]); return r;
}
template <int n, class T> inline T dot(const VectorND<n, T> &a, const VectorND<n, T> &b) { return a.dot(b); }
template <int n, class T> inline T length(const VectorND<n, T> &a) { return a.length(); }
template <int n, class T> inline T length2(const VectorND<n, T> &a) { return a.length2(); }
template <int n> struct Element {};
template <int n, class T> inline VectorND<n, T> normalized(const VectorND<n, T> &a) { return a * (T(1) / a.length()); }

using Vector2 = VectorND<2, real>;
using Vector3 = VectorND<3, real>;
using Vector4 = VectorND<4, real>;
using Vector3f = VectorND<3, float>;
using Vector4f = VectorND<4, float>;
using Vector2i = VectorND<2, int>;
using Vector3i = VectorND<3, int>;

template <int n, class T>
struct MatrixND {
  using Vec = VectorND<n, T>;
  Vec c[n];  // columns
  MatrixND() {}
  MatrixND(T s) { for (int i = 0; i < n; i++) c[i][i] = s; }
  explicit MatrixND(const Vec &diag) { for (int i = 0; i < n; i++) c[i][i] = diag[i]; }
  MatrixND(const Vec &c0, const Vec &c1) { static_assert(n == 2, ""); c[0] = c0; c[1] = c1; }
  MatrixND(const Vec &c0, const Vec &c1, const Vec &c2) { static_assert(n == 3, ""); c[0] = c0; c[1] = c1; c[2] = c2; }
  Vec &operator[](int i) { return c[i]; }
  const Vec &operator[](int i) const { return c[i]; }
  Vec diag() const { Vec r; for (int i = 0; i < n; i++) r[i] = c[i][i]; return r; }
  T trace() const { return diag().sum(); }
  T sum() const { T s = 0; for (int i = 0; i < n; i++) s += c[i].sum(); return s; }
  T frobenius_norm2() const { T s = 0; for (int i = 0; i < n; i++) s += c[i].dot(c[i]); return s; }
  T frobenius_norm() const { return std::sqrt(frobenius_norm2()); }
  MatrixND elementwise_product(const MatrixND &o) const { MatrixND r; for (int i = 0; i < n; i++) r.c[i] = c[i] * o.c[i]; return r; }
  MatrixND transposed() const { MatrixND r; for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) r.c[i][j] = c[j][i]; return r; }
  MatrixND operator-() const { MatrixND r; for (int i = 0; i < n; i++) r.c[i] = -c[i]; return r; }
  MatrixND &operator+=(const MatrixND &o) { for (int i = 0; i < n; i++) c[i] += o.c[i]; return *this; }
  MatrixND &operator-=(const MatrixND &o) { for (int i = 0; i < n; i++) c[i] -= o.c[i]; return *this; }
  static MatrixND outer_product(const Vec &a, const Vec &b) { MatrixND r; for (int i = 0; i < n; i++) r.c[i] = a * b[i]; return r; }
};
template <int n, class T> inline MatrixND<n, T> operator+(const MatrixND<n, T> &a, const MatrixND<n, T> &b) { MatrixND<n, T> r; for (int i = 0; i < n; i++) r[i] = a[i] + b[i]; return r; }
template <int n, class T> inline MatrixND<n, T> operator-(const MatrixND<n, T> &a, const MatrixND<n, T> &b) { MatrixND<n, T> r; for (int i = 0; i < n; i++) r[i] = a[i] - b[i]; return r; }
template <int n, class T> inline MatrixND<n, T> operator*(const MatrixND<n, T> &a, T s) { MatrixND<n, T> r; for (int i = 0; i < n; i++) r[i] = a[i] * s; return r; }
template <int n, class T> inline MatrixND<n, T> operator*(T s, const MatrixND<n, T> &a) { MatrixND<n, T> r; for (int i = 0; i < n; i++) r[i] = s * a[i]; return r; }
template <int n, class T> inline VectorND<n, T> operator*(const MatrixND<n, T> &a, const VectorND<n, T> &v) { VectorND<n, T> r; for (int i = 0; i < n; i++) r += a[i] * v[i]; return r; }
template <int n, class T> inline MatrixND<n, T> operator*(const MatrixND<n, T> &a, const MatrixND<n, T> &b) { MatrixND<n, T> r; for (int i = 0; i < n; i++) r[i] = a * b[i]; return r; }
template <int n, class T> inline MatrixND<n, T> transposed(const MatrixND<n, T> &a) { return a.transposed(); }
template <int n, class T> inline MatrixND<n, T> transpose(const MatrixND<n, T> &a) { return a.transposed(); }
template <class T> inline T determinant(const MatrixND<2, T> &a) { return a[0][0] * a[1][1] - a[1][0] * a[0][1]; }
template <class T> inline T determinant(const MatrixND<3, T> &a) {
  return a[0][0] * (a[1][1] * a[2][2] - a[2][1] * a[1][2]) - a[1][0] * (a[0][1] * a[2][2] - a[2][1] * a[0][2]) +
         a[2][0] * (a[0][1] * a[1][2] - a[1][1] * a[0][2]);
}
template <class T> inline MatrixND<2, T> inversed(const MatrixND<2, T> &a) {
  T id = T(1) / determinant(a);
  return MatrixND<2, T>(VectorND<2, T>(a[1][1] * id, -a[0][1] * id), VectorND<2, T>(-a[1][0] * id, a[0][0] * id));
}
template <class T> inline MatrixND<3, T> inversed(const MatrixND<3, T> &a) {
  // adjugate / determinant, evaluated in double
  double m[3][3];
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) m[j][i] = a[i][j];  // m[row][col]
  double det = m[0][0] * (m[1][1] * m[2][2] - m[1][2] * m[2][1]) - m[0][1] * (m[1][0] * m[2][2] - m[1][2] * m[2][0]) +
               m[0][2] * (m[1][0] * m[2][1] - m[1][1] * m[2][0]);
  double id = 1.0 / det;
  MatrixND<3, T> r;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
      double cof = m[j1][i1] * m[j2][i2] - m[j1][i2] * m[j2][i1];  // cofactor of (j,i) -> inverse entry (i,j)
      r[j][i] = (T)(cof * id);                                      // r(row i, col j)
    }
  return r;
}
// 4x4 (gather_cdf's weighted least squares, src/rigid_transfer.cpp:251-252): cofactor expansion, evaluated in double and rounded
// like the 3x3 inverse above — differences seen by the tests come from the reference's formulas, not from this header
template <class T> inline double det3_rows(const double m[4][4], const int r[3], const int c[3]) {
  return m[r[0]][c[0]] * (m[r[1]][c[1]] * m[r[2]][c[2]] - m[r[1]][c[2]] * m[r[2]][c[1]]) - m[r[0]][c[1]] * (m[r[1]][c[0]] * m[r[2]][c[2]] - m[r[1]][c[2]] * m[r[2]][c[0]]) +
         m[r[0]][c[2]] * (m[r[1]][c[0]] * m[r[2]][c[1]] - m[r[1]][c[1]] * m[r[2]][c[0]]);
}
template <class T> inline double cofactor4(const MatrixND<4, T> &a, int row, int col) {
  double m[4][4];
  for (int c = 0; c < 4; c++) for (int r = 0; r < 4; r++) m[r][c] = a[c][r];
  int rr[3], cc[3], k = 0, l = 0;
  for (int i = 0; i < 4; i++) { if (i != row) rr[k++] = i; if (i != col) cc[l++] = i; }
  return (((row + col) & 1) ? -1.0 : 1.0) * det3_rows<T>(m, rr, cc);
}
template <class T> inline T determinant(const MatrixND<4, T> &a) {
  double d = 0;
  for (int c = 0; c < 4; c++) d += (double)a[c][0] * cofactor4(a, 0, c);
  return (T)d;
}
template <class T> inline MatrixND<4, T> inversed(const MatrixND<4, T> &a) {
  double d = 0;
  for (int c = 0; c < 4; c++) d += (double)a[c][0] * cofactor4(a, 0, c);
  MatrixND<4, T> r;
  for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) r[j][i] = (T)(cofactor4(a, j, i) / d);   // inverse(i,j) = cofactor(j,i) / det
  return r;
}
template <int n, class T> inline VectorND<n, T> cross(const VectorND<n, T> &a, const VectorND<n, T> &b) { return a.cross(b); }
template <int n, class T> inline MatrixND<n, T> inverse(const MatrixND<n, T> &a) { return inversed(a); }


template <int dim>
struct IndexND {
  VectorND<dim, int> i;
  VectorND<dim, real> storage_offset = VectorND<dim, real>(0.5f);
  VectorND<dim, int> get_ipos() const { return i; }
  int &operator[](int k) { return i[k]; }
  int operator[](int k) const { return i[k]; }
  IndexND &operator=(const VectorND<dim, int> &v) { i = v; return *this; }
  VectorND<dim, real> get_pos() const { return i.template cast<real>() + storage_offset; }  // cell centre unless the region says otherwise
};
// iteration space [lo, hi) in lexicographic order, last axis fastest (stencil node n <-> (n/9, n/3%3, n%3), src/transfer.cpp:353-359)
template <int dim> inline VectorND<dim, int> operator+(const IndexND<dim> &a, const VectorND<dim, int> &b) { return a.get_ipos() + b; }
template <int dim>
struct RegionND {
  VectorND<dim, int> lo, hi;
  VectorND<dim, real> storage_offset = VectorND<dim, real>(0.5f);
  RegionND() {}
  RegionND(const VectorND<dim, int> &l, const VectorND<dim, int> &h) : lo(l), hi(h) {}
  RegionND(const VectorND<dim, int> &l, const VectorND<dim, int> &h, const VectorND<dim, real> &o) : lo(l), hi(h), storage_offset(o) {}
  struct iterator {
    IndexND<dim> idx;
    const RegionND *r;
    bool done;
    IndexND<dim> &operator*() { return idx; }
    bool operator!=(const iterator &o) const { return done != o.done; }
    iterator &operator++() {
      for (int a = dim - 1; a >= 0; a--) {
        if (++idx.i[a] < r->hi[a]) return *this;
        idx.i[a] = r->lo[a];
      }
      done = true;
      return *this;
    }
  };
  iterator begin() const { iterator it; it.idx.i = lo; it.idx.storage_offset = storage_offset; it.r = this; it.done = false; for (int a = 0; a < dim; a++) if (lo[a] >= hi[a]) it.done = true; return it; }
  iterator end() const { iterator it; it.r = this; it.done = true; return it; }
};

inline VectorND<3, int>::VectorND(const IndexND<3> &idx) : d{idx.i[0], idx.i[1], idx.i[2]} {}

using Matrix2 = MatrixND<2, real>;
using Matrix3 = MatrixND<3, real>;
}  // namespace taichi
namespace fmt {
template <class... A> inline std::string format(const char *, A &&...) { return std::string(); }
template <class... A> inline void print(FILE *, const char *, A &&...) {}
}  // namespace fmt
