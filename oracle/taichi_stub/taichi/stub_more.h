// TEST INFRASTRUCTURE — see taichi/util.h.  Vocabulary that only src/mpm.cpp needs (seeding from textures /
// meshes / Poisson disks, debug images, serialization, TBB): enough for those parts to COMPILE.  None of it
// runs in the pinned path, which loads particles directly and steps with optimized transfers, a plane level
// set and no rigid bodies.
#pragma once
#include <taichi/util.h>
#include <fstream>
#include <mutex>
#include <numeric>
#if defined(_OPENMP)
#include <omp.h>
#include <parallel/algorithm>
#endif
namespace taichi {
struct Texture {
  template <int d> VectorND<4, real> sample(const VectorND<d, real> &) const { return VectorND<4, real>(0.0f); }
};
struct AssetManager {
  template <class T> static std::shared_ptr<T> get_asset(int) { return std::shared_ptr<T>(); }
};
struct Mesh {
  std::vector<VectorND<3, real>> vertices;
  void initialize(const Config &) {}
};
inline std::string absolute_path(const std::string &s) { return s; }
inline int rand_int() { return std::rand(); }
namespace Time {
struct Timer { explicit Timer(const std::string &) {} };
inline double get_time() { return 0; }
}  // namespace Time

template <int dim, class T>
struct ArrayND {
  VectorND<dim, int> res;
  std::vector<T> data;
  ArrayND() {}
  explicit ArrayND(const VectorND<dim, int> &r, T init = T()) { initialize(r, init); }
  void initialize(const VectorND<dim, int> &r, T init = T()) { res = r; size_t n = 1; for (int i = 0; i < dim; i++) n *= (size_t)std::max(r[i], 0); data.assign(n, init); }
  VectorND<dim, int> get_res() const { return res; }
  size_t lin(const VectorND<dim, int> &i) const { size_t o = 0; for (int k = 0; k < dim; k++) o = o * res[k] + i[k]; return o; }
  T &operator[](const VectorND<dim, int> &i) { return data[lin(i)]; }
  const T &operator[](const VectorND<dim, int> &i) const { return data[lin(i)]; }
  T &operator[](const IndexND<dim> &i) { return data[lin(i.get_ipos())]; }
  const T &operator[](const IndexND<dim> &i) const { return data[lin(i.get_ipos())]; }
  bool inside(const VectorND<dim, int> &i) const { for (int k = 0; k < dim; k++) if (i[k] < 0 || i[k] >= res[k]) return false; return true; }
  RegionND<dim> get_region() const { return RegionND<dim>(VectorND<dim, int>(0), res); }
  void write_as_image(const std::string &) const {}
  void reset_zero() { std::fill(data.begin(), data.end(), T()); }
};
template <class T> using Array2D = ArrayND<2, T>;
template <class T> using Array3D = ArrayND<3, T>;

template <class T> inline void write_to_binary_file(const T &, const std::string &) {}
template <class T> inline void read_from_binary_file(T &, const std::string &) {}
template <class T> inline std::unique_ptr<T> create_instance_unique(const std::string &, const Config & = Config()) { return std::unique_ptr<T>(); }
}  // namespace taichi

namespace tbb {
template <class I, class F> inline void parallel_for(I b, I e, const F &f) {
  const int nt = taichi::stub_num_threads();
  if (nt <= 1) { for (I i = b; i < e; i++) f(i); return; }
#if defined(_OPENMP)
#pragma omp parallel for num_threads(nt) schedule(static)
#endif
  for (I i = b; i < e; i++) f(i);
}
template <class It> inline void parallel_sort(It b, It e) {
#if defined(_OPENMP) && defined(_GLIBCXX_PARALLEL_ALGORITHM_H)
  if (taichi::stub_num_threads() > 1) { __gnu_parallel::sort(b, e); return; }
#endif
  std::sort(b, e);
}
// What src/async/async_mpm.{h,cpp} names of TBB (serial stand-ins: one "thread", vectors are std::vector)
template <class T> struct blocked_range {
  T b, e;
  blocked_range(T b_, T e_) : b(b_), e(e_) {}
  T begin() const { return b; }
  T end() const { return e; }
};
template <class T, class F> inline void parallel_for(const blocked_range<T> &r, const F &f) { f(r); }
template <class T> struct concurrent_vector : std::vector<T> {
  using std::vector<T>::vector;
};
template <class T> struct enumerable_thread_specific {
  T one[1];
  T &local() { return one[0]; }
  T *begin() { return one; }
  T *end() { return one + 1; }
};
}  // namespace tbb
#ifndef TC_ERROR_IF
#define TC_ERROR_IF(cond, ...) do { if (cond) throw std::runtime_error("TC_ERROR_IF"); } while (0)
#endif
#define TC_LOAD_CONFIG(name, default_val) this->name = config.get(#name, default_val)
