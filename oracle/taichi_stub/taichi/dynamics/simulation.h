// TEST INFRASTRUCTURE — see taichi/util.h.  Base class of the solver (vector typedefs, clock, thread count,
// level set slot, the virtual verbs) and the few helper types src/mpm.h names.
#pragma once
#include <taichi/util.h>
#include <taichi/common/meta.h>
#include <taichi/system/threading.h>
namespace taichi {
struct RenderParticle {};
// Static level set made of half-spaces, in GRID units: phi(X) = n.X + d, the largest-penetration (minimum phi)
// plane wins — the same construction as the oracle's planes_sdf.  sample / gradient / time derivative / inside
// are the calls src/mpm.cpp:303-337 makes; `inside` is the domain test that guards sampling.
template <int dim> struct LevelSet {
  real friction = 0;
  std::vector<VectorND<dim + 1, real>> planes;  // (n, d)
};
template <int dim> struct DynamicLevelSet {
  using Vector = VectorND<dim, real>;
  std::shared_ptr<LevelSet<dim>> levelset0;
  int best(const Vector &p, real &phi) const {
    int b = -1;
    phi = 1e30f;
    for (int k = 0; k < (int)levelset0->planes.size(); k++) {
      const auto &pl = levelset0->planes[k];
      real v = pl[dim];
      for (int a = 0; a < dim; a++) v += pl[a] * p[a];
      if (v < phi) { phi = v; b = k; }
    }
    return b;
  }
  real sample(const Vector &p, real) const { real phi; if (!levelset0 || best(p, phi) < 0) return 1e30f; return phi; }
  Vector get_spatial_gradient(const Vector &p, real) const {
    real phi;
    Vector g(0.0f);
    if (!levelset0) return g;
    int b = best(p, phi);
    if (b >= 0) for (int a = 0; a < dim; a++) g[a] = levelset0->planes[b][a];
    return g;
  }
  real get_temporal_derivative(const Vector &, real) const { return 0; }
  bool inside(const Vector &) const { return levelset0 != nullptr; }
};
struct BinaryInputSerializer { void initialize(const std::string &) {} void finalize() {} template <class T> void operator()(T &) {} };
struct BinaryOutputSerializer { void initialize() {} void finalize() {} void write_to_file(const std::string &) {} template <class T> void operator()(const T &) {} };

template <int DIM>
class Simulation : public Unit {
 public:
  static constexpr int dim = DIM;  // explicit specializations of the solver's members name it (src/mpm.cpp:683,1028)
  using Vector = VectorND<dim, real>;
  using VectorP = VectorND<dim + 1, real>;
  using VectorI = VectorND<dim, int>;
  using Vectori = VectorND<dim, int>;
  using Matrix = MatrixND<dim, real>;
  using MatrixP = MatrixND<dim + 1, real>;
  real current_t = 0;
  int num_threads = 1;
  DynamicLevelSet<dim> levelset;
  template <class S> void io(S &) const {}
  virtual std::string add_particles(const Config &) { return ""; }
  virtual void step(real) {}
  virtual std::vector<RenderParticle> get_render_particles() const { return {}; }
  virtual void visualize() const {}
  virtual std::string general_action(const Config &) { return ""; }
  virtual std::string get_debug_information() { return ""; }
  virtual bool test() const { return true; }
  virtual void binary_io(BinaryOutputSerializer &) const {}   // AsyncMPM overrides both (src/async/async_mpm.h:119-124)
  virtual void binary_io(BinaryInputSerializer &) const {}
};
using Simulation2D = Simulation<2>;
using Simulation3D = Simulation<3>;
}  // namespace taichi
