// TEST INFRASTRUCTURE — see taichi/util.h.  Base class of the solver (vector typedefs, clock, thread count,
// level set slot, the virtual verbs) and the few helper types src/mpm.h names.
#pragma once
#include <taichi/util.h>
#include <taichi/common/meta.h>
#include <taichi/system/threading.h>
namespace taichi {
struct RenderParticle {};
template <int dim> struct DynamicLevelSet {
  real friction = 0;
};
struct BinaryInputSerializer { void initialize(const std::string &) {} void finalize() {} template <class T> void operator()(T &) {} };
struct BinaryOutputSerializer { void initialize() {} void finalize() {} void write_to_file(const std::string &) {} template <class T> void operator()(const T &) {} };

template <int dim>
class Simulation : public Unit {
 public:
  using Vector = VectorND<dim, real>;
  using VectorP = VectorND<dim + 1, real>;
  using VectorI = VectorND<dim, int>;
  using Vectori = VectorND<dim, int>;
  using Matrix = MatrixND<dim, real>;
  using MatrixP = MatrixND<dim + 1, real>;
  real current_t = 0;
  int num_threads = 1;
  DynamicLevelSet<dim> levelset;
  template <class S> void io(S &) {}
  virtual std::string add_particles(const Config &) { return ""; }
  virtual void step(real) {}
  virtual std::vector<RenderParticle> get_render_particles() const { return {}; }
  virtual void visualize() const {}
  virtual std::string general_action(const Config &) { return ""; }
  virtual std::string get_debug_information() { return ""; }
  virtual bool test() const { return true; }
};
}  // namespace taichi
