// TEST INFRASTRUCTURE — see taichi/util.h.  The rigid-body interface the transfer code names
// (src/transfer.cpp:196,241-252,...): present so that the coupled branches compile; the pinned fast
// path runs without rigid bodies, so none of it executes.
#pragma once
#include <taichi/util.h>
namespace taichi {
template <int dim>
struct RigidBody {
  using Vector = VectorND<dim, real>;
  using ElementType = int;
  int id = 0;
  real frictions[2] = {0, 0};
  Vector velocity;
  int pos_func_id = -1, rot_func_id = -1;
  using PositionFunctionType = std::function<Vector(real)>;
  using RotationFunctionType = std::function<Vector(real)>;
  PositionFunctionType pos_func;
  RotationFunctionType rot_func;
  void set_as_background() {}
  struct MeshElement { Vector v[dim]; };
  struct MeshType { std::vector<MeshElement> elements; };
  std::shared_ptr<MeshType> mesh;
  MatrixND<dim + 1, real> get_mesh_to_world() const { return MatrixND<dim + 1, real>(1.0f); }
  void reset_tmp_velocity() {}
  void apply_tmp_velocity() {}
  Vector get_velocity_at(const Vector &) const { return Vector(0.0f); }
  void apply_tmp_impulse(const Vector &, const Vector &) {}
};
template <int n> inline VectorND<n, real> transform(const MatrixND<n + 1, real> &, const VectorND<n, real> &v) { return v; }
template <class T> inline void trash(T &&) {}
}  // namespace taichi
