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
  void reset_tmp_velocity() {}
  void apply_tmp_velocity() {}
  Vector get_velocity_at(const Vector &) const { return Vector(0.0f); }
  void apply_tmp_impulse(const Vector &, const Vector &) {}
};
}  // namespace taichi
