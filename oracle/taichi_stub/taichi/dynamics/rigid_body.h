// TEST INFRASTRUCTURE — see taichi/util.h.  Stand-in for the core's RigidBody as the MPM side uses it
// (src/rigid_transfer.cpp:32,66,72; src/transfer.cpp:365,428-445,577-580,763-771,829,956-969; src/boundary_particle.h:48-73).
// Unlike the rest of this directory it has BEHAVIOUR, because the coupled transfers read and write it; the real class cannot be
// read here, so this is the ASSUMED meaning, stated once (the same list stands in include/mpmb.h and oracle/mpm_oracle.cpp):
//   get_velocity_at(p)       = velocity + angular_velocity x (p - position)
//   apply_tmp_impulse(j, p)  : tmp_velocity += inv_mass j ; tmp_angular_velocity += inv_inertia ((p - position) x j)   (under a lock)
//   reset_tmp_velocity / apply_tmp_velocity : zero both accumulators / add them to velocity and angular_velocity
//   get_mesh_to_world() = get_centroid_to_world() = x -> position + rotation x (the mesh is re-centred on the centre of mass at
//                         creation, src/mpm_rigid_body.cpp:190-207)
//   Element<3>: three vertices; get_transformed(M) maps them; world_to_element(e) = [v1 - v0, v2 - v0, n]^-1, n the unit normal
// The body's own dynamics (advance, collisions, articulation) are not restated: the harness keeps the pose fixed over a substep.
#pragma once
#include <taichi/util.h>
#include <mutex>
namespace taichi {
template <int dim>
struct ElementOf {
  using Vector = VectorND<dim, real>;
  Vector v[dim];
  ElementOf get_transformed(const MatrixND<dim + 1, real> &m) const {
    ElementOf r;
    for (int k = 0; k < dim; k++) {
      VectorND<dim + 1, real> h;
      for (int i = 0; i < dim; i++) h[i] = v[k][i];
      h[dim] = 1;
      VectorND<dim + 1, real> t = m * h;
      for (int i = 0; i < dim; i++) r.v[k][i] = t[i];
    }
    return r;
  }
  Vector get_normal() const;
};
template <> inline VectorND<3, real> ElementOf<3>::get_normal() const { return normalized(cross(v[1] - v[0], v[2] - v[0])); }
template <> inline VectorND<2, real> ElementOf<2>::get_normal() const { VectorND<2, real> d = v[1] - v[0]; return normalized(VectorND<2, real>(d[1], -d[0])); }

inline MatrixND<3, real> world_to_element(const ElementOf<3> &e) {
  return inversed(MatrixND<3, real>(e.v[1] - e.v[0], e.v[2] - e.v[0], e.get_normal()));
}
inline MatrixND<2, real> world_to_element(const ElementOf<2> &e) { return inversed(MatrixND<2, real>(e.v[1] - e.v[0], e.get_normal())); }

template <int dim> struct AngularVelocity;
template <> struct AngularVelocity<3> {
  using ValueType = VectorND<3, real>;
  ValueType value = ValueType(0.0f);
  AngularVelocity() {}
  AngularVelocity(const ValueType &v) : value(v) {}
  VectorND<3, real> cross(const VectorND<3, real> &r) const { return taichi::cross(value, r); }
};
template <> struct AngularVelocity<2> {
  using ValueType = real;
  ValueType value = 0;
  AngularVelocity() {}
  AngularVelocity(real v) : value(v) {}
  VectorND<2, real> cross(const VectorND<2, real> &r) const { return VectorND<2, real>(-value * r[1], value * r[0]); }
};
template <int dim> struct Rotation {
  MatrixND<dim, real> value = MatrixND<dim, real>(1.0f);
  Rotation() {}
  explicit Rotation(real) {}
  VectorND<dim, real> rotate(const VectorND<dim, real> &v) const { return value * v; }
};

template <int dim>
struct RigidBody {
  using Vector = VectorND<dim, real>;
  using ElementType = ElementOf<dim>;
  int id = 0;
  real frictions[2] = {0, 0};
  Vector position = Vector(0.0f), velocity = Vector(0.0f), tmp_velocity = Vector(0.0f);
  AngularVelocity<dim> angular_velocity, tmp_angular_velocity;
  Rotation<dim> rotation;
  real inv_mass = 0;
  MatrixND<3, real> inv_inertia = MatrixND<3, real>(0.0f);   // world space (3-D)
  std::mutex mut;
  int pos_func_id = -1, rot_func_id = -1;
  using PositionFunctionType = std::function<Vector(real)>;
  using RotationFunctionType = std::function<Vector(real)>;
  PositionFunctionType pos_func;
  RotationFunctionType rot_func;
  void set_as_background() {}
  struct MeshType { std::vector<ElementType> elements; void initialize(const Config &) {} };
  std::shared_ptr<MeshType> mesh;
  MatrixND<dim + 1, real> get_centroid_to_world() const {
    MatrixND<dim + 1, real> m(1.0f);
    for (int c = 0; c < dim; c++) for (int r = 0; r < dim; r++) m[c][r] = rotation.value[c][r];
    for (int r = 0; r < dim; r++) m[dim][r] = position[r];
    return m;
  }
  MatrixND<dim + 1, real> get_mesh_to_world() const { return get_centroid_to_world(); }
  void reset_tmp_velocity() { tmp_velocity = Vector(0.0f); tmp_angular_velocity = AngularVelocity<dim>(); }
  void apply_tmp_velocity() { velocity = velocity + tmp_velocity; angular_velocity.value = angular_velocity.value + tmp_angular_velocity.value; }
  Vector get_velocity_at(const Vector &p) const { return velocity + angular_velocity.cross(p - position); }
  void apply_tmp_impulse(const Vector &j, const Vector &p);
};
template <> inline void RigidBody<3>::apply_tmp_impulse(const Vector &j, const Vector &p) {
  std::lock_guard<std::mutex> _(mut);
  tmp_velocity = tmp_velocity + inv_mass * j;
  tmp_angular_velocity.value = tmp_angular_velocity.value + inv_inertia * cross(p - position, j);
}
template <> inline void RigidBody<2>::apply_tmp_impulse(const Vector &, const Vector &) {}
template <int n> inline VectorND<n, real> transform(const MatrixND<n + 1, real> &m, const VectorND<n, real> &v) {
  VectorND<n + 1, real> h;
  for (int i = 0; i < n; i++) h[i] = v[i];
  h[n] = 1;
  VectorND<n + 1, real> t = m * h;
  VectorND<n, real> r;
  for (int i = 0; i < n; i++) r[i] = t[i];
  return r;
}
template <class T> inline void trash(T &&) {}
}  // namespace taichi
