// TEST INFRASTRUCTURE.  The reference's OWN solver core, every file included where it lies and unmodified:
//   src/transfer.cpp   P2G (rasterize / rasterize_optimized) and G2P (resample / resample_optimized)
//   src/mpm.cpp        ordering and page maps, grid normalisation, level-set boundary condition, boundary
//                      deletion, MPM<3>::substep() and step()
//   src/visualize.cpp  write_partio (frame dump through the vendored Partio)
//   src/particles.cpp  the registered particle types (constitutive models)
//   src/mpm.h, particle_allocator.h, kernel.h, mpm_fwd.h, articulation.h, boundary_particle.h,
//   poisson_disk_sampler.h, external/SPGrid, external/partio
// compiled against the stand-in core headers oracle/taichi_stub/taichi/*.h (their header says what they restate),
// so that the oracle's restatement is pinned against the reference's lines executed here — stage by stage and as
// whole substeps.
//
// What THIS file adds is scaffolding only:
//   * empty bodies for the solver members defined in translation units outside this build (rigid coupling:
//     src/rigid_transfer.cpp etc.) — none is reachable on the pinned path; add_particles is the reference's own
//     (its texture / mesh / Poisson-disk branches only compile; reft_add_benchmark drives its `benchmark` lattice
//     branch, src/mpm.cpp:155-186, everything else loads particles directly);
//   * populate(): the ordering / page maps / per-node counts of sort_particles_and_populate_grid
//     (src/mpm.cpp:770-918) restated with the real SPGrid calls, used only by the single-transfer entry points
//     (reft_p2g / reft_g2p); reft_substep runs the reference's own sort_particles_and_populate_grid;
//   * C entry points that load particles, set a plane level set, run one stage or whole substeps, read or write
//     node values, and dump a frame.
// standard / stand-in / Partio headers first, so that the `private` trick below only touches the reference's classes
#include <taichi/util.h>
#include <taichi/stub_more.h>
#include <taichi/math/svd.h>
#include <taichi/dynamics/simulation.h>
#include <taichi/dynamics/rigid_body.h>
#include <Partio.h>
#include <array>
#include <atomic>
#include <bitset>
#include <chrono>
#include <iostream>
#include <set>
#include <sstream>
#include <thread>
#include <unordered_map>
#include <sys/mman.h>
#include <SPGrid/Core/SPGrid_Allocator.h>
#include <SPGrid/Core/SPGrid_Page_Map.h>
#define private public   // the integration patch lives inside the reference's classes: offsetof(v_and_m)
#include REF_TRANSFER_SOURCE
#include REF_MPM_SOURCE
#include REF_VISUALIZE_SOURCE
#include REF_PARTICLES_SOURCE
#include REF_RIGID_TRANSFER_SOURCE       // rasterize_rigid_boundary, gather_cdf (src/rigid_transfer.cpp)
#include REF_BOUNDARY_PARTICLE_SOURCE    // registers "rigid_boundary" (src/boundary_particle.cpp)
// AsyncMPM (src/async/async_mpm.{h,cpp}): the scheduler calls MPM<dim>::substep() for every time level (advance(), :329;
// step(), :377).  Those calls are the patch point of the drop-in, and the reference's sources are not
// edited, so within this one include `substep()` expands to a harness hook that either runs the reference's own
// MPM<3>::substep() or hands the pool to libmpmb; `class` -> `struct` only opens AsyncMPM's leading private section to the
// harness (it sets the fields AsyncMPM::initialize would, since MPM::initialize needs the real core's Config).
namespace taichi {
void harness_substep(MPM<3> *m);
inline void harness_substep(MPM<2> *) {}
}  // namespace taichi
#define substep() res; ::taichi::harness_substep(this)
#define class struct
#include REF_ASYNC_SOURCE
#undef class
#undef substep
#undef private
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <dlfcn.h>
#include "../include/mpmb.h"

namespace taichi {
// members of the solver that live in translation units which are not part of this build (src/mpm_rigid_body.cpp):
// empty — rigid bodies keep their pose over a substep here (the harness sets it), so rigidify / advect_rigid_bodies have
// nothing to do; src/rigid_transfer.cpp IS part of the build
template <> void MPM<3>::add_rigid_particle(Config) {}
template <> void MPM<3>::rigidify(real) {}
template <> void MPM<3>::advect_rigid_bodies(real) {}
template <> void MPM<3>::rigid_body_levelset_collision(real, real) {}
template <> void AsyncMPM<3>::visualize() const {}   // src/async/async_visualize.cpp (debug images of the time levels) is not part of the build
// src/mpm.cpp instantiates parts of the 2-D solver explicitly (general_action); the same members, never called
template <> void MPM<2>::rigidify(real) {}
template <> void MPM<2>::advect_rigid_bodies(real) {}
template <> void MPM<2>::rigid_body_levelset_collision(real, real) {}
}  // namespace taichi

namespace {
using namespace taichi;
using Solver = MPM<3>;
using Mask = Solver::SparseMask;

struct Harness {
  std::unique_ptr<Solver> owner;   // MPM<3>, or AsyncMPM<3> (its scheduler calls MPM<3>::substep through harness_substep)
  Solver &m;
  std::vector<int> kind;           // per allocator slot
  // route of harness_substep: empty = the reference's own MPM<3>::substep(); else libmpmb through the C-ABI
  std::string route_lib;
  std::string route_error;
  int64_t routed_substeps = 0;
  MpmbHandle route_engine = nullptr;          // kept from one routed substep to the next: the step changes through mpmb_set_delta_t
  int (*route_destroy)(MpmbHandle) = nullptr;
  explicit Harness(bool async = false) : owner(async ? static_cast<Solver *>(new AsyncMPM<3>()) : new Solver()), m(*owner) {}
};
std::unordered_map<Solver *, Harness *> &harness_registry() { static std::unordered_map<Solver *, Harness *> r; return r; }

// src/mpm.cpp:770-918 with std::sort instead of tbb::parallel_sort, no periodic pool re-pack
void populate(Solver &m) {
  constexpr int index_bits = 32 - Mask::block_bits;
  const size_t n = m.particles.size();
  m.particle_sorter.resize(n);
  auto grid_array = m.grid->Get_Array();
  for (size_t i = 0; i < n; i++) {
    uint64 offset = Mask::Linear_Offset(to_std_array(m.get_grid_base_pos(m.allocator[m.particles[i]]->pos * m.inv_delta_x)));
    m.particle_sorter[i] = ((offset >> Mask::data_bits) << index_bits) + i;
  }
  std::sort(m.particle_sorter.begin(), m.particle_sorter.end());
  std::swap(m.particles, m.particles_);
  m.particles.resize(m.particles_.size());
  for (size_t i = 0; i < n; i++) m.particles[i] = m.particles_[m.particle_sorter[i] & ((1ll << index_bits) - 1)];
  m.page_map->Clear();
  for (size_t i = 0; i < n; i++) m.page_map->Set_Page((m.particle_sorter[i] >> index_bits) << Mask::data_bits);
  m.page_map->Update_Block_Offsets();
  auto blocks = m.page_map->Get_Blocks();
  m.fat_page_map->Clear();
  for (int b = 0; b < (int)blocks.second; b++) {
    auto base_offset = blocks.first[b];
    auto x = 1 << Mask::block_xbits, y = 1 << Mask::block_ybits, z = 1 << Mask::block_zbits;
    auto c = Mask::LinearToCoord(base_offset);
    for (int i = -1 + (c[0] == 0); i < 2; i++)
      for (int j = -1 + (c[1] == 0); j < 2; j++)
        for (int k = -1 + (c[2] == 0); k < 2; k++)
          m.fat_page_map->Set_Page(Mask::Packed_Add(base_offset, Mask::Linear_Offset(x * i, y * j, z * k)));
  }
  m.fat_page_map->Update_Block_Offsets();
  auto fat_blocks = m.fat_page_map->Get_Blocks();
  for (int i = 0; i < (int)fat_blocks.second; i++) std::memset(&grid_array(fat_blocks.first[i]), 0, 1 << Solver::log2_size);
  m.block_meta.clear();
  uint64 last_offset = -1;
  for (uint32 i = 0; i < n; i++) {
    if (last_offset != (m.particle_sorter[i] >> 32)) m.block_meta.push_back({i, 0});
    last_offset = m.particle_sorter[i] >> 32;
  }
  m.block_meta.push_back({(uint32)n, 0});
  for (int b = 0; b < (int)blocks.second; b++) {
    GridState<3> *g = reinterpret_cast<GridState<3> *>(&grid_array(blocks.first[b]));
    for (uint32 i = m.block_meta[b].particle_offset; i < m.block_meta[b + 1].particle_offset; i++) {
      auto base_pos = m.get_grid_base_pos(m.allocator[m.particles[i]]->pos * m.inv_delta_x);
      uint64 offset = Mask::Linear_Offset(to_std_array(base_pos));
      g[(offset >> Mask::data_bits) & ((1 << Mask::block_bits) - 1)].particle_count += 1;
    }
  }
  m.rigid_page_map->Clear();  // no rigid bodies: every block takes block_op_normal (src/transfer.cpp:570-575)
  m.rigid_page_map->Update_Block_Offsets();
}

MatrixND<3, real> load(const float *a) {
  MatrixND<3, real> r;
  for (int c = 0; c < 3; c++) for (int q = 0; q < 3; q++) r[c][q] = a[c * 3 + q];
  return r;
}
void store(const MatrixND<3, real> &a, float *o) {
  for (int c = 0; c < 3; c++) for (int q = 0; q < 3; q++) o[c * 3 + q] = a[c][q];
}
}  // namespace

extern "C" {
void *reft_create(const int *res, float dx, float dt, const float *gravity, int particle_gravity) {
  Harness *h = new Harness();
  Solver &m = h->m;
  m.res = VectorND<3, int>(res[0], res[1], res[2]);
  m.delta_x = dx;
  m.inv_delta_x = 1.0f / dx;
  m.base_delta_t = dt;
  m.gravity = VectorND<3, real>(gravity[0], gravity[1], gravity[2]);
  m.particle_gravity = particle_gravity != 0;
  m.apic = true;
  m.apic_damping = m.rpic_damping = m.affine_damping = m.penalty = 0;
  m.pushing_force = 20000.0f;
  m.cutting_counter = m.plasticity_counter = 0;
  m.reorder_interval = 1000;                               // src/mpm.cpp:45
  m.spgrid_size = 4096;                                   // src/mpm.cpp:50-54
  while (m.spgrid_size / 2 > (m.res.max() + 1)) m.spgrid_size /= 2;
  m.grid = std::make_unique<Solver::SparseGrid>(m.spgrid_size, m.spgrid_size, m.spgrid_size);
  m.page_map = std::make_unique<Solver::PageMap>(*m.grid);
  m.rigid_page_map = std::make_unique<Solver::PageMap>(*m.grid);
  m.fat_page_map = std::make_unique<Solver::PageMap>(*m.grid);
  return h;
}
void reft_destroy(void *hp) {
  Harness *h = static_cast<Harness *>(hp);
  harness_registry().erase(&h->m);
  if (h->route_engine && h->route_destroy) h->route_destroy(h->route_engine);
  delete h;
}
// threads for the stand-in's parallel loops (ThreadedTaskManager / tbb): 1 = serial (the pins); more only to time
void reft_set_threads(void *hp, int n) {
  static_cast<Harness *>(hp)->m.num_threads = n < 1 ? 1 : n;
  stub_num_threads() = n < 1 ? 1 : n;
#if defined(_OPENMP)
  omp_set_num_threads(n < 1 ? 1 : n);
#endif
}
// bulk loader: n particles of one kind (same arguments as reft_add_particle, arrays)
int64_t reft_add_particles(void *hp, int kind, const float *params, int64_t n, const float *x, const float *v, const float *mass,
                           const float *vol, const float *F, const float *b, const float *ps);

// kind / params: the oracle's numbering and parameter vectors.  Returns the particle's id.
int reft_add_particle(void *hp, int kind, const float *params, const float *x, const float *v, float mass, float vol, const float *F,
                      const float *b, float ps) {
  Harness *h = static_cast<Harness *>(hp);
  static const char *names[8] = {"linear", "jelly", "snow", "water", "sand", "elastic", "von_mises", "visco"};
  if (kind < 0 || kind > 7) return -1;
  auto alloc = h->m.allocator.allocate_particle(names[kind]);
  MPMParticle<3> *p = alloc.second;
  Config cfg;
  switch (kind) {
    case 0: p->initialize(cfg); static_cast<LinearParticle<3> *>(p)->mu = params[0]; static_cast<LinearParticle<3> *>(p)->lambda = params[1]; break;
    case 1: p->initialize(cfg); static_cast<JellyParticle<3> *>(p)->mu = params[0]; static_cast<JellyParticle<3> *>(p)->lambda = params[1]; break;
    case 2:
      cfg.set("mu_0", params[0]).set("lambda_0", params[1]).set("hardening", params[2]).set("theta_c", params[3]).set("theta_s", params[4])
          .set("min_Jp", params[5]).set("max_Jp", params[6]).set("Jp", ps);
      p->initialize(cfg);
      break;
    case 3: cfg.set("k", params[0]).set("gamma", params[1]); p->initialize(cfg); static_cast<WaterParticle<3> *>(p)->j = ps; break;
    case 4:
      cfg.set("mu_0", params[0]).set("lambda_0", params[1]).set("cohesion", params[3]).set("beta", params[4]);
      p->initialize(cfg);
      static_cast<SandParticle<3> *>(p)->alpha = params[2];
      static_cast<SandParticle<3> *>(p)->logJp = ps;
      break;
    case 5: p->initialize(cfg); static_cast<ElasticParticle<3> *>(p)->mu_0 = params[0]; static_cast<ElasticParticle<3> *>(p)->lambda_0 = params[1]; break;
    case 6:
      cfg.set("yield_stress", params[2]);
      p->initialize(cfg);
      static_cast<VonMisesParticle<3> *>(p)->mu_0 = params[0];
      static_cast<VonMisesParticle<3> *>(p)->lambda_0 = params[1];
      break;
    case 7:
      cfg.set("nu", params[2]).set("kappa", params[3]).set("base_delta_t", params[4]).set("tau", ps);
      p->initialize(cfg);
      static_cast<ViscoParticle<3> *>(p)->mu_0 = params[0];
      static_cast<ViscoParticle<3> *>(p)->lambda_0 = params[1];
      break;
  }
  p->pos = VectorND<3, real>(x[0], x[1], x[2]);
  p->set_mass(mass);
  p->set_velocity(VectorND<3, real>(v[0], v[1], v[2]));
  p->vol = vol;
  p->dg_e = load(F);
  p->apic_b = load(b);
  h->m.particles.push_back(alloc.first);
  h->kind.push_back(kind);
  return (int)p->id;
}

int64_t reft_add_particles(void *hp, int kind, const float *params, int64_t n, const float *x, const float *v, const float *mass,
                           const float *vol, const float *F, const float *b, const float *ps) {
  for (int64_t i = 0; i < n; i++)
    if (reft_add_particle(hp, kind, params, x + 3 * i, v + 3 * i, mass[i], vol[i], F + 9 * i, b + 9 * i, ps[i]) < 0) return -1;
  return n;
}

// P2G of one substep: ordering + cleared fat blocks, then rasterize(dt) (scalar, src/transfer.cpp:193-278)
// or rasterize_optimized(dt) (src/transfer.cpp:361-581)
void reft_p2g(void *hp, int optimized) {
  Solver &m = static_cast<Harness *>(hp)->m;
  populate(m);
  if (optimized) m.rasterize_optimized(m.base_delta_t);
  else m.rasterize(m.base_delta_t, true);
}
// dense [(res+1)^3][4] copy of velocity_and_mass (momentum, mass after P2G)
void reft_get_grid(void *hp, float *out) {
  Solver &m = static_cast<Harness *>(hp)->m;
  const int nx = m.res[0] + 1, ny = m.res[1] + 1, nz = m.res[2] + 1;
  auto fat = m.fat_page_map->Get_Blocks();
  std::memset(out, 0, sizeof(float) * 4 * size_t(nx) * ny * nz);
  auto grid_array = m.grid->Get_Array();
  for (int b = 0; b < (int)fat.second; b++) {
    auto c = Mask::LinearToCoord(fat.first[b]);
    for (int i = 0; i < (1 << Mask::block_xbits); i++)
      for (int j = 0; j < (1 << Mask::block_ybits); j++)
        for (int k = 0; k < (1 << Mask::block_zbits); k++) {
          int X = c[0] + i, Y = c[1] + j, Z = c[2] + k;
          if (X >= nx || Y >= ny || Z >= nz) continue;
          const GridState<3> &g = grid_array(std::array<int, 3>{X, Y, Z});
          float *o = out + 4 * ((size_t(X) * ny + Y) * nz + Z);
          for (int q = 0; q < 4; q++) o[q] = g.velocity_and_mass[q];
        }
  }
}
// overwrite velocity_and_mass of every node of the fat blocks from a dense array (node velocities before G2P)
void reft_set_grid(void *hp, const float *in) {
  Solver &m = static_cast<Harness *>(hp)->m;
  const int nx = m.res[0] + 1, ny = m.res[1] + 1, nz = m.res[2] + 1;
  auto fat = m.fat_page_map->Get_Blocks();
  auto grid_array = m.grid->Get_Array();
  for (int b = 0; b < (int)fat.second; b++) {
    auto c = Mask::LinearToCoord(fat.first[b]);
    for (int i = 0; i < (1 << Mask::block_xbits); i++)
      for (int j = 0; j < (1 << Mask::block_ybits); j++)
        for (int k = 0; k < (1 << Mask::block_zbits); k++) {
          int X = c[0] + i, Y = c[1] + j, Z = c[2] + k;
          if (X >= nx || Y >= ny || Z >= nz) continue;
          GridState<3> &g = grid_array(std::array<int, 3>{X, Y, Z});
          const float *o = in + 4 * ((size_t(X) * ny + Y) * nz + Z);
          for (int q = 0; q < 4; q++) g.velocity_and_mass[q] = o[q];
        }
  }
}
// G2P of one substep: resample() (scalar, src/transfer.cpp:585-687) or resample_optimized() (702-968)
void reft_g2p(void *hp, int optimized) {
  Solver &m = static_cast<Harness *>(hp)->m;
  if (optimized) m.resample_optimized();
  else m.resample();
}
// static level set of half-spaces, grid units: phi(X) = n.X + d (planes4[k] = nx, ny, nz, d), friction as set_friction
void reft_set_planes(void *hp, int n, const float *planes4, float friction) {
  Solver &m = static_cast<Harness *>(hp)->m;
  auto ls = std::make_shared<LevelSet<3>>();
  ls->friction = friction;
  for (int k = 0; k < n; k++) ls->planes.push_back(VectorND<4, real>(planes4[4 * k], planes4[4 * k + 1], planes4[4 * k + 2], planes4[4 * k + 3]));
  m.levelset.levelset0 = ls;
}
// the grid update between the transfers, as substep() does it (src/mpm.cpp:519-540): normalisation (+ gravity on
// the grid when particle_gravity is off), then the level-set boundary condition
void reft_grid_update(void *hp) {
  Solver &m = static_cast<Harness *>(hp)->m;
  VectorND<3, real> inc = m.gravity * m.base_delta_t;
  if (m.particle_gravity) inc = VectorND<3, real>(0.0f);
  m.normalize_grid_and_apply_external_force(inc);
  if (m.levelset.levelset0) m.apply_grid_boundary_conditions(m.levelset, m.current_t);
}
// n whole substeps by MPM<3>::substep() itself (src/mpm.cpp:452-575): its own ordering, optimized transfers,
// grid update and boundary deletion.  Returns the number of live particles.
int64_t reft_substep(void *hp, int n) {
  Solver &m = static_cast<Harness *>(hp)->m;
  if (!m.levelset.levelset0) m.levelset.levelset0 = std::make_shared<LevelSet<3>>();  // no planes: phi = +inf everywhere
  for (int i = 0; i < n; i++) m.substep();
  return (int64_t)m.particles.size();
}
// ---- AsyncMPM (src/async/async_mpm.{h,cpp}), SURVEY §8f row 3: the reference's own scheduler object.  reft_create_async sets
// what AsyncMPM<3>::initialize sets (src/async/async_mpm.cpp:13-58) after the solver fields of reft_create; particles are loaded
// with reft_add_particle(s) and then handed to the block pools (the tail of AsyncMPM::add_particles, :62-76).
void *reft_create_async(const int *res, float dx, const float *gravity, int particle_gravity, float unit_delta_t, int64_t max_units,
                        float cfl_dt_mul, float strength_dt_mul) {
  Harness *h = new Harness(true);
  auto &m = *static_cast<AsyncMPM<3> *>(h->owner.get());
  m.res = VectorND<3, int>(res[0], res[1], res[2]);
  m.delta_x = dx;
  m.inv_delta_x = 1.0f / dx;
  m.base_delta_t = unit_delta_t;
  m.gravity = VectorND<3, real>(gravity[0], gravity[1], gravity[2]);
  m.particle_gravity = particle_gravity != 0;
  m.apic = true;
  m.apic_damping = m.rpic_damping = m.affine_damping = m.penalty = 0;
  m.pushing_force = 20000.0f;
  m.cutting_counter = m.plasticity_counter = 0;
  m.reorder_interval = 0;                                 // AsyncMPM keeps id == pool slot (gather_from_pool asserts it)
  m.spgrid_size = 4096;
  while (m.spgrid_size / 2 > (m.res.max() + 1)) m.spgrid_size /= 2;
  m.grid = std::make_unique<Solver::SparseGrid>(m.spgrid_size, m.spgrid_size, m.spgrid_size);
  m.page_map = std::make_unique<Solver::PageMap>(*m.grid);
  m.rigid_page_map = std::make_unique<Solver::PageMap>(*m.grid);
  m.fat_page_map = std::make_unique<Solver::PageMap>(*m.grid);
  m.levelset.levelset0 = std::make_shared<LevelSet<3>>();
  // AsyncMPM<3>::initialize, src/async/async_mpm.cpp:16-58
  {
    auto bs = m.grid_block_size();
    m.scheduler_size = (uint64)m.spgrid_size * m.spgrid_size * m.spgrid_size / (uint64)(bs[0] * bs[1] * bs[2]);
  }
  m.scheduler_mask = m.scheduler_size - 1;
  m.unit_delta_t = unit_delta_t;
  m.max_units = max_units;
  m.cfl_dt_mul = cfl_dt_mul;
  m.strength_dt_mul = strength_dt_mul;
  m.blocks.resize(m.scheduler_size);
  std::memset(&m.blocks[0], 0, sizeof(m.blocks[0]) * m.blocks.size());
  for (uint64 o = 0; o < m.scheduler_size; ++o) {
    m.blocks[o].strength_dt_limit = 1LL << 31;
    m.blocks[o].cfl_dt_limit = 1LL << 31;
    m.blocks[o].continuous_dt_limit = 1;
    m.blocks[o].local_min_dt_limit = 1;
  }
  m.precompute_neighbor_pairs();
  harness_registry()[&h->m] = h;
  return h;
}
// the tail of AsyncMPM::add_particles (src/async/async_mpm.cpp:64-75): every loaded particle into the pool of its block
void reft_async_distribute(void *hp) {
  Harness *h = static_cast<Harness *>(hp);
  auto &m = *static_cast<AsyncMPM<3> *>(h->owner.get());
  for (auto p : m.particles) {
    uint64 grid_offset = Mask::Linear_Offset(to_std_array(m.get_grid_base_pos(m.allocator[p]->pos * m.inv_delta_x)));
    uint64 offset = (grid_offset >> Mask::data_bits >> Mask::block_bits) & m.scheduler_mask;
    m.particle_pool[offset].push_back(m.allocator.pool[p]);
  }
  m.particles.clear();
}
// AsyncMPM<3>::step(dt) itself (src/async/async_mpm.cpp:375-421).  out[4] = {update_counter, current_t_int, min and max
// continuous_dt_limit over the non-empty blocks (the spread of time levels the scheduler chose)}
int64_t reft_async_step(void *hp, float dt, int64_t *out) {
  Harness *h = static_cast<Harness *>(hp);
  auto &m = *static_cast<AsyncMPM<3> *>(h->owner.get());
  h->route_error.clear();
  m.step(dt);
  int64_t lo = 1LL << 40, hi = 0, n = 0;
  for (uint64 o = 0; o < m.scheduler_size; ++o)
    if (!m.particle_pool[o].empty()) {
      lo = std::min<int64_t>(lo, m.blocks[o].continuous_dt_limit);
      hi = std::max<int64_t>(hi, m.blocks[o].continuous_dt_limit);
      n += (int64_t)m.particle_pool[o].size();
    }
  if (out) { out[0] = (int64_t)m.update_counter; out[1] = m.current_t_int; out[2] = lo; out[3] = hi; }
  return h->route_error.empty() ? n : -1;
}
const char *reft_route_error(void *hp) { return static_cast<Harness *>(hp)->route_error.c_str(); }
// particle state by id out of the block pools (every live particle sits in exactly one particle_pool)
int64_t reft_async_get_particles(void *hp, float *x, float *v, float *F, float *b, float *ps, uint8_t *alive) {
  Harness *h = static_cast<Harness *>(hp);
  auto &m = *static_cast<AsyncMPM<3> *>(h->owner.get());
  int64_t n = 0;
  for (uint64 o = 0; o < m.scheduler_size; ++o)
    for (const auto &c : m.particle_pool[o]) {
      const MPMParticle<3> *p = reinterpret_cast<const MPMParticle<3> *>(&c);
      const int id = p->id;
      for (int d = 0; d < 3; d++) { x[3 * id + d] = p->pos[d]; v[3 * id + d] = p->get_velocity()[d]; }
      store(p->dg_e, F + 9 * id);
      store(p->apic_b, b + 9 * id);
      switch (h->kind[id]) {
        case 2: ps[id] = static_cast<const SnowParticle<3> *>(p)->Jp; break;
        case 3: ps[id] = static_cast<const WaterParticle<3> *>(p)->j; break;
        case 4: ps[id] = static_cast<const SandParticle<3> *>(p)->logJp; break;
        case 7: ps[id] = static_cast<const ViscoParticle<3> *>(p)->visco_tau; break;
        default: ps[id] = 0;
      }
      alive[id] = 1;
      n++;
    }
  return n;
}
// from now on every MPM<3>::substep() the scheduler asks for goes to libmpmb (lib_path; empty string: back to the reference)
void reft_route_through(void *hp, const char *lib_path) { static_cast<Harness *>(hp)->route_lib = lib_path ? lib_path : ""; }
int64_t reft_routed_substeps(void *hp) { return static_cast<Harness *>(hp)->routed_substeps; }

// ---- rigid bodies (CPIC).  Bodies 1..n_bodies-1 become MPM::rigids[1..] (row 0 is the background body MPM::initialize
// creates, src/mpm.cpp:72-74); every sample becomes a RigidBoundaryParticle in the reference's own pool
// (add_boundry_particle, src/mpm_rigid_body.cpp:153-169), aligned with its body.  Call after the MPM particles are loaded.
int reft_set_rigid(void *hp, int n_bodies, const float *position, const float *rot, const float *velocity, const float *angular_velocity,
                   const float *inv_mass, const float *inv_inertia, const float *frictions, int64_t n_samples, const float *offset,
                   const float *tri, const int32_t *sample_rigid, float penalty, float pushing_force) {
  Harness *h = static_cast<Harness *>(hp);
  Solver &m = h->m;
  if (!m.rigids.empty()) return -1;
  using V3 = VectorND<3, real>;
  for (int b = 0; b < n_bodies; b++) {
    auto r = std::make_unique<RigidBody<3>>();
    r->id = b;
    r->position = V3(position[3 * b], position[3 * b + 1], position[3 * b + 2]);
    r->velocity = V3(velocity[3 * b], velocity[3 * b + 1], velocity[3 * b + 2]);
    r->angular_velocity.value = V3(angular_velocity[3 * b], angular_velocity[3 * b + 1], angular_velocity[3 * b + 2]);
    r->rotation.value = load(rot + 9 * b);
    r->inv_mass = inv_mass[b];
    r->inv_inertia = load(inv_inertia + 9 * b);
    r->frictions[0] = frictions[2 * b];
    r->frictions[1] = frictions[2 * b + 1];
    m.rigids.push_back(std::move(r));
  }
  m.penalty = penalty;
  m.pushing_force = pushing_force;
  m.config_backup.set("rigid_body_collision", false);   // rigidify() is outside this build anyway
  for (int64_t s = 0; s < n_samples; s++) {
    auto alloc = m.allocator.allocate_particle("rigid_boundary");
    auto *p = static_cast<RigidBoundaryParticle<3> *>(alloc.second);
    p->rigid = m.rigids[sample_rigid[s]].get();
    p->offset = V3(offset[3 * s], offset[3 * s + 1], offset[3 * s + 2]);
    for (int k = 0; k < 3; k++) p->untransformed_element.v[k] = V3(tri[9 * s + 3 * k], tri[9 * s + 3 * k + 1], tri[9 * s + 3 * k + 2]);
    p->original_normal = p->untransformed_element.get_normal();
    p->set_mass(0.0f);
    p->align_with_rigid_body();
    m.particles.push_back(alloc.first);
    h->kind.push_back(-1);
  }
  return 0;
}
// pose and velocities of every body for the next substep; the boundary particles follow (align_with_rigid_body,
// what advect_rigid_bodies does at the end of a substep, src/mpm_rigid_body.cpp:278-283)
void reft_set_rigid_state(void *hp, const float *position, const float *rot, const float *velocity, const float *angular_velocity) {
  Solver &m = static_cast<Harness *>(hp)->m;
  using V3 = VectorND<3, real>;
  for (size_t b = 0; b < m.rigids.size(); b++) {
    auto &r = *m.rigids[b];
    r.position = V3(position[3 * b], position[3 * b + 1], position[3 * b + 2]);
    r.velocity = V3(velocity[3 * b], velocity[3 * b + 1], velocity[3 * b + 2]);
    r.angular_velocity.value = V3(angular_velocity[3 * b], angular_velocity[3 * b + 1], angular_velocity[3 * b + 2]);
    r.rotation.value = load(rot + 9 * b);
  }
  for (auto ptr : m.particles) {
    MPMParticle<3> *p = m.allocator[ptr];
    if (p->is_rigid()) static_cast<RigidBoundaryParticle<3> *>(p)->align_with_rigid_body();
  }
}
void reft_get_rigid_state(void *hp, float *velocity, float *angular_velocity) {
  Solver &m = static_cast<Harness *>(hp)->m;
  for (size_t b = 0; b < m.rigids.size(); b++)
    for (int d = 0; d < 3; d++) { velocity[3 * b + d] = m.rigids[b]->velocity[d]; angular_velocity[3 * b + d] = m.rigids[b]->angular_velocity.value[d]; }
}
// MPMParticle::states in (by id), states / boundary_normal / boundary_distance / near_boundary_ out (by id)
void reft_set_states(void *hp, const uint32_t *states) {
  Solver &m = static_cast<Harness *>(hp)->m;
  for (auto ptr : m.particles) { MPMParticle<3> *p = m.allocator[ptr]; if (!p->is_rigid()) p->states = states[p->id]; }
}
void reft_get_cdf_particles(void *hp, uint32_t *states, float *normal, float *dist, uint8_t *near) {
  Solver &m = static_cast<Harness *>(hp)->m;
  for (auto ptr : m.particles) {
    MPMParticle<3> *p = m.allocator[ptr];
    if (p->is_rigid()) continue;
    const int id = p->id;
    states[id] = p->states;
    for (int d = 0; d < 3; d++) normal[3 * id + d] = p->boundary_normal[d];
    dist[id] = p->boundary_distance;
    near[id] = p->near_boundary_ ? 1 : 0;
  }
}
// GridState::states (tags | (rigid id + 1) << 24) and GridState::distance of every node of the fat blocks, dense
void reft_get_cdf_grid(void *hp, uint32_t *states, float *dist) {
  Solver &m = static_cast<Harness *>(hp)->m;
  const int nx = m.res[0] + 1, ny = m.res[1] + 1, nz = m.res[2] + 1;
  auto fat = m.fat_page_map->Get_Blocks();
  auto grid_array = m.grid->Get_Array();
  for (int b = 0; b < (int)fat.second; b++) {
    auto c = Mask::LinearToCoord(fat.first[b]);
    for (int i = 0; i < (1 << Mask::block_xbits); i++)
      for (int j = 0; j < (1 << Mask::block_ybits); j++)
        for (int k = 0; k < (1 << Mask::block_zbits); k++) {
          int X = c[0] + i, Y = c[1] + j, Z = c[2] + k;
          if (X >= nx || Y >= ny || Z >= nz) continue;
          const GridState<3> &g = grid_array(std::array<int, 3>{X, Y, Z});
          states[(size_t(X) * ny + Y) * nz + Z] = g.states;
          dist[(size_t(X) * ny + Y) * nz + Z] = g.distance;
        }
  }
}
// 1 if the block holding node (X,Y,Z) is a rigid page (update_rigid_page_map, src/mpm.cpp:1026-1076)
int reft_is_rigid_page(void *hp, int X, int Y, int Z) {
  Solver &m = static_cast<Harness *>(hp)->m;
  return m.rigid_page_map->Test_Page(Mask::Linear_Offset(std::array<int, 3>{X, Y, Z})) ? 1 : 0;
}
// one coupled substep in stages, exactly the calls MPM<3>::substep makes between rigidify and advect_rigid_bodies
// (src/mpm.cpp:464-565), so that the grid can be read after each transfer:
//   stage 0: sort_particles_and_populate_grid (with update_rigid_page_map), rasterize_rigid_boundary, gather_cdf
//   stage 1: rasterize_optimized            stage 2: grid update            stage 3: resample_optimized + clear_boundary_particles
void reft_coupled_stage(void *hp, int stage) {
  Solver &m = static_cast<Harness *>(hp)->m;
  if (!m.levelset.levelset0) m.levelset.levelset0 = std::make_shared<LevelSet<3>>();
  if (stage == 0) { m.sort_particles_and_populate_grid(); m.rasterize_rigid_boundary(); m.gather_cdf(); }
  if (stage == 1) m.rasterize_optimized(m.base_delta_t);
  if (stage == 2) reft_grid_update(hp);
  if (stage == 3) { m.resample_optimized(); m.clear_boundary_particles(); m.current_t += m.base_delta_t; m.substep_counter += 1; }
}
// ids of the live particles, in the solver's current order
void reft_alive_ids(void *hp, int32_t *ids) {
  Solver &m = static_cast<Harness *>(hp)->m;
  size_t o = 0;
  for (size_t k = 0; k < m.particles.size(); k++)
    if (!m.allocator[m.particles[k]]->is_rigid()) ids[o++] = m.allocator[m.particles[k]]->id;
}
// the reference's own benchmark seeding (add_particles with benchmark = 125 | 8000, src/mpm.cpp:155-186): a cube of
// res*0.2 (resp. res*0.8) cells per axis, 8 particles per cell.  type = registered particle name.  Returns the count.
int64_t reft_add_benchmark(void *hp, const char *type, int benchmark, float density) {
  Harness *h = static_cast<Harness *>(hp);
  Config cfg;
  cfg.set("type", std::string(type)).set("benchmark", benchmark).set("density", density);
  const size_t before = h->m.particles.size();
  h->m.add_particles(cfg);
  static const char *names[5] = {"linear", "jelly", "snow", "water", "sand"};
  int kind = -1;
  for (int k = 0; k < 5; k++) if (std::string(type) == names[k]) kind = k;
  for (size_t i = before; i < h->m.particles.size(); i++) h->kind.push_back(kind);
  return (int64_t)(h->m.particles.size() - before);
}
// mass and volume by id (what add_particles assigned)
void reft_get_mass_vol(void *hp, float *mass, float *vol) {
  Solver &m = static_cast<Harness *>(hp)->m;
  for (auto ptr : m.particles) { MPMParticle<3> *p = m.allocator[ptr]; mass[p->id] = p->get_mass(); vol[p->id] = p->vol; }
}
// ---- the drop-in, executed: INTEGRATION.md §2.  The reference's own MPM<3> object hands its AoS particle pool to
// libmpmb through the C-ABI (slot layout by offsetof on the reference's own types), the engine runs `n` substeps,
// and the pool and the `particles` index vector are refreshed from the device — after which the reference goes on
// (substep(), write_partio, ...) as if it had stepped itself.  `lib_path`: libmpmb.so (B200) or its SIMT-emulator
// build (tests/simt); loaded with dlopen so that this checker library does not link the product.
// Single material per call (MpmbAosLayout carries one plastic-scalar offset).  Returns survivors, or a negative status.
int64_t reft_substep_via_mpmb(void *hp, const char *lib_path, int n, char *err, int err_len) {
  Harness *h = static_cast<Harness *>(hp);
  Solver &m = h->m;
  auto fail = [&](const char *msg, int code) -> int64_t { if (err && err_len > 0) { std::strncpy(err, msg, err_len - 1); err[err_len - 1] = 0; } return code; };
  if (m.particles.empty()) return 0;
  void *lib = dlopen(lib_path, RTLD_NOW | RTLD_LOCAL);
  if (!lib) return fail(dlerror(), -100);
#define SYM(name) auto name##_ = reinterpret_cast<decltype(&name)>(dlsym(lib, #name)); if (!name##_) return fail("missing symbol " #name, -101)
  SYM(mpmb_create); SYM(mpmb_destroy); SYM(mpmb_last_error); SYM(mpmb_set_material); SYM(mpmb_set_planes); SYM(mpmb_upload_aos);
  SYM(mpmb_substep); SYM(mpmb_download_aos); SYM(mpmb_set_rigid_samples); SYM(mpmb_set_rigid_coupling); SYM(mpmb_set_rigid_state);
  SYM(mpmb_get_rigid_state); SYM(mpmb_set_particle_states); SYM(mpmb_get_particle_cdf); SYM(mpmb_set_delta_t);
#undef SYM
  // with rigid bodies: the RigidBoundaryParticles leave the index vector for the duration — the engine
  // takes them as a sample list — and come back behind the survivors
  const bool coupled = m.rigids.size() > 1;
  std::vector<Solver::ParticlePtr> rigid_ptrs, mpm_ptrs;
  for (auto ptr : m.particles) (m.allocator[ptr]->is_rigid() ? rigid_ptrs : mpm_ptrs).push_back(ptr);
  m.particles = mpm_ptrs;
  if (m.particles.empty()) { m.particles.insert(m.particles.end(), rigid_ptrs.begin(), rigid_ptrs.end()); return 0; }
  MpmbConfig c{};
  for (int d = 0; d < 3; d++) { c.res[d] = m.res[d]; c.gravity[d] = m.gravity[d]; }
  c.dx = m.delta_x; c.dt = m.base_delta_t;
  c.particle_gravity = m.particle_gravity;
  c.clean_boundary = m.config_backup.get("clean_boundary", true);
  c.device = 0; c.world = 1;
  // a routed solver (AsyncMPM's scheduler) keeps ONE engine and changes its step before every substep, as the reference changes
  // base_delta_t (src/async/async_mpm.cpp:407-409); the one-shot drop-in test creates and destroys its own
  const bool keep = !h->route_lib.empty();
  MpmbHandle e = keep ? h->route_engine : nullptr;
  if (!e) {
    if (mpmb_create_(&c, &e) != MPMB_OK) return fail(mpmb_last_error_(nullptr), -102);
  } else if (mpmb_set_delta_t_(e, m.base_delta_t) != MPMB_OK) {
    return fail(mpmb_last_error_(e), -104);
  }
  if (keep) { h->route_engine = e; h->route_destroy = mpmb_destroy_; }
  auto mpmb_release = [&](MpmbHandle x) { if (!keep) mpmb_destroy_(x); };
  // one material group: kind and parameters read back from the first particle (all are of one registered type here)
  const int kind = h->kind[m.allocator[m.particles[0]]->id];
  float prm[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  using P3 = MPMParticle<3>;
  MpmbAosLayout L{};
  L.stride = (int32_t)sizeof(ParticleContainer<3>);
  L.off_pos = (int32_t)offsetof(P3, pos);
  L.off_v_and_m = (int32_t)offsetof(P3, v_and_m);
  L.off_dg_e = (int32_t)offsetof(P3, dg_e);
  L.off_apic_b = (int32_t)offsetof(P3, apic_b);
  L.col_pitch = (int32_t)sizeof(VectorND<3, real>);
  L.off_vol = (int32_t)offsetof(P3, vol);
  L.off_scalar = -1;
  P3 *p0 = m.allocator[m.particles[0]];
  switch (kind) {
    case 0: prm[0] = static_cast<LinearParticle<3> *>(p0)->mu; prm[1] = static_cast<LinearParticle<3> *>(p0)->lambda; break;
    case 1: prm[0] = static_cast<JellyParticle<3> *>(p0)->mu; prm[1] = static_cast<JellyParticle<3> *>(p0)->lambda; break;
    case 2: { auto *q = static_cast<SnowParticle<3> *>(p0); prm[0] = q->mu_0; prm[1] = q->lambda_0; prm[2] = q->hardening; prm[3] = q->theta_c; prm[4] = q->theta_s;
              prm[5] = q->min_Jp; prm[6] = q->max_Jp; L.off_scalar = (int32_t)offsetof(SnowParticle<3>, Jp); break; }
    case 3: { auto *q = static_cast<WaterParticle<3> *>(p0); prm[0] = q->k; prm[1] = q->gamma; L.off_scalar = (int32_t)offsetof(WaterParticle<3>, j); break; }
    case 4: { auto *q = static_cast<SandParticle<3> *>(p0); prm[0] = q->mu_0; prm[1] = q->lambda_0; prm[2] = q->alpha; prm[3] = q->cohesion; prm[4] = q->beta;
              L.off_scalar = (int32_t)offsetof(SandParticle<3>, logJp); break; }
    case 5: prm[0] = static_cast<ElasticParticle<3> *>(p0)->mu_0; prm[1] = static_cast<ElasticParticle<3> *>(p0)->lambda_0; break;
    case 6: { auto *q = static_cast<VonMisesParticle<3> *>(p0); prm[0] = q->mu_0; prm[1] = q->lambda_0; prm[2] = q->yield_stress; break; }
    case 7: { auto *q = static_cast<ViscoParticle<3> *>(p0); prm[0] = q->mu_0; prm[1] = q->lambda_0; prm[2] = q->visco_nu; prm[3] = q->visco_kappa; prm[4] = q->dt;
              L.off_scalar = (int32_t)offsetof(ViscoParticle<3>, visco_tau); break; }
    default: mpmb_release(e); return fail("unknown particle type", -103);
  }
  int rc = mpmb_set_material_(e, 0, kind, prm, 8);
  if (rc == MPMB_OK && m.levelset.levelset0 && !m.levelset.levelset0->planes.empty()) {
    std::vector<float> pl;
    for (auto &q : m.levelset.levelset0->planes) for (int k = 0; k < 4; k++) pl.push_back(q[k]);
    rc = mpmb_set_planes_(e, (int32_t)m.levelset.levelset0->planes.size(), pl.data(), m.levelset.levelset0->friction);
  }
  const int64_t n_up = (int64_t)m.particles.size();
  if (rc == MPMB_OK && coupled) {
    std::vector<float> off, tri;
    std::vector<int32_t> rid;
    for (auto ptr : rigid_ptrs) {
      auto *p = static_cast<RigidBoundaryParticle<3> *>(m.allocator[ptr]);
      for (int d = 0; d < 3; d++) off.push_back(p->offset[d]);
      for (int k = 0; k < 3; k++) for (int d = 0; d < 3; d++) tri.push_back(p->untransformed_element.v[k][d]);
      rid.push_back(p->rigid->id);
    }
    rc = mpmb_set_rigid_samples_(e, (int32_t)m.rigids.size(), (int64_t)rid.size(), off.data(), tri.data(), rid.data());
    if (rc == MPMB_OK) rc = mpmb_set_rigid_coupling_(e, m.penalty, m.pushing_force);
  }
  if (rc == MPMB_OK) rc = mpmb_upload_aos_(e, n_up, m.allocator.pool.data(), (int64_t)m.allocator.pool.size(), m.particles.data(), &L, nullptr);
  if (rc == MPMB_OK && coupled) {  // MPMParticle::states travel by id (= position in the upload)
    std::vector<uint32_t> st((size_t)n_up);
    for (int64_t k = 0; k < n_up; k++) st[k] = m.allocator[mpm_ptrs[k]]->states;
    rc = mpmb_set_particle_states_(e, n_up, st.data());
  }
  if (rc == MPMB_OK && !coupled) rc = mpmb_substep_(e, n);
  for (int s = 0; rc == MPMB_OK && coupled && s < n; s++) {   // the substep loop of INTEGRATION.md §2c: pose in, one substep, velocities out
    std::vector<MpmbRigidBody> rb(m.rigids.size());
    for (size_t b = 0; b < m.rigids.size(); b++) {
      const RigidBody<3> &r = *m.rigids[b];
      for (int d = 0; d < 3; d++) { rb[b].position[d] = r.position[d]; rb[b].velocity[d] = r.velocity[d]; rb[b].angular_velocity[d] = r.angular_velocity.value[d]; }
      store(r.rotation.value, rb[b].rot);
      store(r.inv_inertia, rb[b].inv_inertia);
      rb[b].inv_mass = r.inv_mass;
      rb[b].frictions[0] = r.frictions[0]; rb[b].frictions[1] = r.frictions[1];
    }
    rc = mpmb_set_rigid_state_(e, (int32_t)rb.size(), rb.data());
    if (rc == MPMB_OK) rc = mpmb_substep_(e, 1);
    if (rc == MPMB_OK) rc = mpmb_get_rigid_state_(e, (int32_t)rb.size(), rb.data());
    for (size_t b = 1; rc == MPMB_OK && b < m.rigids.size(); b++) {
      m.rigids[b]->velocity = VectorND<3, real>(rb[b].velocity[0], rb[b].velocity[1], rb[b].velocity[2]);
      m.rigids[b]->angular_velocity.value = VectorND<3, real>(rb[b].angular_velocity[0], rb[b].angular_velocity[1], rb[b].angular_velocity[2]);
    }
    m.advect_rigid_bodies(m.base_delta_t);   // host side, as in substep() (src/mpm.cpp:567-569); empty in this harness
  }
  int64_t alive = 0;
  if (rc == MPMB_OK) rc = mpmb_download_aos_(e, m.allocator.pool.data(), (int64_t)m.allocator.pool.size(), m.particles.data(), (int64_t)m.particles.size(), &L, &alive);
  if (rc == MPMB_OK && coupled) {
    std::vector<uint32_t> st((size_t)n_up);
    std::vector<float> nrm((size_t)n_up * 3), dist((size_t)n_up);
    std::vector<uint8_t> nearb((size_t)n_up);
    rc = mpmb_get_particle_cdf_(e, n_up, st.data(), nrm.data(), dist.data(), nearb.data());
    for (int64_t k = 0; rc == MPMB_OK && k < n_up; k++) {
      MPMParticle<3> *p = m.allocator[mpm_ptrs[k]];
      p->states = st[k];
      p->boundary_normal = VectorND<3, real>(nrm[3 * k], nrm[3 * k + 1], nrm[3 * k + 2]);
      p->boundary_distance = dist[k];
      p->near_boundary_ = nearb[k] != 0;
    }
  }
  if (rc != MPMB_OK) { fail(mpmb_last_error_(e), rc); mpmb_release(e); m.particles.insert(m.particles.end(), rigid_ptrs.begin(), rigid_ptrs.end()); return rc; }
  m.particles.resize((size_t)alive);                      // == what clear_boundary_particles leaves (src/mpm.cpp:583-633)
  m.particles.insert(m.particles.end(), rigid_ptrs.begin(), rigid_ptrs.end());
  m.current_t += n * m.base_delta_t;
  m.substep_counter += n;
  mpmb_release(e);
  return alive;
}

// offsetof() on the reference's own particle classes (src/particles.h, src/particles.cpp): the slot layout the drop-in
// adapter hands to mpmb_upload_aos.  out[8] = stride, pos, v_and_m, dg_e, apic_b, col_pitch, vol, scalar(kind) (-1: none)
void reft_aos_layout(int kind, int32_t *out) {
  using P3 = MPMParticle<3>;
  out[0] = (int32_t)sizeof(ParticleContainer<3>);
  out[1] = (int32_t)offsetof(P3, pos);
  out[2] = (int32_t)offsetof(P3, v_and_m);
  out[3] = (int32_t)offsetof(P3, dg_e);
  out[4] = (int32_t)offsetof(P3, apic_b);
  out[5] = (int32_t)sizeof(VectorND<3, real>);
  out[6] = (int32_t)offsetof(P3, vol);
  out[7] = kind == 2 ? (int32_t)offsetof(SnowParticle<3>, Jp) : kind == 3 ? (int32_t)offsetof(WaterParticle<3>, j)
           : kind == 4 ? (int32_t)offsetof(SandParticle<3>, logJp) : kind == 7 ? (int32_t)offsetof(ViscoParticle<3>, visco_tau) : -1;
}

// MPM<3>::step(dt) itself (src/mpm.cpp:428-450): `real` = float clocks decide how many substeps a frame runs
// (request_t += dt; while (current_t + base_delta_t < request_t) substep()).  Returns the substep counter.
int64_t reft_step(void *hp, float dt, float *current_t, float *request_t) {
  Solver &m = static_cast<Harness *>(hp)->m;
  m.step(dt);
  if (current_t) *current_t = m.current_t;
  if (request_t) *request_t = m.request_t;
  return (int64_t)m.substep_counter;
}

// frame dump by MPM<3>::write_partio itself (src/visualize.cpp:16-100) through the vendored Partio
void reft_write_partio(void *hp, const char *file_name) { static_cast<Harness *>(hp)->m.write_partio(file_name); }
int64_t reft_num_particles(void *hp) {   // MPM particles (RigidBoundaryParticles not counted)
  Solver &m = static_cast<Harness *>(hp)->m;
  int64_t n = 0;
  for (auto ptr : m.particles) n += m.allocator[ptr]->is_rigid() ? 0 : 1;
  return n;
}
// particle state by id (= order of reft_add_particle)
void reft_get_particles(void *hp, float *x, float *v, float *F, float *b, float *ps) {
  Harness *h = static_cast<Harness *>(hp);
  Solver &m = h->m;
  for (auto ptr : m.particles) {
    MPMParticle<3> *p = m.allocator[ptr];
    if (p->is_rigid()) continue;   // RigidBoundaryParticles (ids after the MPM particles') are not part of the read-back
    const int id = p->id;
    for (int d = 0; d < 3; d++) { x[3 * id + d] = p->pos[d]; v[3 * id + d] = p->get_velocity()[d]; }
    store(p->dg_e, F + 9 * id);
    store(p->apic_b, b + 9 * id);
    switch (h->kind[id]) {
      case 2: ps[id] = static_cast<SnowParticle<3> *>(p)->Jp; break;
      case 3: ps[id] = static_cast<WaterParticle<3> *>(p)->j; break;
      case 4: ps[id] = static_cast<SandParticle<3> *>(p)->logJp; break;
      case 7: ps[id] = static_cast<ViscoParticle<3> *>(p)->visco_tau; break;
      default: ps[id] = 0;
    }
  }
}
}

namespace taichi {
// the patch point of INTEGRATION.md §2, for a solver object whose substep() calls cannot be edited (AsyncMPM's scheduler):
// the reference's own MPM<3>::substep(), or one substep on libmpmb with the step the caller has just set in base_delta_t
void harness_substep(MPM<3> *m) {
  auto it = harness_registry().find(m);
  Harness *h = it == harness_registry().end() ? nullptr : it->second;
  if (!h || h->route_lib.empty()) { m->substep(); return; }
  char err[512] = {0};
  const int64_t rc = reft_substep_via_mpmb(h, h->route_lib.c_str(), 1, err, (int)sizeof(err));
  if (rc < 0) h->route_error = err;
  h->routed_substeps++;
}
}  // namespace taichi
