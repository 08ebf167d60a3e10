/*
 * mpmb.h — C-ABI of the B200-native MLS-MPM substep engine (libmpmb.so).
 *
 * This is the drop-in boundary for ONE hot path of yuanming-hu/taichi_mpm: the calls made by
 * MPM<3>::substep() (reference src/mpm.cpp:452-575) between particle ordering and particle
 * deletion.  Every entry point names the reference interface it replaces (path:line under the
 * reference tree).  Plain C types only: opaque handle, pointers and sizes; no exceptions cross the
 * boundary; every call returns 0 on success or a negative MpmbStatus, and the text of the last
 * failure is available from mpmb_last_error().
 *
 * Ownership: the library owns all device memory; the caller owns every host buffer it passes and
 * no host pointer is retained after a call returns.  Entry points are not re-entrant per handle;
 * distinct handles are independent.  One handle drives one GPU (one process per GPU; multi-GPU
 * z-slab runs exchange the buffers exposed by the mpmb_halo_* / mpmb_migrate_* calls with
 * whatever transport the host owns — torch.distributed/NCCL in this repository).
 *
 * Numeric contract: fp32 state (reference `real = float`, README.md:316-317), 3x3 matrices stored
 * column-major as 9 floats (m[c*3+r]; reference `M[i]` is column i, README.md:314), `b` is the
 * reference's `apic_b` (src/particles.h:24-45): C = -4/dx * b in the 88-line notation.
 */
#ifndef MPMB_H_
#define MPMB_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MPMB_VERSION 1
#define MPMB_MAX_GROUPS 16  /* material groups per handle */
#define MPMB_MAT_PARAMS 8

typedef struct MpmbEngine *MpmbHandle;

typedef enum MpmbStatus {
  MPMB_OK = 0,
  MPMB_ERR_INVALID = -1,   /* bad argument */
  MPMB_ERR_CUDA = -2,      /* a CUDA call failed (sticky per handle) */
  MPMB_ERR_CAPACITY = -3,  /* particle / tile / exchange-buffer capacity exceeded */
  MPMB_ERR_STATE = -4      /* call not valid in the current state */
} MpmbStatus;

/* Material kinds = the registered particle type names of the reference
 * (TC_REGISTER_MPM_PARTICLE, src/particles.cpp:845-856; all eight deformable types).  Parameter vectors (unused entries 0):
 *   LINEAR  "linear" (src/particles.cpp:297-363): [0]=mu [1]=lambda
 *   JELLY   "jelly"  (src/particles.cpp:365-438): [0]=mu [1]=lambda
 *   SNOW    "snow"   (src/particles.cpp:165-295): [0]=mu_0 [1]=lambda_0 [2]=hardening [3]=theta_c
 *                                                  [4]=theta_s [5]=min_Jp [6]=max_Jp ; scalar = Jp
 *   WATER   "water"  (src/particles.cpp:440-499): [0]=k [1]=gamma ; scalar = j
 *   SAND    "sand"   (src/particles.cpp:563-676): [0]=mu_0 [1]=lambda_0 [2]=alpha [3]=cohesion
 *                                                  [4]=beta ; scalar = logJp
 *   ELASTIC "elastic" (src/particles.cpp:764-841): [0]=mu_0 [1]=lambda_0   (Hencky, no return map)
 *   VON_MISES "von_mises" (src/particles.cpp:679-761): [0]=mu_0 [1]=lambda_0 [2]=yield_stress
 *   VISCO   "visco"  (src/particles.cpp:40-163): [0]=mu_0 [1]=lambda_0 [2]=visco_nu [3]=visco_kappa
 *                                                  [4]=dt (the particle's own `base_delta_t` key, default 1e-4,
 *                                                  src/particles.cpp:66) ; scalar = visco_tau              */
typedef enum MpmbMaterial {
  MPMB_MAT_LINEAR = 0,
  MPMB_MAT_JELLY = 1,
  MPMB_MAT_SNOW = 2,
  MPMB_MAT_WATER = 3,
  MPMB_MAT_SAND = 4,
  MPMB_MAT_ELASTIC = 5,
  MPMB_MAT_VON_MISES = 6,
  MPMB_MAT_VISCO = 7
} MpmbMaterial;

/* Mirrors the keys MPM<dim>::initialize reads from its Config (src/mpm.cpp:27-75). */
typedef struct MpmbConfig {
  int32_t res[3];            /* cells per axis (`res`); nodes = res+1 (src/mpm.cpp:66)            */
  float dx;                  /* `delta_x`                                                         */
  float dt;                  /* `base_delta_t` (already multiplied by dt_multiplier)              */
  float gravity[3];          /* `gravity` (src/mpm.cpp:38)                                        */
  int32_t particle_gravity;  /* src/mpm.cpp:47 (default 1): kick particles in P2G, not the grid   */
  int32_t clean_boundary;    /* src/mpm.cpp:563 (default 1): delete particles in the 7-cell band  */
  int32_t device;            /* CUDA device ordinal                                               */
  int64_t capacity;          /* max resident particles; 0 = sized at the first upload             */
  /* z-slab decomposition.  world==1: whole domain.  A rank owns the particles whose
   * base node lies in tile layers [tile_z0, tile_z1) (tiles are 4x4x4 nodes).                    */
  int32_t rank, world;
  int32_t tile_z0, tile_z1;
  int64_t migrate_capacity;  /* max particles leaving through one face per substep (world>1)      */
  int32_t halo_capacity;     /* max active tiles in one boundary layer; 0 = the whole cross-section */
  int32_t no_graph;          /* 1: launch every kernel from the host; 0 (default): mpmb_substep replays pairs of
                              * substeps as a CUDA graph (same kernels, same order, same results)              */
  int32_t reserved[6];       /* must be 0                                                         */
} MpmbConfig;

/* Byte offsets of the fields of one reference particle slot (ParticleContainer<3>, 320 B,
 * src/mpm_fwd.h:59-67; fields src/particles.h:24-45).  Obtained with offsetof() on the reference
 * side; never hard-coded here.  Offset < 0 = field absent.                       */
typedef struct MpmbAosLayout {
  int32_t stride;      /* bytes per slot (320)                                       */
  int32_t off_pos;     /* float[3]   `pos`                                           */
  int32_t off_v_and_m; /* float[4]   `v_and_m` = (vx,vy,vz,mass)                     */
  int32_t off_dg_e;    /* 3 x float[4] padded columns of `dg_e` (column pitch below) */
  int32_t off_apic_b;  /* 3 x float[4] padded columns of `apic_b`                    */
  int32_t col_pitch;   /* bytes between matrix columns (16)                          */
  int32_t off_vol;     /* float      `vol`                                           */
  int32_t off_scalar;  /* float      Jp / j / logJp of the subclass, or -1           */
  int32_t reserved[4];
} MpmbAosLayout;

/* ------------------------------------------------------------------------------ lifetime */
/* Replaces MPM<3>::initialize (src/mpm.cpp:27-75): sizes the tile grid, allocates device state. */
int mpmb_create(const MpmbConfig *cfg, MpmbHandle *out);
int mpmb_destroy(MpmbHandle h);
/* Text of the last error on this handle ("" if none).  h may be NULL for create-time errors.    */
const char *mpmb_last_error(MpmbHandle h);
int mpmb_version(void);
/* All work is enqueued on `cuda_stream` (a cudaStream_t cast to void*; NULL = default stream).   */
int mpmb_set_stream(MpmbHandle h, void *cuda_stream);
int mpmb_synchronize(MpmbHandle h);

/* ------------------------------------------------------------------------------ scene     */
/* Replaces Particle::initialize(config) of the registered type (e.g. SandParticle::initialize,
 * src/particles.cpp:587-597) for every particle of `group`.                                      */
int mpmb_set_material(MpmbHandle h, int32_t group, int32_t kind, const float *params, int32_t n_params);
/* (Called with particles of `group` already resident — between substeps — it also rebuilds their cached
 * affine matrices, so the next rasterize uses the new material's stress; apic_b must be current, i.e. the
 * call follows an upload or a completed mpmb_substep / mpmb_resample.)                              */
/* `base_delta_t` of the following substeps.  MPM<3> keeps it fixed, but AsyncMPM sets it before every MPM<dim>::substep() it
 * schedules (src/async/async_mpm.cpp:407-409: base_delta_t = unit_delta_t * delta_t_int), and a host that adapts its step does
 * the same.  With particles resident their cached affine matrices are rebuilt for the new step (apic_b must be current, as for
 * mpmb_set_material); an upload after the call needs nothing.  Between substeps only.                                      */
int mpmb_set_delta_t(MpmbHandle h, float dt);
/* Replaces Simulation::set_levelset + DynamicLevelSet::sample/get_spatial_gradient as used by
 * apply_grid_boundary_conditions (src/mpm.cpp:296-372) for a static level set.  `sdf4` is a host
 * array [res0+1][res1+1][res2+1][4] = (n_x,n_y,n_z,phi), phi in grid units at the node, n the
 * unit spatial gradient.  NULL removes the boundary.  `friction` = levelset0->friction
 * (-1 sticky, <=-2 slip, >=0 separate with Coulomb friction; src/mpm_fwd.h:25-57).               */
int mpmb_set_sdf(MpmbHandle h, const float *sdf4, float friction);
/* Same boundary built on the device from half-spaces phi_i(X) = n_i . X + d_i (grid units, n unit):
 * phi = min_i phi_i, n = n of the minimiser (levelset.add_plane in scripts/mls-cpic/sand_sweep.py:13-19). */
int mpmb_set_planes(MpmbHandle h, int32_t n_planes, const float *planes4, float friction);

/* The level set rasterised on the device from analytic solids (what the reference's scripts build with
 * levelset.add_plane / add_sphere / add_cuboid, e.g. scripts/mls-cpic/sand_sweep.py:13-19, scripts/async/sand.py:35,
 * scripts/mls-cpic/sand_stir.py:9): phi = min over the shapes of their signed distance in GRID units (negative inside
 * the obstacle), n = unit gradient of the minimiser; inside_out makes a container of a solid.  Parameters (grid units):
 *   PLANE  p = {n_x, n_y, n_z, d}       phi = n.X + d  (n unit)
 *   SPHERE p = {c_x, c_y, c_z, r}       phi = |X - c| - r
 *   CUBOID p = {lo_x, lo_y, lo_z, hi_x, hi_y, hi_z}   exact box distance
 * The shape formulas of the reference live in its un-vendored core (LevelSet3D): planes are pinned by the in-place build
 * (DESIGN.md §2), sphere and cuboid follow the published signed-distance definitions — parity unpinned for those two. */
typedef enum MpmbShapeKind { MPMB_SHAPE_PLANE = 0, MPMB_SHAPE_SPHERE = 1, MPMB_SHAPE_CUBOID = 2 } MpmbShapeKind;
typedef struct MpmbShape {
  int32_t kind;        /* MpmbShapeKind */
  int32_t inside_out;  /* 0: the solid is the obstacle; 1: its complement is (a container) */
  float p[6];
} MpmbShape;
int mpmb_set_levelset_shapes(MpmbHandle h, int32_t n_shapes, const MpmbShape *shapes, float friction);

/* ------------------------------------------------------------------------------ particles */
/* Replaces MPM<3>::add_particles' writes into the particle pool (src/mpm.cpp:93-148) for n
 * particles given field-wise: x[n][3], v[n][3], F[n][9], b[n][9] (apic_b), mass[n], vol[n],
 * scalar[n] (Jp/j/logJp), group[n].  F, b, scalar, group may be NULL (identity, 0, material
 * default, group 0).  Replaces the resident set.  Particle k gets id k.                           */
int mpmb_upload_particles(MpmbHandle h, int64_t n, const float *x, const float *v, const float *F, const float *b,
                          const float *mass, const float *vol, const float *scalar, const int32_t *group);
/* Particle k of the next upload gets id `base + k` (default 0); z-slab ranks use disjoint ranges so
 * that ids stay unique after migration.  Ids must stay below 2^28.                                 */
int mpmb_set_id_base(MpmbHandle h, int64_t base);
/* Same, reading the reference's own AoS pool: slot indices[k] of `pool` (ParticleAllocator::pool,
 * src/particle_allocator.h:39; MPM::particles index vector, src/mpm.h:116).  group[k] may be NULL. */
int mpmb_upload_aos(MpmbHandle h, int64_t n, const void *pool, int64_t pool_slots, const uint32_t *indices,
                    const MpmbAosLayout *layout, const int32_t *group);
/* Replaces the `benchmark` branch of MPM<3>::add_particles (src/mpm.cpp:149-186: 8 particles per cell of the
 * block [lo_cell, hi_cell) at the cell centre +- 0.25 dx, F = I, apic_b = 0, velocity v0) ON THE DEVICE — no host
 * array, which is what makes 10^7..10^8-particle scenes start in milliseconds.  `vol` / `mass` are the per-particle
 * values the caller derives as add_particles does (vol = dx^3 / maximum, mass = vol * density, src/mpm.cpp:134-135).
 * `jitter` (grid units, < 0.25; 0 = the reference's lattice) displaces every particle by a hash of its lattice index
 * and `seed`, so positions do not depend on the z-slab partition.  Particle ids are id_base + lattice index
 * (((ix*ny + iy)*nz + iz)*8 + corner, cells relative to lo_cell); a z-slab rank keeps the particles it owns;
 * particles within 7 cells of a face are not created (src/mpm.cpp:129-132).  Replaces the resident set.           */
int mpmb_seed_lattice(MpmbHandle h, const int32_t lo_cell[3], const int32_t hi_cell[3], float vol, float mass, float jitter,
                      uint32_t seed, int32_t group, const float v0[3], int64_t *n_seeded);
/* Number of live particles, counted on the device (4 bytes cross the bus; synchronises).
 * Reference: particles.size().                                                                    */
int mpmb_num_particles(MpmbHandle h, int64_t *n);
/* Sum over all substeps so far of the particles each one updated — the reference's `update_counter +=
 * particles.size()` per substep (src/mpm.cpp:436,449), counted on the device (exact when particles are
 * deleted in the middle of a frame).                                                               */
int mpmb_get_update_count(MpmbHandle h, int64_t *updates);
/* Copies the live particles out, any pointer may be NULL.  Row k of every array belongs to the
 * particle whose id (upload index) is id[k]; rows are in the engine's storage order.  `cap` rows
 * are available in each array; *n_out receives the number written.                                */
int mpmb_download_particles(MpmbHandle h, int64_t cap, int64_t *n_out, uint32_t *id, float *x, float *v, float *F,
                            float *b, float *mass, float *vol, float *scalar, int32_t *group);
/* Writes the live particles back into the reference's AoS pool (slot = indices[id]); returns the
 * number of survivors and compacts `indices` to the survivors (in id order), which is what
 * clear_boundary_particles (src/mpm.cpp:583-633) leaves in MPM::particles.  The scatter runs on the
 * device into the image of the pool kept since mpmb_upload_aos; both directions move only the window of
 * every slot that holds the layout's fields (a 2-D copy with the slot stride as pitch: 204 of the
 * reference's 320 bytes), so bytes outside it (vptr, flags, ...) are never touched on the host.  Fields
 * inside the window that the engine does not own (e.g. boundary_normal) return as uploaded: do not
 * modify the pool between the two calls (the device owns the particles during step()).
 * Without a preceding mpmb_upload_aos of the same pool the image is read first.                      */
int mpmb_download_aos(MpmbHandle h, void *pool, int64_t pool_slots, uint32_t *indices, int64_t n_indices,
                      const MpmbAosLayout *layout, int64_t *n_alive);

/* Frame dump without a host pass over the particles: the per-point records of MPM<3>::write_partio through Partio's
 * BGEO writer (src/visualize.cpp:16-100, external/partio/src/io/BGEO.cpp:131-150) — big-endian words position xyz,
 * w = 1, type = 0, index = id, limit = (1,1,1), v xyz: 48 bytes per particle, non-verbose dump — packed on the device in
 * id order and copied into `records` (cap_records x 48 bytes).  `id_range`: ids are below id_base + id_range.  The caller
 * writes the file header and trailer around the block (taichi_mpm_b200/bgeo.py does).                              */
int mpmb_download_bgeo_points(MpmbHandle h, int64_t id_range, void *records, int64_t cap_records, int64_t *n_out);

/* ------------------------------------------------------------------------------ hot path  */
/* The whole of MPM<3>::substep() on the fast path, `nsub` times (src/mpm.cpp:452-575):
 * sort_particles_and_populate_grid (464-465) -> rasterize_optimized (510-512) ->
 * normalize_grid_and_apply_external_force (526-533) -> apply_grid_boundary_conditions (539-540) ->
 * resample_optimized (548-549) -> clear_boundary_particles (563-565).  Asynchronous.              */
int mpmb_substep(MpmbHandle h, int32_t nsub);
/* The same stages one at a time (for parity tests and for a host that interleaves its own work,
 * e.g. rigid bodies).  Must be called in this order; mpmb_substep == these three.                  */
int mpmb_sort_particles_and_populate_grid(MpmbHandle h); /* src/mpm.cpp:770-918                   */
int mpmb_rasterize(MpmbHandle h);                        /* src/transfer.cpp:361-581 (P2G)        */
int mpmb_resample(MpmbHandle h);                         /* src/mpm.cpp:277-372 + src/transfer.cpp:702-970 + src/mpm.cpp:583-633 */

/* z-slab runs (world>1): the same two stages, each in two launches — part 1 = the tiles of the
 * slab's two boundary layers (they produce and consume halo data), part 2 = all other tiles — so
 * that the host can overlap the halo exchange with the interior work:
 *   sort; rasterize_part(1); halo_pack; [send/recv starts]; rasterize_part(2); resample_part(2);
 *   [send/recv done]; halo_unpack; resample_part(1); migrate_*.                                     */
int mpmb_rasterize_part(MpmbHandle h, int32_t part);
int mpmb_resample_part(MpmbHandle h, int32_t part);

/* Parity/debug: dense node grid [res0+1][res1+1][res2+1][4] on the host.
 * which=0: (p_x,p_y,p_z,m) after P2G; which=1: (v_x,v_y,v_z,m) after normalise + boundary.
 * Valid between mpmb_rasterize and the next mpmb_sort_particles_and_populate_grid.                 */
int mpmb_download_grid(MpmbHandle h, int32_t which, float *dense4);

/* ------------------------------------------------------------------------------ rigid bodies (CPIC) */
/* Two-way coupling with rigid bodies as the reference's optimized path does it.  On the device:
 * update_rigid_page_map (src/mpm.cpp:1026-1076), rasterize_rigid_boundary and gather_cdf (src/rigid_transfer.cpp:18-113,
 * 120-274) and the block_op_rigid branches of rasterize_optimized / resample_optimized (src/transfer.cpp:367-463, 706-835),
 * selected per 4x4x8-node page exactly as block_op_switch does.  On the host, as in the reference: the bodies themselves
 * (RigidBody: integration, scripted motion, rigid-rigid collisions, articulation — src/mpm_rigid_body.cpp:252-330 call into
 * the reference's un-vendored core for these).  Per substep the host hands over poses and velocities and reads the
 * velocities back:
 *     mpmb_set_rigid_state -> mpmb_substep(h, 1) -> mpmb_get_rigid_state -> host: advect_rigid_bodies
 * (mpmb_substep(h, n > 1) keeps the pose fixed over the n substeps.)  CUDA-graph replay is off while bodies are present.
 * Single-GPU engines only (world == 1).
 * What the device assumes of the core's RigidBody (it cannot be read here):
 *   get_velocity_at(p) = velocity + angular_velocity x (p - position);
 *   apply_tmp_impulse(j, p): tmp_velocity += inv_mass j, tmp_angular_velocity += inv_inertia ((p - position) x j);
 *   apply_tmp_velocity(): velocity += tmp_velocity, angular_velocity += tmp_angular_velocity (after each transfer);
 *   get_mesh_to_world() = get_centroid_to_world() = x -> position + rot x;
 *   world_to_element(e) = [v1 - v0, v2 - v0, n]^-1 with n the unit normal (v1 - v0) x (v2 - v0).                      */
typedef struct MpmbRigidBody {
  float position[3];          /* centre of mass, world units                                  */
  float rot[9];               /* mesh -> world linear part, column-major                      */
  float velocity[3];
  float angular_velocity[3];
  float inv_mass;             /* 0: scripted / infinite mass (set_infinity_mass)               */
  float inv_inertia[9];       /* world-space inverse inertia, column-major; 0: scripted rotation */
  float frictions[2];         /* RigidBody::frictions: [side of the particle's colour bit]     */
} MpmbRigidBody;
/* Replaces the RigidBoundaryParticles of MPM<3>::add_rigid_particle (src/mpm_rigid_body.cpp:137-250,
 * src/boundary_particle.h): sample s belongs to body rigid_id[s] (the index into MPM::rigids: 1..n_bodies-1; 0 is the
 * background body, src/mpm.cpp:72-74), sits at `offset` in the body's centroid frame and carries its triangle
 * `untransformed_element` (v0, v1, v2; 9 floats).  n_samples = 0 switches the coupling off.  n_bodies <= 12
 * (GridState::max_num_rigid_bodies, src/mpm_fwd.h:79).                                                                 */
int mpmb_set_rigid_samples(MpmbHandle h, int32_t n_bodies, int64_t n_samples, const float *offset3, const float *tri9,
                           const int32_t *rigid_id);
/* `penalty` (src/mpm.cpp:35, default 0) and `pushing_force` (src/mpm.cpp:40, default 20000).                          */
int mpmb_set_rigid_coupling(MpmbHandle h, float penalty, float pushing_force);
/* Poses and velocities of all n_bodies bodies (entry 0 ignored) for the next substep(s).                              */
int mpmb_set_rigid_state(MpmbHandle h, int32_t n_bodies, const MpmbRigidBody *bodies);
/* The same records with velocity / angular_velocity as the two transfers left them (synchronises).                    */
int mpmb_get_rigid_state(MpmbHandle h, int32_t n_bodies, MpmbRigidBody *bodies);
/* MPMParticle::states (src/particles.h) by particle id - id_base: set (default 0 for uploaded particles) / read together
 * with boundary_normal, boundary_distance and near_boundary_ of the last substep.  Any pointer may be NULL.            */
int mpmb_set_particle_states(MpmbHandle h, int64_t n, const uint32_t *states);
int mpmb_get_particle_cdf(MpmbHandle h, int64_t n, uint32_t *states, float *normal3, float *distance, uint8_t *near_boundary);
/* Parity/debug: the node colour field of the last substep, dense [res0+1][res1+1][res2+1]: GridState::states
 * (tags | (rigid id + 1) << 24) and GridState::distance (world units).                                                 */
int mpmb_download_cdf(MpmbHandle h, uint32_t *node_states, float *node_distance);

/* ------------------------------------------------------------------------------ profiling */
#define MPMB_N_STAGES 5 /* 0 sort+tiles, 1 P2G, 2 G2P, 3 exchange pack/unpack, 4 grid update */
/* When enabled, CUDA events bracket every stage on the engine's stream.                          */
int mpmb_set_profiling(MpmbHandle h, int32_t enabled);
/* Accumulated milliseconds and launch counts per stage since the last reset (synchronises).      */
int mpmb_get_profile(MpmbHandle h, double ms[MPMB_N_STAGES], int64_t launches[MPMB_N_STAGES], int32_t reset);
/* Device counters: active tiles of the last substep, live particles, kernels launched so far.    */
int mpmb_get_counters(MpmbHandle h, int64_t *active_tiles, int64_t *alive, int64_t *kernel_launches);

/* State of the incremental ordering: rows of the current storage, particles that changed tile in the last substep
 * (they are holes of their old runs until the next ordering re-bins them), ghost tiles received from the neighbours. */
int mpmb_get_ordering_stats(MpmbHandle h, int64_t *rows, int64_t *movers, int64_t *ghost_tiles);

/* ------------------------------------------------------------------------------ multi-GPU */
/* z-slab runs (world>1).  The host moves the bytes; the engine packs and unpacks on the device.
 * Halo: after mpmb_rasterize, face 0 (-z) / 1 (+z) tile-layer partial sums of (p,m) are packed
 * into a device buffer of mpmb_halo_bytes() bytes; the neighbour unpacks them as ghost tiles
 * before mpmb_resample.  Migration: after mpmb_resample, particles that left the slab through a
 * face are packed (mpmb_migrate_bytes() bytes per face) and appended by the neighbour before its
 * next sort.  All pointers are DEVICE pointers owned by the caller.                               */
int64_t mpmb_halo_bytes(MpmbHandle h);
int mpmb_halo_pack(MpmbHandle h, int32_t face, void *dev_buf);
int mpmb_halo_unpack(MpmbHandle h, int32_t face, const void *dev_buf);
int64_t mpmb_migrate_bytes(MpmbHandle h);
int mpmb_migrate_pack(MpmbHandle h, int32_t face, void *dev_buf);
int mpmb_migrate_unpack(MpmbHandle h, int32_t face, const void *dev_buf);

/* Peer-memory exchange: with the neighbours' receive buffers mapped (CUDA IPC between processes, or
 * plain pointers inside one process) the engine needs no host transport at all — the pack kernels
 * store straight into the neighbour GPU over NVLink, a {count, seq} header is released system-wide
 * and the neighbour's unpack waits on it.  mpmb_substep(h, n) then runs whole z-slab substeps,
 * exchanges included, without returning to the host.
 *   kind 0 = halo arenas, kind 1 = migrating particles.  Through face f a rank writes into the
 *   neighbour's buffer of the OPPOSITE face.                                                        */
int mpmb_xchg_buffer(MpmbHandle h, int32_t kind, int32_t face, void **dev_ptr);      /* my receive buffer   */
int mpmb_xchg_ipc_handle(MpmbHandle h, int32_t kind, int32_t face, void *out64);     /* its 64-byte IPC handle */
int mpmb_xchg_connect(MpmbHandle h, int32_t kind, int32_t face, const void *handle64, void *same_process_ptr);
/* the four exchange steps of one substep, for hosts that drive the stages themselves */
int mpmb_halo_send(MpmbHandle h, int32_t face);     /* after mpmb_rasterize                      */
int mpmb_halo_recv(MpmbHandle h, int32_t face);     /* before mpmb_resample                      */
int mpmb_migrate_send(MpmbHandle h, int32_t face);  /* after mpmb_resample                       */
int mpmb_migrate_recv(MpmbHandle h, int32_t face);  /* before the next ordering                  */

#ifdef __cplusplus
}
#endif
#endif /* MPMB_H_ */
