// libmpmb.so — B200-native MLS-MPM substep engine (sm_100a).  C-ABI in include/mpmb.h.
//
// Data layout in HBM:
//   * particles: SoA of ten float4 streams (160 B/particle), double buffered.
//       P2G set  q0=(x,y,z,mass) q1=(vx,vy,vz,A0) q2=(A1..A4) q3=(A5..A8)
//       G2P set  q4=(F0..F3) q5=(F4..F7) q6=(F8,scalar,vol,tag)     tag = group<<28 | id
//       state    q7=(b0..b3) q8=(b4..b7) q9=(b8,-,-,-)             b = apic_b
//     A = calculate_force()*(-4 dt/dx) + apic_b*(4 m) is the affine matrix rasterize needs
//     (src/transfer.cpp:503,521-522).  G2P produces it for the NEXT substep from the same
//     constitutive evaluation as the return map, so P2G reads 64 B/particle and does no
//     constitutive math, and each particle is factorised at most once per substep.
//   * order: every active tile (4x4x4 nodes) owns a contiguous RUN of storage rows plus a short
//     ARRIVAL list.  G2P writes each tile's particles (cell-sorted) into one contiguous run of the
//     other buffer, so storage is re-compacted every substep for free; a particle whose base node
//     left its tile (~0.5 % per substep) is appended to a mover list and re-binned by four tiny
//     kernels (count, scan, place, rank).  There is NO per-substep radix sort and no permutation
//     array: run rows are read with unit stride.  The radix sort runs once, at upload.
//     This replaces sort_particles_and_populate_grid (src/mpm.cpp:770-918).
//   * grid: no dense grid.  P2G leaves one 6x6x6 float4 "arena" (tile + the +2 stencil halo,
//     the reference's GridCache footprint src/transfer.cpp:59-63) per active tile; k_grid rebuilds
//     each node as the fixed-order sum of the <=8 arenas covering it, normalises and applies the
//     level-set boundary.  No global float atomics, no float atomics at all: single-GPU results
//     are bit-reproducible run to run.
#include <cuda_runtime.h>
#include <cub/block/block_reduce.cuh>
#include <cub/block/block_scan.cuh>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/mpmb.h"
#include "mpmb_math.cuh"

namespace mpmb {

constexpr int ARENA = 216;  // 6*6*6 nodes
constexpr int N_Q = 10;
// tag = group << TAG_ID_BITS | id: MPMB_MAX_GROUPS = 16 groups, ids below 2^28 = 268 M (BASELINE config 5 holds 64 M)
constexpr int TAG_ID_BITS = 28;
constexpr uint32_t TAG_ID_MASK = (1u << TAG_ID_BITS) - 1u;
static_assert(MPMB_MAX_GROUPS <= (1 << (32 - TAG_ID_BITS)), "group bits");
// keys: [0, ntiles_total) = tile index; ntiles_total + {0,1,2} = left through -z / +z face, dead
enum { SPECIAL_MIG_DOWN = 0, SPECIAL_MIG_UP = 1, SPECIAL_DEAD = 2 };

struct Params {
  int res[3];
  int nnode[3];
  // DENSE tile grid of this engine.  world == 1: the whole domain.  z-slab rank: its own tile layers plus one ghost
  // layer on each side, i.e. local layer l holds global layer l + tz_off — the ordering scans, the slot map and the
  // keys are slab-local, so their cost does not grow with the number of ranks.
  int nt[3];
  int tz_off;
  int ntiles_total;
  float dx, inv_dx, dt;
  float gdt[3];
  int particle_gravity, clean_boundary;
  float friction;
  int has_sdf;
  int world, tile_z0, tile_z1;
  Material mats[MPMB_MAX_GROUPS];
};

struct Counters {  // device-resident
  int n_store;        // rows of the current storage that hold a particle record
  int n_tiles;        // active tiles of this substep
  int n_ghost;        // ghost tiles appended after the owned ones (world>1)
  int error;          // sticky device-side error flags
  int n_movers;       // mover list consumed by the next ordering
  int n_movers_next;  // mover list being filled by G2P / migrate_unpack
  int n_alive;        // particles binned by the last ordering
  int n_live;         // scratch of k_count_live (mpmb_num_particles)
  unsigned long long updates;  // sum over substeps of the particles binned: the reference's update_counter (src/mpm.cpp:436)
  int g2p_done;       // CTAs of the running k_g2p that have finished (the last one commits the substep)
  int xstep;          // substeps completed since the peers were connected: the sequence number of the peer exchange
                      // (kept on the device so that a captured CUDA graph of a substep stays valid from replay to replay)
  int chk[8];         // MPMB_CHECKED builds: first violated index check {site, a, b, c, ...}
};
// Bounds-checked debug build (-DMPMB_CHECKED): an index that would leave its array is recorded (first one wins) and the
// access skipped, so a corrupted ordering shows up as a report (mpmb_debug_check) instead of an illegal-address fault.
#ifdef MPMB_CHECKED
#define MPMB_CHK(cnt, cond, site, a, b, c)                                                                        \
  ((cond) ? true : (atomicCAS(&(cnt)->chk[0], 0, (site)) == 0 ? ((cnt)->chk[1] = (int)(a), (cnt)->chk[2] = (int)(b), (cnt)->chk[3] = (int)(c), false) : false))
#else
#define MPMB_CHK(cnt, cond, site, a, b, c) true
#endif
enum { DEVERR_TILE_CAPACITY = 1, DEVERR_MIGRATE_CAPACITY = 2, DEVERR_PARTICLE_CAPACITY = 4, DEVERR_PEER_TIMEOUT = 8, DEVERR_BAD_INPUT = 16 };

struct TileMeta {  // one 32-byte record per active tile (slot)
  int tile, run_begin, run_len, arr_off, arr_len, out_begin, xy, z;  // xy = tile x | y << 16, z = GLOBAL tile layer
};
// TileMeta::pad0/pad1 carry the tile's (x | y<<16, z) coordinates, decoded once per tile by k_order_c, so the tile
// kernels skip two runtime integer divisions per CTA per tile/chunk.
// Measured A/B of round 2 (profiles/README.md): kept TILE_XYZ (-0.017 ms/substep) and the per-warp P2G arenas
// (-0.018 ms); dropped the 192-thread x-plane P2G (0.317 vs 0.283 ms: +16 % instructions and 3x row reads cost
// more than 18 warps/SM hide) and the per-tile level-set flags for k_grid (no change: the band test is not its limiter).

struct View {  // raw pointers handed to kernels
  float4 *q[N_Q];        // current (read) buffer
  float4 *qn[N_Q];       // next (written by G2P)
  uint32_t *keys;        // tile of every current row (specials: dead / migrating)
  uint32_t *keys_next;
  uint32_t *outpos;      // per live current row: its row in the next storage (written by P2G)
  // dense per-tile ordering state
  int *run_begin, *run_len;      // run of the tile in the current storage
  int *out_begin, *total;        // run of the tile in the next storage (= next substep's run)
  int *stay_cnt, *stay_next;     // rows of the run still owned by the tile (now / after this G2P)
  int *arr_cnt, *arr_off, *arr_len, *arr_cur;
  int *slot_map;
  TileMeta *meta;                // [slot]
  uint32_t *mover_dst, *mover_idx;      // consumed by the ordering
  uint32_t *mover_dst_n, *mover_idx_n;  // produced by G2P
  uint32_t *arrivals, *arrivals_sorted;
  int *blocksum;
  float4 *arena;
  const float4 *sdf4;
  Counters *cnt;
  int cap_tiles;
  int cap_particles;
  const unsigned char *rigid_flag;  // [slot] 1 = the tile lies in a rigid page: k_g2p leaves it to k_g2p_rigid (null: no rigid bodies)
};

// ------------------------------------------------------------------------------ helpers
// Asynchronous 16-byte global->shared copies (LDGSTS): the gathered particle rows of a tile are
// put in flight in one batch, no register staging, and waited for once.
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src) {
  unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

// TMA bulk copies (cp.async.bulk, SASS UBLKCP) completing on an mbarrier: ONE elected thread moves a whole unit-stride
// run chunk (4 KB per stream) or a tile's 3.4 KB node block global -> shared; nobody computes addresses per row, and the
// consumers wait on the barrier's phase parity instead of on their own copy groups.
#ifndef MPMB_SIMT_HOST
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// the elected thread's arrival + the number of bytes the copies of this phase will deliver
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gmem_src, unsigned bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
#else
// SIMT emulator (tests/simt): a phase completes at the elected thread's arrival (the word counts completed phases and a
// waiter yields to the other coroutines until the phase of its parity is over), but the bulk copies LAND LATE — they are
// queued and performed by the first thread that waits on their barrier — so a consumer that reads without waiting, or
// another copy that races with a bulk copy's bytes, shows up in the emulated parity runs.
struct SimtBulk { void *dst; const void *src; unsigned bytes; uint64_t *bar; };
static std::vector<SimtBulk> g_simt_bulk;
__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned) { *bar = 0; }
__device__ __forceinline__ void mbar_init_fence() {}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, unsigned) { *bar += 1; }
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gmem_src, unsigned bytes, uint64_t *bar) {
  g_simt_bulk.push_back(SimtBulk{smem_dst, gmem_src, bytes, bar});
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity) {
  while ((unsigned)(*(volatile uint64_t *)bar & 1u) == parity) simt::yield();
  for (size_t i = 0; i < g_simt_bulk.size();) {
    if (g_simt_bulk[i].bar == bar) {
      memcpy(g_simt_bulk[i].dst, g_simt_bulk[i].src, g_simt_bulk[i].bytes);
      g_simt_bulk.erase(g_simt_bulk.begin() + i);
    } else {
      i++;
    }
  }
}
#endif

// A float4 quantity (x, y | z, w) and its weights as pairs.  The pair form was written for Blackwell's packed fp32
// (FFMA2 / FMUL2 / FADD2, `fma.rn.f32x2`); measured on a B200 (profiles/README.md, round 2) the packed instructions were
// SLOWER than the scalar ones in both tile kernels (k_p2g 0.253 vs 0.241 ms, k_g2p 0.354 vs 0.346 ms: FFMA2 occupies
// the fma pipe for two issue slots and the aligned register pairs cost k_g2p 28 registers), so the operations below
// are the scalar fmaf / mul / sub — the node loops keep the pair structure, which the scalar code generator likes
// (-0.017 ms in k_p2g against the float[27][4] accumulators of round 1).
struct F4 {
  float2 lo, hi;  // (x, y), (z, w)
};
__device__ __forceinline__ float2 bc2(float w) { return make_float2(w, w); }
__device__ __forceinline__ F4 f4_load(const float4 &v) { return F4{make_float2(v.x, v.y), make_float2(v.z, v.w)}; }
__device__ __forceinline__ F4 f4_zero() { return F4{make_float2(0.f, 0.f), make_float2(0.f, 0.f)}; }
__device__ __forceinline__ float2 mul2(float2 a, float2 b) { return make_float2(a.x * b.x, a.y * b.y); }
__device__ __forceinline__ F4 f4_fma(float2 w, const F4 &a, const F4 &c) {
  return F4{make_float2(fmaf(w.x, a.lo.x, c.lo.x), fmaf(w.y, a.lo.y, c.lo.y)), make_float2(fmaf(w.x, a.hi.x, c.hi.x), fmaf(w.y, a.hi.y, c.hi.y))};
}
__device__ __forceinline__ F4 f4_mul(float2 w, const F4 &a) { return F4{make_float2(w.x * a.lo.x, w.y * a.lo.y), make_float2(w.x * a.hi.x, w.y * a.hi.y)}; }
__device__ __forceinline__ F4 f4_sub(const F4 &a, const F4 &b) {
  return F4{make_float2(a.lo.x - b.lo.x, a.lo.y - b.lo.y), make_float2(a.hi.x - b.hi.x, a.hi.y - b.hi.y)};
}

__device__ __forceinline__ void base_rel(float x, float inv_dx, int &base, float &rel) {
  // pos_ = p.pos * inv_delta_x (src/transfer.cpp:490); base = int(x - 0.5f) (src/kernel.h:119-121).
  // Explicit round-to-nearest ops so that no FMA contraction changes the cell assignment.
  float X = __fmul_rn(x, inv_dx);
  base = (int)__fsub_rn(X, 0.5f);
  rel = __fsub_rn(X, (float)base);
}

__device__ __forceinline__ uint32_t make_key(const Params &P, float x, float y, float z) {
  int bx, by, bz;
  float r;
  base_rel(x, P.inv_dx, bx, r);
  base_rel(y, P.inv_dx, by, r);
  base_rel(z, P.inv_dx, bz, r);
  // the 27-node stencil [base, base+2] must stay on the node grid
  bool in_domain = (bx >= 0) && (by >= 0) && (bz >= 0) && (bx + 2 < P.nnode[0]) && (by + 2 < P.nnode[1]) && (bz + 2 < P.nnode[2]);
  if (!in_domain) return (uint32_t)(P.ntiles_total + SPECIAL_DEAD);
  int tx = bx >> 2, ty = by >> 2, tz = bz >> 2;
  if (P.world > 1) {
    if (tz < P.tile_z0) return (uint32_t)(P.ntiles_total + SPECIAL_MIG_DOWN);
    if (tz >= P.tile_z1) return (uint32_t)(P.ntiles_total + SPECIAL_MIG_UP);
  }
  return (uint32_t)((tx * P.nt[1] + ty) * P.nt[2] + (tz - P.tz_off));
}

// near_boundary + abnormal (src/mpm.h:269-276, src/mpm.cpp:595-598)
__device__ __forceinline__ bool reference_deletes(const Params &P, float3 x, float3 v) {
  float X = x.x * P.inv_dx, Y = x.y * P.inv_dx, Z = x.z * P.inv_dx;
  float mn = fminf(X, fminf(Y, Z));
  float mx = fmaxf(X - (float)P.res[0], fmaxf(Y - (float)P.res[1], Z - (float)P.res[2]));
  bool bad = (mn < 7.0f) || (mx > -7.0f);
  bad |= !isfinite((x.x + x.y + x.z) + (v.x + v.y + v.z));  // any inf/NaN poisons the sum (pos.abnormal() || v.abnormal())
  return bad;
}

// z-slab runs split every tile kernel in two launches so that the halo exchange overlaps compute:
// part 1 = tiles of the slab's boundary layers (they produce / consume halo data), part 2 = the rest,
// part 0 = all tiles.
__device__ __forceinline__ bool tile_in_part(const Params &P, int tile, int part) {
  if (part == 0) return true;
  const int tz = tile % P.nt[2] + P.tz_off;
  const bool boundary = (P.world > 1) && (tz == P.tile_z0 || tz == P.tile_z1 - 1);
  return part == 1 ? boundary : !boundary;
}

// affine = stress * (-4 inv_dx dt) + apic_b * (inv_D * mass)  (src/transfer.cpp:465,503,521-522)
__device__ __forceinline__ void make_affine(const Mat3 &force, const Mat3 &b, float mass, float S, Mat3 &A) {
  const float bm = 4.0f * mass;
#pragma unroll
  for (int k = 0; k < 9; k++) A.m[k] = fmaf(force.m[k], S, b.m[k] * bm);
}

template <bool STORE_B = true>
__device__ __forceinline__ void store_particle(float4 *const *q, size_t i, float3 x, float mass, float3 v, const Mat3 &A, const Mat3 &F,
                                               float ps, float vol, uint32_t tag, const Mat3 &b) {
  q[0][i] = make_float4(x.x, x.y, x.z, mass);
  q[1][i] = make_float4(v.x, v.y, v.z, A.m[0]);
  q[2][i] = make_float4(A.m[1], A.m[2], A.m[3], A.m[4]);
  q[3][i] = make_float4(A.m[5], A.m[6], A.m[7], A.m[8]);
  q[4][i] = make_float4(F.m[0], F.m[1], F.m[2], F.m[3]);
  q[5][i] = make_float4(F.m[4], F.m[5], F.m[6], F.m[7]);
  q[6][i] = make_float4(F.m[8], ps, vol, __uint_as_float(tag));
  if (STORE_B) {
    q[7][i] = make_float4(b.m[0], b.m[1], b.m[2], b.m[3]);
    q[8][i] = make_float4(b.m[4], b.m[5], b.m[6], b.m[7]);
    q[9][i] = make_float4(b.m[8], 0.f, 0.f, 0.f);
  }
}

// ------------------------------------------------------------------------------ upload kernels
__global__ void k_iota(uint32_t *a, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) a[i] = (uint32_t)i;
}

__device__ __forceinline__ void pack_one(View &V, const Params &P, int i, uint32_t id_base, float3 x, float3 v, Mat3 F, Mat3 b, float mass, float vol, float ps,
                                         int g, uint32_t *keys) {
  // a hole is encoded by the sign of the mass and the group by 4 tag bits: reject what cannot be represented
  if (g < 0 || g >= MPMB_MAX_GROUPS || !(mass > 0.f)) {
    atomicOr(&V.cnt->error, DEVERR_BAD_INPUT);
    g = min(max(g, 0), MPMB_MAX_GROUPS - 1);
    mass = 1e-30f;
  }
  Mat3 force, A;
  calculate_force(P.mats[g], F, ps, vol, force);
  make_affine(force, b, mass, -4.0f * P.inv_dx * P.dt, A);
  store_particle(V.q, (size_t)i, x, mass, v, A, F, ps, vol, ((uint32_t)g << TAG_ID_BITS) | (id_base + (uint32_t)i), b);
  keys[i] = make_key(P, x.x, x.y, x.z);
}

// Field-wise host arrays (staged on the device) -> q streams.
__global__ void k_pack_particles(View V, Params P, int n, const float *x, const float *v, const float *F, const float *b,
                                 const float *mass, const float *vol, const float *scalar, const int *group, uint32_t *keys, uint32_t id_base) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Mat3 f, bb;
#pragma unroll
  for (int k = 0; k < 9; k++) {
    f.m[k] = F ? F[9 * (size_t)i + k] : ((k % 4 == 0) ? 1.f : 0.f);
    bb.m[k] = b ? b[9 * (size_t)i + k] : 0.f;
  }
  int g = group ? group[i] : 0;
  float ps;
  if (scalar) ps = scalar[i];
  else {
    int kind = P.mats[g].kind;
    ps = (kind == MAT_SNOW || kind == MAT_WATER) ? 1.0f : 0.0f;  // Jp=1 (205), j=1 (461), logJp=0 (595)
  }
  float3 xx = make_float3(x[3 * (size_t)i], x[3 * (size_t)i + 1], x[3 * (size_t)i + 2]);
  float3 vv = make_float3(v[3 * (size_t)i], v[3 * (size_t)i + 1], v[3 * (size_t)i + 2]);
  pack_one(V, P, i, id_base, xx, vv, f, bb, mass[i], vol[i], ps, g, keys);
}

// Reference AoS slots (staged on the device) -> q streams.
__global__ void k_pack_aos(View V, Params P, int n, const unsigned char *pool, const uint32_t *indices, MpmbAosLayout L,
                           const int *group, uint32_t *keys, uint32_t id_base) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned char *s = pool + (size_t)indices[i] * L.stride;
  const float *pos = (const float *)(s + L.off_pos);
  const float *vm = (const float *)(s + L.off_v_and_m);
  Mat3 f, bb;
  for (int c = 0; c < 3; c++) {
    const float *fc = (const float *)(s + L.off_dg_e + c * L.col_pitch);
    const float *bc = (const float *)(s + L.off_apic_b + c * L.col_pitch);
    for (int r = 0; r < 3; r++) {
      f.m[c * 3 + r] = fc[r];
      bb.m[c * 3 + r] = bc[r];
    }
  }
  int g = group ? group[i] : 0;
  float vol = *(const float *)(s + L.off_vol);
  float ps = L.off_scalar >= 0 ? *(const float *)(s + L.off_scalar) : 0.f;
  pack_one(V, P, i, id_base, make_float3(pos[0], pos[1], pos[2]), make_float3(vm[0], vm[1], vm[2]), f, bb, vm[3], vol, ps, g, keys);
}

// q streams -> field-wise arrays, compacting live particles (storage order).
__global__ void k_unpack_particles(View V, const uint32_t *keys, int n, uint32_t special_min, uint32_t *id, float *x, float *v, float *F,
                                   float *b, float *mass, float *vol, float *scalar, int *group, const int *prefix) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (keys[i] >= special_min) return;
  int o = prefix[i];
  float4 q0 = V.q[0][i], q1 = V.q[1][i], q4 = V.q[4][i], q5 = V.q[5][i], q6 = V.q[6][i], q7 = V.q[7][i], q8 = V.q[8][i], q9 = V.q[9][i];
  uint32_t tag = __float_as_uint(q6.w);
  if (id) id[o] = tag & TAG_ID_MASK;
  if (group) group[o] = (int)(tag >> TAG_ID_BITS);
  if (x) { x[3 * (size_t)o] = q0.x; x[3 * (size_t)o + 1] = q0.y; x[3 * (size_t)o + 2] = q0.z; }
  if (mass) mass[o] = fabsf(q0.w);
  if (v) { v[3 * (size_t)o] = q1.x; v[3 * (size_t)o + 1] = q1.y; v[3 * (size_t)o + 2] = q1.z; }
  if (F) {
    float *f = F + 9 * (size_t)o;
    f[0] = q4.x; f[1] = q4.y; f[2] = q4.z; f[3] = q4.w; f[4] = q5.x; f[5] = q5.y; f[6] = q5.z; f[7] = q5.w; f[8] = q6.x;
  }
  if (scalar) scalar[o] = q6.y;
  if (vol) vol[o] = q6.z;
  if (b) {
    float *p = b + 9 * (size_t)o;
    p[0] = q7.x; p[1] = q7.y; p[2] = q7.z; p[3] = q7.w; p[4] = q8.x; p[5] = q8.y; p[6] = q8.z; p[7] = q8.w; p[8] = q9.x;
  }
}

// live rows of the current storage (keys below the specials), reduced on the device: mpmb_num_particles reads 4 bytes
__global__ void __launch_bounds__(256) k_count_live(const uint32_t *keys, uint32_t special_min, Counters *cnt) {
  typedef cub::BlockReduce<int, 256> Red;
  __shared__ typename Red::TempStorage tmp;
  const int n = cnt->n_store;
  int c = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) c += keys[i] < special_min;
  c = Red(tmp).Sum(c);
  if (threadIdx.x == 0 && c) atomicAdd(&cnt->n_live, c);
}

__global__ void k_alive_flags(const uint32_t *keys, int n, uint32_t special_min, int *flags) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) flags[i] = keys[i] < special_min ? 1 : 0;
}

__global__ void k_fill_u32(uint32_t *a, int n, uint32_t v) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) a[i] = v;
}

// ------------------------------------------------------------------------------ device-side seeding
// Replaces the `benchmark` lattice of MPM<3>::add_particles (src/mpm.cpp:149-186: 8 particles per cell at the cell
// centre +- 0.25 dx) without a host array: optional jitter from a counter-based hash of the particle's LATTICE index, so
// the same particle gets the same position whatever the z-slab partition (and from the numpy twin
// scenes.lattice_block_hashed, the parity reference).  id = lattice index; a z-slab rank keeps the particles whose base
// node lies in its tile layers; particles in the 7-cell boundary band are not created (src/mpm.cpp:129-132).
struct SeedBox {
  int lo[3], n[3];      // first cell and number of cells of the whole block
  int z_first, z_count; // candidate cell layers of this rank (relative to lo[2])
  float jitter, vol, mass, v0[3];
  float ps;             // plastic scalar of the material's fresh state
  uint32_t seed, id_base;
  int group;
};
__host__ __device__ __forceinline__ uint32_t seed_hash(uint32_t idx, uint32_t axis, uint32_t seed) {
  uint32_t h = (idx * 3u + axis) ^ seed;
  h *= 0x9E3779B1u; h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;
  return h;
}
// candidate c -> lattice index, position; returns whether this rank creates the particle
__device__ __forceinline__ bool seed_candidate(const Params &P, const SeedBox &B, size_t c, uint32_t &lattice, float3 &x, uint32_t &key) {
  const int corner = (int)(c & 7u);
  size_t cell = c >> 3;
  const int kz = (int)(cell % (size_t)B.z_count) + B.z_first;
  cell /= (size_t)B.z_count;
  const int ky = (int)(cell % (size_t)B.n[1]), kx = (int)(cell / (size_t)B.n[1]);
  lattice = (uint32_t)((((size_t)kx * B.n[1] + ky) * B.n[2] + kz) * 8u + corner);
  const int ic[3] = {B.lo[0] + kx, B.lo[1] + ky, B.lo[2] + kz};
  float X[3];
#pragma unroll
  for (int a = 0; a < 3; a++) {
    const float off = ((corner >> a) & 1) ? 0.75f : 0.25f;
    const float u = __fsub_rn((float)(seed_hash(lattice, (uint32_t)a, B.seed) >> 8) * (1.0f / 8388608.0f), 1.0f);  // [-1, 1)
    X[a] = __fadd_rn(__fadd_rn((float)ic[a], off), __fmul_rn(B.jitter, u));                                        // grid units
  }
  x = make_float3(__fmul_rn(X[0], P.dx), __fmul_rn(X[1], P.dx), __fmul_rn(X[2], P.dx));
  const float mn = fminf(X[0], fminf(X[1], X[2]));
  const float mx = fmaxf(X[0] - (float)P.res[0], fmaxf(X[1] - (float)P.res[1], X[2] - (float)P.res[2]));
  if (mn < 7.0f || mx > -7.0f) return false;  // near_boundary (src/mpm.h:269-276)
  key = make_key(P, x.x, x.y, x.z);
  return key < (uint32_t)P.ntiles_total;      // owned by this rank (not a migration / dead key)
}
__global__ void k_seed_flags(Params P, SeedBox B, size_t n_cand, int *flags) {
  for (size_t c = blockIdx.x * (size_t)blockDim.x + threadIdx.x; c < n_cand; c += (size_t)gridDim.x * blockDim.x) {
    uint32_t lattice, key;
    float3 x;
    flags[c] = seed_candidate(P, B, c, lattice, x, key) ? 1 : 0;
  }
}
__global__ void k_seed_write(View V, Params P, SeedBox B, size_t n_cand, const int *flags, const int *prefix, uint32_t *keys) {
  for (size_t c = blockIdx.x * (size_t)blockDim.x + threadIdx.x; c < n_cand; c += (size_t)gridDim.x * blockDim.x) {
    if (!flags[c]) continue;
    uint32_t lattice, key;
    float3 x;
    seed_candidate(P, B, c, lattice, x, key);
    Mat3 F, b, force, A;
#pragma unroll
    for (int k = 0; k < 9; k++) { F.m[k] = (k % 4 == 0) ? 1.f : 0.f; b.m[k] = 0.f; }
    calculate_force(P.mats[B.group], F, B.ps, B.vol, force);
    make_affine(force, b, B.mass, -4.0f * P.inv_dx * P.dt, A);
    const size_t o = (size_t)prefix[c];
    store_particle(V.q, o, x, B.mass, make_float3(B.v0[0], B.v0[1], B.v0[2]), A, F, B.ps, B.vol, ((uint32_t)B.group << TAG_ID_BITS) | (B.id_base + lattice), b);
    keys[o] = key;
  }
}

// mpmb_set_material on RESIDENT particles: the cached affine matrix A (q1.w..q3) was computed with the old parameters
// (pack_one at upload, or the last G2P); rebuild it from (F, scalar, vol, apic_b) for the rows of that group so that
// the next rasterize uses the new material's stress (src/transfer.cpp:503,509,521-522).
__global__ void k_refresh_affine(View V, Params P, const uint32_t *keys, int n, uint32_t special_min, int group) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || keys[i] >= special_min) return;
  const float4 q6 = V.q[6][i];
  if ((int)(__float_as_uint(q6.w) >> TAG_ID_BITS) != group) return;
  const float4 q0 = V.q[0][i], q1 = V.q[1][i], q4 = V.q[4][i], q5 = V.q[5][i], q7 = V.q[7][i], q8 = V.q[8][i], q9 = V.q[9][i];
  Mat3 F, b, force, A;
  F.m[0] = q4.x; F.m[1] = q4.y; F.m[2] = q4.z; F.m[3] = q4.w; F.m[4] = q5.x; F.m[5] = q5.y; F.m[6] = q5.z; F.m[7] = q5.w; F.m[8] = q6.x;
  b.m[0] = q7.x; b.m[1] = q7.y; b.m[2] = q7.z; b.m[3] = q7.w; b.m[4] = q8.x; b.m[5] = q8.y; b.m[6] = q8.z; b.m[7] = q8.w; b.m[8] = q9.x;
  calculate_force(P.mats[group], F, q6.y, q6.z, force);
  make_affine(force, b, fabsf(q0.w), -4.0f * P.inv_dx * P.dt, A);
  V.q[1][i] = make_float4(q1.x, q1.y, q1.z, A.m[0]);
  V.q[2][i] = make_float4(A.m[1], A.m[2], A.m[3], A.m[4]);
  V.q[3][i] = make_float4(A.m[5], A.m[6], A.m[7], A.m[8]);
}

// ------------------------------------------------------------------------------ frame dump
// The per-point records of MPM<3>::write_partio -> Partio's BGEO writer (src/visualize.cpp:16-100,
// external/partio/src/io/BGEO.cpp:131-150) packed on the device: big-endian 4-byte words
//   position xyz, w = 1.0f, type = 0 (is_rigid), index = id, limit = (1,1,1), v xyz        (12 words, non-verbose dump)
// in id order, which is the order the reference sorts its particles into before writing (visualize.cpp:39-43).
__device__ __forceinline__ uint32_t be32(uint32_t v) { return __byte_perm(v, 0, 0x0123); }
__global__ void k_id_flags(View V, const uint32_t *keys, int n, uint32_t special_min, uint32_t id_base, int64_t id_range, int *flags, Counters *cnt) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || keys[i] >= special_min) return;
  const uint32_t id = (__float_as_uint(V.q[6][i].w) & TAG_ID_MASK) - id_base;
  if ((int64_t)id >= id_range) { atomicOr(&cnt->error, DEVERR_BAD_INPUT); return; }
  flags[id] = 1;
}
__global__ void k_bgeo_points(View V, const uint32_t *keys, int n, uint32_t special_min, uint32_t id_base, int64_t id_range, const int *prefix, uint32_t *out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || keys[i] >= special_min) return;
  const float4 q0 = V.q[0][i], q1 = V.q[1][i];
  const uint32_t gid = __float_as_uint(V.q[6][i].w) & TAG_ID_MASK, id = gid - id_base;
  if ((int64_t)id >= id_range) return;
  uint4 *rec = reinterpret_cast<uint4 *>(out + (size_t)prefix[id] * 12);
  rec[0] = make_uint4(be32(__float_as_uint(q0.x)), be32(__float_as_uint(q0.y)), be32(__float_as_uint(q0.z)), be32(__float_as_uint(1.0f)));
  rec[1] = make_uint4(0u, be32(gid), be32(1u), be32(1u));
  rec[2] = make_uint4(be32(1u), be32(__float_as_uint(q1.x)), be32(__float_as_uint(q1.y)), be32(__float_as_uint(q1.z)));
}

// ------------------------------------------------------------------------------ AoS write-back
// mpmb_download_aos on the device: every live row goes back into its slot of the (device image of the) reference's
// pool and raises the alive flag of its id; the survivors' slots, in id order, are then one stream compaction away.
__global__ void k_aos_scatter(View V, const uint32_t *keys, int n, uint32_t special_min, unsigned char *pool, const uint32_t *indices,
                              int64_t n_indices, int64_t pool_slots, MpmbAosLayout L, uint32_t id_base, int *alive_flag, Counters *cnt) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (keys[i] >= special_min) return;
  const float4 q0 = V.q[0][i], q1 = V.q[1][i], q4 = V.q[4][i], q5 = V.q[5][i], q6 = V.q[6][i], q7 = V.q[7][i], q8 = V.q[8][i], q9 = V.q[9][i];
  const uint32_t id = (__float_as_uint(q6.w) & TAG_ID_MASK) - id_base;
  if ((int64_t)id >= n_indices) { atomicOr(&cnt->error, DEVERR_BAD_INPUT); return; }
  const uint32_t slot = indices[id];
  if ((int64_t)slot >= pool_slots) { atomicOr(&cnt->error, DEVERR_BAD_INPUT); return; }
  alive_flag[id] = 1;
  unsigned char *sl = pool + (size_t)slot * L.stride;
  float *pos = (float *)(sl + L.off_pos), *vm = (float *)(sl + L.off_v_and_m);
  pos[0] = q0.x; pos[1] = q0.y; pos[2] = q0.z;
  vm[0] = q1.x; vm[1] = q1.y; vm[2] = q1.z; vm[3] = fabsf(q0.w);
  const float Fm[9] = {q4.x, q4.y, q4.z, q4.w, q5.x, q5.y, q5.z, q5.w, q6.x}, bm[9] = {q7.x, q7.y, q7.z, q7.w, q8.x, q8.y, q8.z, q8.w, q9.x};
#pragma unroll
  for (int c = 0; c < 3; c++) {
    float *fc = (float *)(sl + L.off_dg_e + c * L.col_pitch), *bc = (float *)(sl + L.off_apic_b + c * L.col_pitch);
#pragma unroll
    for (int r = 0; r < 3; r++) { fc[r] = Fm[c * 3 + r]; bc[r] = bm[c * 3 + r]; }
  }
  if (L.off_scalar >= 0) *(float *)(sl + L.off_scalar) = q6.y;
}
__global__ void k_compact_survivors(const uint32_t *indices, const int *alive_flag, const int *prefix, int64_t n, uint32_t *out) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n && alive_flag[i]) out[prefix[i]] = indices[i];
}

// ------------------------------------------------------------------------------ ordering
// Replaces sort_particles_and_populate_grid (src/mpm.cpp:770-918) incrementally.
//   (arr_cnt[dst]++ for every mover is done where the mover is appended: k_g2p, the migration unpack kernels)
//   k_order_a/b/c : exclusive scans over the dense tile arrays of (stay+arrivals, arrivals, active)
//                   -> next-run offsets, arrival-segment offsets, compact active-tile list + slot_map
//   k_mover_place : movers into their tile's arrival segment (order arbitrary)
//   k_mover_rank  : segment sorted by storage row => deterministic visiting order
constexpr int ORD_B = 256, ORD_IPT = 4, ORD_TILE = ORD_B * ORD_IPT;

// arrival counted where the mover is appended (k_g2p, the migration unpack kernels): no counting pass in the ordering
#define MPMB_COUNT_ARRIVAL(V, P, key) if ((key) < (uint32_t)(P).ntiles_total) atomicAdd(&(V).arr_cnt[key], 1)
// arrivals per tile straight from radix-sorted keys (upload path: every particle is an arrival)
__global__ void k_count_sorted(View V, const uint32_t *keys_sorted, int n, int ntot) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t t = keys_sorted[i];
  if (t >= (uint32_t)ntot) return;
  if (i > 0 && keys_sorted[i - 1] == t) return;
  int lo = i + 1, hi = n;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (keys_sorted[mid] == t) lo = mid + 1;
    else hi = mid;
  }
  V.arr_cnt[t] = lo - i;
}

__global__ void __launch_bounds__(ORD_B) k_order_a(View V, int ntot) {
  typedef cub::BlockReduce<int, ORD_B> Red;
  __shared__ typename Red::TempStorage tmp;
  int s_tot = 0, s_arr = 0, s_act = 0;
#pragma unroll
  for (int k = 0; k < ORD_IPT; k++) {
    const int t = blockIdx.x * ORD_TILE + threadIdx.x * ORD_IPT + k;
    if (t < ntot) {
      const int a = V.arr_cnt[t], tot = V.stay_cnt[t] + a;
      s_tot += tot; s_arr += a; s_act += tot > 0;
    }
  }
  s_tot = Red(tmp).Sum(s_tot); __syncthreads();
  s_arr = Red(tmp).Sum(s_arr); __syncthreads();
  s_act = Red(tmp).Sum(s_act);
  if (threadIdx.x == 0) { V.blocksum[3 * blockIdx.x] = s_tot; V.blocksum[3 * blockIdx.x + 1] = s_arr; V.blocksum[3 * blockIdx.x + 2] = s_act; }
}

__global__ void __launch_bounds__(1024) k_order_b(View V, int nblocks) {
  // single CTA: exclusive scan of the three block sums (nblocks is a few thousand at most)
  typedef cub::BlockScan<int, 1024> Scan;
  __shared__ typename Scan::TempStorage tmp;
  int carry[3] = {0, 0, 0};
  for (int base = 0; base < nblocks; base += 1024) {
    const int i = base + threadIdx.x;
#pragma unroll
    for (int c = 0; c < 3; c++) {
      int v = i < nblocks ? V.blocksum[3 * i + c] : 0, ex, agg;
      Scan(tmp).ExclusiveSum(v, ex, agg);
      __syncthreads();
      if (i < nblocks) V.blocksum[3 * i + c] = ex + carry[c];
      carry[c] += agg;
    }
  }
  if (threadIdx.x == 0) {
    V.cnt->n_alive = carry[0];
    V.cnt->updates += (unsigned long long)carry[0];
    V.cnt->n_tiles = carry[2];
    V.cnt->n_ghost = 0;
    if (carry[2] > V.cap_tiles) atomicOr(&V.cnt->error, DEVERR_TILE_CAPACITY);
    if (carry[0] > V.cap_particles) atomicOr(&V.cnt->error, DEVERR_PARTICLE_CAPACITY);
  }
}

#define MPMB_TILE_XYZ(P, tm, tx, ty, tz) const int tx = (tm).xy & 0xffff, ty = (tm).xy >> 16, tz = (tm).z
__global__ void __launch_bounds__(ORD_B) k_order_c(View V, int ntot, int nt1, int nt2, int tz_off) {
  typedef cub::BlockScan<int, ORD_B> Scan;
  __shared__ typename Scan::TempStorage tmp;
  int tot[ORD_IPT], arr[ORD_IPT], act[ORD_IPT];
  int s_tot = 0, s_arr = 0, s_act = 0;
#pragma unroll
  for (int k = 0; k < ORD_IPT; k++) {
    const int t = blockIdx.x * ORD_TILE + threadIdx.x * ORD_IPT + k;
    tot[k] = arr[k] = act[k] = 0;
    if (t < ntot) {
      arr[k] = V.arr_cnt[t];
      tot[k] = V.stay_cnt[t] + arr[k];
      act[k] = tot[k] > 0;
    }
    s_tot += tot[k]; s_arr += arr[k]; s_act += act[k];
  }
  int e_tot, e_arr, e_act;
  Scan(tmp).ExclusiveSum(s_tot, e_tot); __syncthreads();
  Scan(tmp).ExclusiveSum(s_arr, e_arr); __syncthreads();
  Scan(tmp).ExclusiveSum(s_act, e_act);
  e_tot += V.blocksum[3 * blockIdx.x];
  e_arr += V.blocksum[3 * blockIdx.x + 1];
  e_act += V.blocksum[3 * blockIdx.x + 2];
#pragma unroll
  for (int k = 0; k < ORD_IPT; k++) {
    const int t = blockIdx.x * ORD_TILE + threadIdx.x * ORD_IPT + k;
    if (t < ntot) {
      V.out_begin[t] = e_tot;
      V.total[t] = tot[k];
      V.arr_off[t] = e_arr;
      V.arr_len[t] = arr[k];
      V.arr_cnt[t] = 0;   // ready for the next substep's count
      V.arr_cur[t] = 0;   // cursor of k_mover_place
      V.stay_next[t] = 0; // G2P writes it for the tiles it visits
      int slot = -1;
      if (act[k] && e_act < V.cap_tiles) {
        slot = e_act;
        TileMeta m;
        m.tile = t; m.run_begin = V.run_begin[t]; m.run_len = V.run_len[t]; m.arr_off = e_arr; m.arr_len = arr[k]; m.out_begin = e_tot;
        m.xy = (t / (nt2 * nt1)) | (((t / nt2) % nt1) << 16); m.z = t % nt2 + tz_off;
        V.meta[slot] = m;
      }
      V.slot_map[t] = slot;
    }
    e_tot += tot[k]; e_arr += arr[k]; e_act += act[k];
  }
}

__global__ void k_mover_place(View V, int ntot) {
  const int n = V.cnt->n_movers;
  for (int m = blockIdx.x * blockDim.x + threadIdx.x; m < n; m += gridDim.x * blockDim.x) {
    const uint32_t d = V.mover_dst[m];
    if (d < (uint32_t)ntot) {
      const int pos = atomicAdd(&V.arr_cur[d], 1);
      if (!MPMB_CHK(V.cnt, pos < V.arr_len[d] && V.arr_off[d] + pos < V.cap_particles && V.mover_idx[m] < (uint32_t)V.cap_particles, 1, d, pos, V.arr_len[d])) continue;
      V.arrivals[V.arr_off[d] + pos] = V.mover_idx[m];
    }
  }
}

__global__ void k_mover_rank(View V, int ntot) {
  const int n = V.cnt->n_movers;
  for (int m = blockIdx.x * blockDim.x + threadIdx.x; m < n; m += gridDim.x * blockDim.x) {
    const uint32_t d = V.mover_dst[m];
    if (d < (uint32_t)ntot) {
      const uint32_t idx = V.mover_idx[m];
      const int off = V.arr_off[d], len = V.arr_len[d];
      int rank = 0;
      for (int e = 0; e < len; e++) rank += V.arrivals[off + e] < idx;
      if (!MPMB_CHK(V.cnt, off + rank < V.cap_particles && rank < len, 2, d, rank, len)) continue;
      V.arrivals_sorted[off + rank] = idx;
    }
  }
}

#ifdef MPMB_VALIDATE
// debug: after the ordering every tile must have received exactly the arrivals that were counted for it, its arrival
// rows must be live storage rows, and the runs must tile the next storage without overlap
__global__ void k_validate_order(View V, int ntot) {
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < ntot; t += gridDim.x * blockDim.x) {
    if (V.arr_cur[t] != V.arr_len[t]) {
      if (atomicCAS(&V.cnt->chk[0], 0, 20) == 0) { V.cnt->chk[1] = t; V.cnt->chk[2] = V.arr_cur[t]; V.cnt->chk[3] = V.arr_len[t]; }
    }
    if (V.run_begin[t] + V.run_len[t] > V.cnt->n_store || V.run_len[t] < V.stay_cnt[t]) {
      if (atomicCAS(&V.cnt->chk[0], 0, 21) == 0) { V.cnt->chk[1] = t; V.cnt->chk[2] = V.run_len[t]; V.cnt->chk[3] = V.stay_cnt[t]; }
    }
    for (int e = 0; e < V.arr_len[t]; e++) {
      const uint32_t row = V.arrivals_sorted[V.arr_off[t] + e];
      if (row >= (uint32_t)V.cnt->n_store || V.keys[row] != (uint32_t)t || (e > 0 && V.arrivals_sorted[V.arr_off[t] + e - 1] >= row)) {
        if (atomicCAS(&V.cnt->chk[0], 0, 22) == 0) { V.cnt->chk[1] = t; V.cnt->chk[2] = (int)row; V.cnt->chk[3] = row < (uint32_t)V.cnt->n_store ? (int)V.keys[row] : -7; }
      }
    }
  }
}
extern "C" int mpmb_debug_check(MpmbHandle h, int *out8);
#endif

// ------------------------------------------------------------------------------ P2G
// Replaces MPM<3>::rasterize_optimized / block_op_normal (src/transfer.cpp:467-569).
// One 64-thread CTA per active tile (persistent round-robin).  A tile's rows are its run (unit
// stride, holes where a particle left) followed by its arrivals.  Per chunk of <=CH rows:
//   1. stage: 16-byte async copies (LDGSTS) of the P2G set (64 B/particle) into padded shared rows,
//      all in flight at once;
//   2. sort:  stable counting sort of the rows by cell (warp match + per-(pass,warp) histograms),
//             bit-reproducible; publishes each particle's row in the next storage (outpos);
//   3. accumulate: thread c owns cell c and streams that cell's particles, accumulating all
//      27 nodes x (p_x,p_y,p_z,m) in 108 REGISTERS — every particle of a cell shares one stencil;
//   4. flush: each warp adds its registers into the shared 6x6x6 arena, conflict-free by layout
//      (node strides 68/8/1 put the 32 cells of a warp on 32 banks), warp after warp;
// then one coalesced store of the arena.  No atomics on floats anywhere.
constexpr int P2G_T = 64;          // threads = cells per tile
// particles staged per chunk: a full interior tile is 512 rows at 8 particles per cell, and in a developed flow a quarter
// of the tiles sit a little above that (profiles/r02_flow_stats.jsonl: p90 526, max ~630 rows) — 576 keeps most of them
// in ONE chunk at 4 CTAs per SM; the per-row passes below skip the empty tail of a short chunk
#ifndef MPMB_P2G_CH
#define MPMB_P2G_CH 576
#endif
constexpr int P2G_CH = MPMB_P2G_CH;
constexpr int P2G_K = P2G_CH / P2G_T;
constexpr int P2G_ROWS = P2G_CH + P2G_CH / 8;  // padded row index r + (r>>3): cell-run reads hit distinct banks
constexpr int P2G_DYN_BYTES = 4 * P2G_ROWS * (int)sizeof(float4);
constexpr int AR_SX = 68, AR_SY = 8;           // arena strides in shared memory
constexpr int AR_SIZE = 6 * AR_SX;

// Occupancy, measured and rejected (round 2, profiles/r02_ab_p2g_occupancy.log): this kernel holds 243 registers and 51.7 KB
// of shared memory, i.e. 4 CTAs = 8 warps per SM by both limits.  Laying BOTH flush arenas over the row staging area (-6.5 KB)
// and bounding the launch to 5 CTAs per SM gives a spill-free 168-register build that fits 5 CTAs with 512-row chunks — and
// is slower everywhere: 0.2628 / 0.3773 ms (at rest / flowing) against 0.2463 / 0.3275; the 168-register code alone, still at
// 4 CTAs with 576-row chunks, 0.2638 / 0.3559.  The time is in-warp latency of the per-particle dependency chain, which the
// wider register allocation schedules around; two more warps do not buy it back.
__global__ void __launch_bounds__(P2G_T) k_p2g(View V, Params P, int part) {
  // the row staging area is dynamic shared memory (static + dynamic = 51.7 KB, above the 48 KB static limit)
#ifndef MPMB_SIMT_HOST
  extern __shared__ __align__(16) unsigned char p2g_dyn[];
#else
  __shared__ __align__(16) unsigned char p2g_dyn[P2G_DYN_BYTES];
#endif
  float4 (*s_rows)[P2G_ROWS] = reinterpret_cast<float4 (*)[P2G_ROWS]>(p2g_dyn);
  __shared__ unsigned short s_order[P2G_CH];
  __shared__ unsigned short s_hist[P2G_K * 2][64];
  __shared__ int s_start[65];
  __shared__ float s_arena[4][AR_SIZE];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n_tiles = V.cnt->n_tiles;
  const int cx = tid >> 4, cy = (tid >> 2) & 3, cz = tid & 3;
  for (int slot = blockIdx.x; slot < n_tiles; slot += gridDim.x) {
    const TileMeta tm = V.meta[slot];
    const int tile = tm.tile;
    if (!tile_in_part(P, tile, part)) continue;  // uniform per CTA
    const int nrow_tile = tm.run_len + tm.arr_len;
    if (!MPMB_CHK(V.cnt, tm.run_begin >= 0 && tm.run_len >= 0 && tm.arr_len >= 0 && tm.run_begin + tm.run_len <= V.cap_particles && tm.arr_off + tm.arr_len <= V.cap_particles &&
                             tm.out_begin + nrow_tile <= V.cap_particles, 3, slot, tm.run_begin, tm.run_len)) continue;
    MPMB_TILE_XYZ(P, tm, tx, ty, tz);
    const float fbx = (float)(tx * 4 + cx), fby = (float)(ty * 4 + cy), fbz = (float)(tz * 4 + cz);
    F4 acc[27];  // (p_x, p_y | p_z, m) of the 27 nodes of my cell's stencil
#pragma unroll
    for (int n = 0; n < 27; n++) acc[n] = f4_zero();
    for (int n = tid; n < AR_SIZE; n += P2G_T) { s_arena[0][n] = 0.f; s_arena[1][n] = 0.f; s_arena[2][n] = 0.f; s_arena[3][n] = 0.f; }
    int vbase = 0;  // valid rows in the chunks already processed
    // storage row of tile-row g: run rows first, then arrivals
    auto row_of = [&](int g) -> uint32_t {
      return g < tm.run_len ? (uint32_t)(tm.run_begin + g) : V.arrivals_sorted[tm.arr_off + (g - tm.run_len)];
    };

    for (int cb = 0; cb < nrow_tile; cb += P2G_CH) {
      const int nrows = min(P2G_CH, nrow_tile - cb);
#pragma unroll
      for (int e = 0; e < P2G_K * 2; e++) s_hist[e][tid] = 0;
      // ---- 1: stage rows: all copies of the chunk in flight at once
      uint32_t pidx[P2G_K];
      if (cb + nrows <= tm.run_len) {  // whole chunk inside the run (the common case): pure arithmetic, no loads
#pragma unroll
        for (int k = 0; k < P2G_K; k++) pidx[k] = (uint32_t)(tm.run_begin + cb + k * P2G_T + tid);
      } else {
#pragma unroll
        for (int k = 0; k < P2G_K; k++) {
          const int r = k * P2G_T + tid;
          pidx[k] = r < nrows ? row_of(cb + r) : 0u;
          if (!MPMB_CHK(V.cnt, pidx[k] < (uint32_t)V.cap_particles, 4, slot, pidx[k], cb + r)) pidx[k] = 0u;
        }
      }
#pragma unroll
      for (int k = 0; k < P2G_K; k++) {
        const int r = k * P2G_T + tid;
        if (r < nrows) {
          const int ri = r + (r >> 3);
          cp_async16(&s_rows[0][ri], &V.q[0][pidx[k]]);
          cp_async16(&s_rows[1][ri], &V.q[1][pidx[k]]);
          cp_async16(&s_rows[2][ri], &V.q[2][pidx[k]]);
          cp_async16(&s_rows[3][ri], &V.q[3][pidx[k]]);
        }
      }
      cp_async_commit();
      cp_async_wait_all();
      __syncthreads();
      // ---- 2a: per-(pass,warp) cell histograms (each thread reads back its own rows)
      uint32_t cr[P2G_K];
#pragma unroll
      for (int k = 0; k < P2G_K; k++) {
        const int r = k * P2G_T + tid;
        int cell = 64 + warp;  // holes and rows past the end: a private bucket
        if (k * P2G_T >= nrows) {  // uniform: nothing staged for this pass
          cr[k] = (uint32_t)cell << 16;
          continue;
        }
        float4 a0 = make_float4(0.f, 0.f, 0.f, -1.f);
        if (r < nrows) a0 = s_rows[0][r + (r >> 3)];
        // G2P stores the mass NEGATED when the particle's base node left the tile of the run it was
        // written to: run rows with mass < 0 are holes (the particle is in another tile's arrivals)
        if (r < nrows && (a0.w > 0.f || cb + r >= tm.run_len)) {
          int bx, by, bz;
          float rr;
          base_rel(a0.x, P.inv_dx, bx, rr);
          base_rel(a0.y, P.inv_dx, by, rr);
          base_rel(a0.z, P.inv_dx, bz, rr);
          cell = (((bx - tx * 4) & 3) << 4) | (((by - ty * 4) & 3) << 2) | ((bz - tz * 4) & 3);
        }
        const unsigned m = __match_any_sync(0xffffffffu, cell);
        const int rank = __popc(m & ((1u << lane) - 1u));
        if (cell < 64) {
          if (rank == 0) s_hist[k * 2 + warp][cell] = (unsigned short)__popc(m);
        }
        cr[k] = ((uint32_t)cell << 16) | (uint32_t)rank;
      }
      __syncthreads();
      // ---- 2b: exclusive scan over (pass,warp) per cell, then over cells
      {
        const int c = tid;
        int run = 0;
#pragma unroll
        for (int e = 0; e < P2G_K * 2; e++) {
          const int h = s_hist[e][c];
          s_hist[e][c] = (unsigned short)run;
          run += h;
        }
        int incl = run;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int y = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl += y;
        }
        if (tid == 31) s_start[64] = incl;  // total of warp 0, fixed up below
        __syncthreads();
        const int base0 = warp ? s_start[64] : 0;
        s_start[c] = base0 + incl - run;
        __syncthreads();
        if (tid == 63) s_start[64] = base0 + incl;
      }
      // ---- 2c: scatter row ids to their sorted position; publish the output row for G2P
#pragma unroll
      for (int k = 0; k < P2G_K; k++) {
        const int r = k * P2G_T + tid;
        const int cell = (int)(cr[k] >> 16);
        if (cell < 64) {
          const int pos = s_start[cell] + s_hist[k * 2 + warp][cell] + (int)(cr[k] & 0xffffu);
          s_order[pos] = (unsigned short)r;
          V.outpos[pidx[k]] = (uint32_t)(tm.out_begin + vbase + pos);
        }
        // holes are NOT marked here: the same storage row is a hole of its old tile and an arrival of
        // its new one, and two CTAs must not write different values to one outpos entry
      }
      __syncthreads();
      // ---- 3: accumulate my cell's run in registers.  A warp iterates as long as its fullest cell has particles; in a
      // developed flow (cells hold 0..30 particles: mean/max per warp 0.56, profiles/r02_flow_stats.jsonl) that is the
      // kernel's cost.  Measured and rejected (round 2): capping the loop and adding the excess particles with one lane
      // per stencil node — 0.403 instead of 0.329 ms, the serialised excess pass costs more than the idle tail.
      // Also measured and rejected (round 2, profiles/r02_ab_p2g_balance.log): re-assigning cells to threads per tile
      // so that warp 0 owns the fuller cells — by bank pairs (flush stays conflict-free) +0.002 ms in both states, by
      // population rank +0.015 / +0.020 ms: the kernel is latency-bound at 8 warps per SM, an early warp's issue slots
      // are not what it lacks.
      const int i0 = s_start[tid], i1 = s_start[tid + 1];
      // (software-pipelining this loop — the next particle's row fetched one iteration ahead, 252 registers, no spills — was
      // measured slower as well: 0.2665 / 0.3644 ms against 0.2465 / 0.3275, profiles/r02_ab_p2g_prefetch.log)
      for (int it = i0; it < i1; it++) {
        const int r = s_order[it];
        const int ri = r + (r >> 3);
        const float4 a0 = s_rows[0][ri], a1 = s_rows[1][ri], a2 = s_rows[2][ri], a3 = s_rows[3][ri];
        const float mass = fabsf(a0.w);
        float vx = a1.x, vy = a1.y, vz = a1.z;
        if (P.particle_gravity) {  // src/transfer.cpp:485-487
          vx += P.gdt[0]; vy += P.gdt[1]; vz += P.gdt[2];
        }
        // rel = pos/dx - base node of this cell (src/transfer.cpp:490,518)
        const float rx = __fsub_rn(__fmul_rn(a0.x, P.inv_dx), fbx), ry = __fsub_rn(__fmul_rn(a0.y, P.inv_dx), fby),
                    rz = __fsub_rn(__fmul_rn(a0.z, P.inv_dx), fbz);
        float wx[3], wy[3], wz[3];
        bspline_weights(rx, wx);
        bspline_weights(ry, wy);
        bspline_weights(rz, wz);
        // A columns: c0 = (a1.w,a2.x,a2.y) c1 = (a2.z,a2.w,a3.x) c2 = (a3.y,a3.z,a3.w)
        // q = m v + A rel ;  node (i,j,k): q - i c0 - j c1 - k c2   (dpos = rel - node, 528-536)
        const float q0 = fmaf(a3.y, rz, fmaf(a2.z, ry, fmaf(a1.w, rx, mass * vx)));
        const float q1 = fmaf(a3.z, rz, fmaf(a2.w, ry, fmaf(a2.x, rx, mass * vy)));
        const float q2 = fmaf(a3.w, rz, fmaf(a3.x, ry, fmaf(a2.y, rx, mass * vz)));
        // (momentum | mass) as register pairs; the mass lane rides along with zero column entries
        const F4 Q = {make_float2(q0, q1), make_float2(q2, mass)};
        const F4 c0 = {make_float2(a1.w, a2.x), make_float2(a2.y, 0.f)}, c1 = {make_float2(a2.z, a2.w), make_float2(a3.x, 0.f)},
                 c2 = {make_float2(a3.y, a3.z), make_float2(a3.w, 0.f)};
        const float2 m2 = bc2(-2.0f);
        const float2 wx2[3] = {bc2(wx[0]), bc2(wx[1]), bc2(wx[2])}, wy2[3] = {bc2(wy[0]), bc2(wy[1]), bc2(wy[2])},
                     wz2[3] = {bc2(wz[0]), bc2(wz[1]), bc2(wz[2])};
#pragma unroll
        for (int i = 0; i < 3; i++) {
          const F4 U = i == 0 ? Q : (i == 1 ? f4_sub(Q, c0) : f4_fma(m2, c0, Q));    // q - i c0
#pragma unroll
          for (int j = 0; j < 3; j++) {
            const F4 T = j == 0 ? U : (j == 1 ? f4_sub(U, c1) : f4_fma(m2, c1, U));  // ... - j c1
            const float2 wij = mul2(wx2[i], wy2[j]);
#pragma unroll
            for (int k = 0; k < 3; k++) {
              const float2 w = mul2(wij, wz2[k]);
              const F4 N = k == 0 ? T : (k == 1 ? f4_sub(T, c2) : f4_fma(m2, c2, T));  // ... - k c2
              acc[i * 9 + j * 3 + k] = f4_fma(w, N, acc[i * 9 + j * 3 + k]);
            }
          }
        }
      }
      vbase += s_start[64];
      __syncthreads();  // rows / order / hist are reused by the next chunk
    }
    // ---- 4: both warps flush at once, warp 0 into s_arena, warp 1 into a second arena laid over the
    // row staging area (free after the last chunk's barrier; zeroed by warp 1 itself, so a __syncwarp
    // is all it needs).  The store below sums the two in a fixed order: still bit-reproducible.
    float (*ar1)[AR_SIZE] = reinterpret_cast<float (*)[AR_SIZE]>(&s_rows[0][0]);
    {
      if (warp == 1) {
        for (int n = lane; n < 4 * AR_SIZE; n += 32) (&ar1[0][0])[n] = 0.f;
        __syncwarp();
      }
      float (*ar)[AR_SIZE] = warp == 0 ? s_arena : ar1;
      const int nb = cx * AR_SX + cy * AR_SY + cz;
#pragma unroll
      for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++)
#pragma unroll
          for (int k = 0; k < 3; k++) {
            const int node = nb + i * AR_SX + j * AR_SY + k;
            const F4 &a = acc[i * 9 + j * 3 + k];
            ar[0][node] += a.lo.x;
            ar[1][node] += a.lo.y;
            ar[2][node] += a.hi.x;
            ar[3][node] += a.hi.y;
            __syncwarp();
          }
    }
    __syncthreads();
    float4 *out = V.arena + (size_t)slot * ARENA;
    for (int n = tid; n < ARENA; n += P2G_T) {
      const int a = n / 36, b = (n / 6) % 6, c = n % 6;
      const int node = a * AR_SX + b * AR_SY + c;
      out[n] = make_float4(s_arena[0][node] + ar1[0][node], s_arena[1][node] + ar1[1][node], s_arena[2][node] + ar1[2][node],
                           s_arena[3][node] + ar1[3][node]);
    }
    __syncthreads();
  }
}


// ------------------------------------------------------------------------------ grid node
// Momentum/mass of global node g = fixed-order sum of the arenas that cover it:
// owner tile T=(g>>2) holds it at local l=g&3; tile T-o (o in {0,1}^3) holds it at l+4o (needs l<=1).
// The <=8 loads are predicated, not branched around, so they stay independent and issue back to back.
template <class SlotOf>
__device__ __forceinline__ float4 gather_node(const float4 *arena, int zero_slot, int lx, int ly, int lz, SlotOf slot_of) {
  (void)zero_slot;
  float4 a[8];
#pragma unroll
  for (int o = 0; o < 8; o++) {
    const int ox = o >> 2, oy = (o >> 1) & 1, oz = o & 1;
    const int slot = slot_of(ox, oy, oz);
    const bool covered = !(ox && lx > 1) && !(oy && ly > 1) && !(oz && lz > 1) && slot >= 0;
    const int idx = ((lx + 4 * ox) * 6 + (ly + 4 * oy)) * 6 + (lz + 4 * oz);
    a[o] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (covered) a[o] = arena[(size_t)slot * ARENA + idx];
  }
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int o = 0; o < 8; o++) { acc.x += a[o].x; acc.y += a[o].y; acc.z += a[o].z; acc.w += a[o].w; }
  return acc;
}

// normalize_grid_and_apply_external_force (src/mpm.cpp:277-294) + apply_grid_boundary_conditions
// (src/mpm.cpp:296-372, static level set) for one node.
__device__ __forceinline__ float4 node_update(const Params &P, const float4 *sdf4, float4 g, int gx, int gy, int gz, bool use_sdf = true) {
  float m = g.w;
  if (m > 0.f) {
    float inv = 1.0f / m;
    float ix = P.particle_gravity ? 0.f : P.gdt[0], iy = P.particle_gravity ? 0.f : P.gdt[1], iz = P.particle_gravity ? 0.f : P.gdt[2];
    g.x = fmaf(g.x, inv, ix);
    g.y = fmaf(g.y, inv, iy);
    g.z = fmaf(g.z, inv, iz);
  }
  if (P.has_sdf && use_sdf && m != 0.f && gx < P.nnode[0] && gy < P.nnode[1] && gz < P.nnode[2]) {
    float4 s = sdf4[((size_t)gx * P.nnode[1] + gy) * P.nnode[2] + gz];
    if (!(s.w < -3.0f || 0.0f < s.w)) {
      float3 v = friction_project0(make_float3(g.x, g.y, g.z), make_float3(s.x, s.y, s.z), P.friction);
      g.x = v.x; g.y = v.y; g.z = v.z;
    }
  }
  return g;
}

// ------------------------------------------------------------------------------ grid update
// Replaces normalize_grid_and_apply_external_force (src/mpm.cpp:277-294) + apply_grid_boundary_conditions
// (src/mpm.cpp:296-372): for every active tile the 6x6x6 nodes its particles gather from are
// rebuilt (fixed-order sum of the covering arenas), normalised, projected on the level set and
// stored as one contiguous 3.4 KB block vel[slot][216].  Node-parallel, so the dependent
// slot_map -> arena -> sdf loads are hidden by occupancy instead of stalling a G2P CTA.
// launch bound (256, 1): 64 registers, 0.058 ms; capping at 40 / 32 registers for 6 / 8 CTAs per SM spills and is not
// faster (0.064 / 0.056 ms, profiles/r02_ab_grid_occupancy.log)
__global__ void __launch_bounds__(256, 1) k_grid(View V, Params P, float4 *vel, int part) {
  // one WARP per tile: the 27 neighbour slots live in lanes 0..26 and are fetched with shuffles, so
  // there is no block barrier and every warp of the grid has its own tile in flight
  const int lane = threadIdx.x & 31;
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  const int n_tiles = V.cnt->n_tiles;
  for (int slot = gw; slot < n_tiles; slot += nw) {
    const TileMeta tmg = V.meta[slot];
    const int tile = tmg.tile;
    if (!tile_in_part(P, tile, part)) continue;  // uniform per warp
    MPMB_TILE_XYZ(P, tmg, tx, ty, tz);
    int my_nb = -1;
    if (lane < 27) {
      int ox = lane / 9 - 1, oy = (lane / 3) % 3 - 1, oz = lane % 3 - 1;
      int x = tx + ox, y = ty + oy, z = tz + oz - P.tz_off;  // z: local layer
      if (x >= 0 && y >= 0 && z >= 0 && x < P.nt[0] && y < P.nt[1] && z < P.nt[2]) my_nb = V.slot_map[(x * P.nt[1] + y) * P.nt[2] + z];
      if (!MPMB_CHK(V.cnt, my_nb >= -1 && my_nb < V.cap_tiles, 5, slot, my_nb, lane)) my_nb = -1;
    }
#pragma unroll 2
    for (int n0 = 0; n0 < ARENA; n0 += 32) {
      const int n = min(n0 + lane, ARENA - 1);  // the last pass is partial: clamp, store guarded
      int a = n / 36, b = (n / 6) % 6, c = n % 6;
      int wx_ = a >> 2, wy_ = b >> 2, wz_ = c >> 2;  // owner tile offset (0/1)
      int lx = a & 3, ly = b & 3, lz = c & 3;
      int sl[8];
#pragma unroll
      for (int o = 0; o < 8; o++) {
        const int ox = o >> 2, oy = (o >> 1) & 1, oz = o & 1;
        sl[o] = __shfl_sync(0xffffffffu, my_nb, (wx_ - ox + 1) * 9 + (wy_ - oy + 1) * 3 + (wz_ - oz + 1));
      }
      float4 g = gather_node(V.arena, V.cap_tiles, lx, ly, lz, [&](int ox, int oy, int oz) { return sl[ox * 4 + oy * 2 + oz]; });
      g = node_update(P, V.sdf4, g, tx * 4 + a, ty * 4 + b, tz * 4 + c);
      if (n0 + lane < ARENA) vel[(size_t)slot * ARENA + n] = g;
    }
  }
}

// ------------------------------------------------------------------------------ G2P
// Replaces MPM<3>::resample_optimized / block_op_normal (src/transfer.cpp:837-954) + Particle::plasticity +
// clear_boundary_particles (src/mpm.cpp:583-633); produces the affine matrix of the NEXT rasterize
// (calculate_force of the updated state), writes the tile's particles cell-sorted into one contiguous
// run of the other buffer, and appends the particles whose base node left the tile to the mover list.
// Persistent CTAs walk their tiles chunk by chunk through a two-stage cp.async pipeline: while the
// warps compute chunk i out of one shared buffer, the rows (64 B/particle) of chunk i+1 and, at a
// tile change, that tile's 216 node velocities are already in flight into the other.
constexpr int G2P_CH = 256;  // rows per pipeline stage
// STORE_B = false skips the apic_b streams (48 B/particle): no kernel reads apic_b — rasterize uses the
// affine matrix A — so inside mpmb_substep(h, n) only the LAST substep has to leave it behind for the
// host (downloads, visualize, save).
// EXT_MATS = false leaves the elastic / von Mises / visco branches of material_step out (see mpmb_math.cuh).
template <int BLOCK, bool STORE_B, bool EXT_MATS>
__global__ void __launch_bounds__(BLOCK, 5) k_g2p(View V, Params P, const float4 *vel, int part, int commit) {
  __shared__ __align__(16) float4 s_vel[2][ARENA];
  __shared__ __align__(16) float4 s_in[2][4][G2P_CH];
  // output row of every staged row: run rows at [sh + r] (sh = 16-byte alignment shift of the run's outpos words),
  // arrival rows at [r + 8] — clear of the up to 3 words by which the bulk copy of the run part is rounded up
  __shared__ __align__(16) uint32_t s_out[2][G2P_CH + 8];
  __shared__ __align__(8) uint64_t s_bar[2];               // one mbarrier per pipeline stage
  // particles of the current tile that stay in it.  TWO slots, alternating per tile: thread 0 reads and clears a tile's
  // slot after the tile's last barrier while the other warps are already adding to the next tile's slot — with one slot
  // (round 1, when a barrier at the top of the loop covered it) a late warp 0 folded the next tile's first additions
  // into this tile's count: stay counts drifted, runs overlapped (found on a B200 with -DMPMB_VALIDATE, round 2)
  __shared__ int s_stay[2];
  constexpr int KPT = G2P_CH / BLOCK;
  const int tid = threadIdx.x;
  const int n_tiles = V.cnt->n_tiles;
  const float scale = -4.0f * P.inv_dx * P.dt;  // src/transfer.cpp:938
  struct Item { int slot, rb, first; TileMeta tm; };
  auto first_item = [&](int slot) {
    Item it;
    it.rb = 0; it.first = 1;
    if (part != 0)  // next tile of this CTA's round-robin share that belongs to the part
      while (slot < n_tiles && !tile_in_part(P, V.meta[slot].tile, part)) slot += (int)gridDim.x;
    if (V.rigid_flag)  // CPIC scenes: tiles of rigid pages belong to k_g2p_rigid
      while (slot < n_tiles && V.rigid_flag[slot]) slot += (int)gridDim.x;
    it.slot = slot;
    if (slot < n_tiles) it.tm = V.meta[slot];
    else {
      it.tm.run_len = 0; it.tm.arr_len = 0; it.tm.tile = 0; it.tm.run_begin = 0; it.tm.arr_off = 0; it.tm.out_begin = 0;
      it.tm.xy = 0; it.tm.z = 0;
    }
    return it;
  };
  auto next_item = [&](const Item &it) {
    if (it.rb + G2P_CH < it.tm.run_len + it.tm.arr_len) { Item n = it; n.rb += G2P_CH; n.first = 0; return n; }
    return first_item(it.slot + (int)gridDim.x);
  };
  // rows of the chunk that lie inside the tile's run (unit stride in storage) / alignment shift of their outpos words
  auto run_rows = [&](const Item &it) { return max(0, min(G2P_CH, it.tm.run_len - it.rb)); };
  auto out_shift = [&](const Item &it) { return run_rows(it) > 0 ? (int)((uint32_t)(it.tm.run_begin + it.rb) & 3u) : 0; };
  // puts one item in flight: rows of the G2P set (x|mass, F, scalar|vol|tag), their output rows and, for the first
  // chunk of a tile, the tile's node velocities.  The run part of the chunk and the node block are TMA bulk copies
  // issued by thread 0 and tracked by the stage's mbarrier; the (few) arrival rows are gathered with 16-byte
  // LDGSTS copies by the thread that will consume them, tracked by its own copy groups.
  auto prefetch = [&](const Item &it, int buf, int vbuf) {
    if (it.slot < n_tiles) {
      const int nrow_tile = it.tm.run_len + it.tm.arr_len;
      const int nrows = min(G2P_CH, nrow_tile - it.rb);
      const int nrun = run_rows(it);
      const int sh = out_shift(it);
      if (tid == 0) {
        const uint32_t r0 = (uint32_t)(it.tm.run_begin + it.rb);
        const unsigned row_bytes = (unsigned)nrun * 16u, out_bytes = nrun > 0 ? (((unsigned)(sh + nrun) * 4u + 15u) & ~15u) : 0u;
        mbar_arrive_expect_tx(&s_bar[buf], 4u * row_bytes + out_bytes + (it.first ? (unsigned)(ARENA * sizeof(float4)) : 0u));
        if (nrun > 0) {
          bulk_g2s(&s_in[buf][0][0], &V.q[0][r0], row_bytes, &s_bar[buf]);
          bulk_g2s(&s_in[buf][1][0], &V.q[4][r0], row_bytes, &s_bar[buf]);
          bulk_g2s(&s_in[buf][2][0], &V.q[5][r0], row_bytes, &s_bar[buf]);
          bulk_g2s(&s_in[buf][3][0], &V.q[6][r0], row_bytes, &s_bar[buf]);
          bulk_g2s(&s_out[buf][0], &V.outpos[r0 - (uint32_t)sh], out_bytes, &s_bar[buf]);
        }
        if (it.first) bulk_g2s(&s_vel[vbuf][0], vel + (size_t)it.slot * ARENA, (unsigned)(ARENA * sizeof(float4)), &s_bar[buf]);
      }
      if (nrun < nrows) {
#pragma unroll
        for (int k = 0; k < KPT; k++) {
          const int r = k * BLOCK + tid;
          if (r >= nrun && r < nrows) {
            uint32_t pidx = V.arrivals_sorted[it.tm.arr_off + (it.rb + r - it.tm.run_len)];
            if (!MPMB_CHK(V.cnt, pidx < (uint32_t)V.cap_particles, 6, it.slot, pidx, r)) pidx = 0u;
            cp_async16(&s_in[buf][0][r], &V.q[0][pidx]);
            cp_async16(&s_in[buf][1][r], &V.q[4][pidx]);
            cp_async16(&s_in[buf][2][r], &V.q[5][pidx]);
            cp_async16(&s_in[buf][3][r], &V.q[6][pidx]);
            unsigned d = (unsigned)__cvta_generic_to_shared(&s_out[buf][r + 8]);
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(d), "l"(&V.outpos[pidx]));
          }
        }
      }
    }
    cp_async_commit();
  };
  if (tid == 0) {
    mbar_init(&s_bar[0], 1);
    mbar_init(&s_bar[1], 1);
    mbar_init_fence();
    s_stay[0] = 0;
    s_stay[1] = 0;
  }
  __syncthreads();
  Item cur = first_item(blockIdx.x);
  int buf = 0, vbuf = 0, sp = 0;
  unsigned phase = 0;  // bit b = parity of the phase stage b's barrier is in
  prefetch(cur, 0, 0);
  while (cur.slot < n_tiles) {
    const Item nxt = next_item(cur);
    const int nvbuf = nxt.first ? (vbuf ^ 1) : vbuf;
    prefetch(nxt, buf ^ 1, nvbuf);
    asm volatile("cp.async.wait_group 1;\n" ::: "memory");  // my arrival rows of the current item have landed; the next item stays in flight
    mbar_wait(&s_bar[buf], (phase >> buf) & 1u);              // ... and so have its bulk copies (every thread reads only rows it waited for)
    phase ^= 1u << buf;
    {
      const int tile = cur.tm.tile;
      const int sh = out_shift(cur), nrun = run_rows(cur);
      MPMB_TILE_XYZ(P, cur.tm, tx, ty, tz);
      const int nrows = min(G2P_CH, cur.tm.run_len + cur.tm.arr_len - cur.rb);
      const float4 *sv = s_vel[vbuf];
      int my_stay = 0;
      for (int r = tid; r < nrows; r += BLOCK) {
        const float4 q0 = s_in[buf][0][r];
        // run rows with negative mass are holes (the particle now belongs to another tile's arrivals);
        // same rule as k_p2g, so both kernels visit exactly the same particles
        if (cur.rb + r < cur.tm.run_len && !(q0.w > 0.f)) continue;
        const float mass = fabsf(q0.w);
        int bx, by, bz;
        float rx, ry, rz;
        base_rel(q0.x, P.inv_dx, bx, rx);
        base_rel(q0.y, P.inv_dx, by, ry);
        base_rel(q0.z, P.inv_dx, bz, rz);
        bx -= tx * 4; by -= ty * 4; bz -= tz * 4;
        if (!MPMB_CHK(V.cnt, (unsigned)bx < 4u && (unsigned)by < 4u && (unsigned)bz < 4u, 8, cur.slot, (bx & 255) | ((by & 255) << 8) | ((bz & 255) << 16), cur.rb + r)) continue;
        float wx[3], wy[3], wz[3];
        bspline_weights(rx, wx);
        bspline_weights(ry, wy);
        bspline_weights(rz, wz);
        // v = sum w g ; b = sum w g (x) (rel - node) = v (x) rel - [sum_i i S_i | sum_j j T_j | sum_k k R_k]
        // (src/transfer.cpp:888-904), evaluated slab by slab: ~250 FMA instead of 27*16.
        // every node value (v_x,v_y | v_z,m) as two pairs; the mass lane of the sums is never read
        const float2 wz0 = bc2(wz[0]), wz1 = bc2(wz[1]), wz2 = bc2(wz[2]), wz2x = bc2(2.0f * wz[2]);
        const float2 wy2[3] = {bc2(wy[0]), bc2(wy[1]), bc2(wy[2])}, wx2[3] = {bc2(wx[0]), bc2(wx[1]), bc2(wx[2])};
        const float2 wyj[3] = {wy2[0], wy2[1], bc2(2.0f * wy[2])}, wxi[3] = {wx2[0], wx2[1], bc2(2.0f * wx[2])};
        F4 V4 = f4_zero(), CX = f4_zero(), CY = f4_zero(), CZ = f4_zero();
#pragma unroll
        for (int i = 0; i < 3; i++) {
          F4 pi = f4_zero(), jy = f4_zero(), kz = f4_zero();
#pragma unroll
          for (int jn = 0; jn < 3; jn++) {
            const int row = ((bx + i) * 6 + (by + jn)) * 6 + bz;
            const F4 g0 = f4_load(sv[row]), g1 = f4_load(sv[row + 1]), g2 = f4_load(sv[row + 2]);
            const F4 a = f4_fma(wz2, g2, f4_fma(wz1, g1, f4_mul(wz0, g0)));
            const F4 c = f4_fma(wz2x, g2, f4_mul(wz1, g1));
            pi = f4_fma(wy2[jn], a, pi);
            kz = f4_fma(wy2[jn], c, kz);
            if (jn > 0) jy = f4_fma(wyj[jn], a, jy);
          }
          V4 = f4_fma(wx2[i], pi, V4);
          CY = f4_fma(wx2[i], jy, CY);
          CZ = f4_fma(wx2[i], kz, CZ);
          if (i > 0) CX = f4_fma(wxi[i], pi, CX);
        }
        const float3 v = make_float3(V4.lo.x, V4.lo.y, V4.hi.x), colx = make_float3(CX.lo.x, CX.lo.y, CX.hi.x),
                     coly = make_float3(CY.lo.x, CY.lo.y, CY.hi.x), colz = make_float3(CZ.lo.x, CZ.lo.y, CZ.hi.x);
        // the rest of the particle's row is read only now: nothing of it is live across the gather (registers)
        asm volatile("" ::: "memory");
        const float4 q4 = s_in[buf][1][r], q5 = s_in[buf][2][r], q6 = s_in[buf][3][r];
        const size_t o = s_out[buf][r < nrun ? sh + r : r + 8];
        if (!MPMB_CHK(V.cnt, o < (size_t)V.cap_particles && (int)o >= cur.tm.out_begin && (int)o < cur.tm.out_begin + cur.tm.run_len + cur.tm.arr_len, 7, cur.slot, o, r)) continue;
        const float vol = q6.z;
        const uint32_t tag = __float_as_uint(q6.w);
        const Material &mat = P.mats[tag >> TAG_ID_BITS];
        Mat3 B;
        B.m[0] = fmaf(v.x, rx, -colx.x); B.m[1] = fmaf(v.y, rx, -colx.y); B.m[2] = fmaf(v.z, rx, -colx.z);
        B.m[3] = fmaf(v.x, ry, -coly.x); B.m[4] = fmaf(v.y, ry, -coly.y); B.m[5] = fmaf(v.z, ry, -coly.z);
        B.m[6] = fmaf(v.x, rz, -colz.x); B.m[7] = fmaf(v.y, rz, -colz.y); B.m[8] = fmaf(v.z, rz, -colz.z);
        Mat3 cdg;  // cdg = I + (-4 inv_dx dt) b   (src/transfer.cpp:938-942)
#pragma unroll
        for (int k = 0; k < 9; k++) cdg.m[k] = fmaf(scale, B.m[k], (k % 4 == 0) ? 1.f : 0.f);
        Mat3 F;
        F.m[0] = q4.x; F.m[1] = q4.y; F.m[2] = q4.z; F.m[3] = q4.w; F.m[4] = q5.x; F.m[5] = q5.y; F.m[6] = q5.z; F.m[7] = q5.w; F.m[8] = q6.x;
        float ps = q6.y;
        Mat3 force, A;
        material_step<EXT_MATS>(mat, cdg, F, ps, vol, force);
        make_affine(force, B, mass, scale, A);
        float3 x = make_float3(fmaf(v.x, P.dt, q0.x), fmaf(v.y, P.dt, q0.y), fmaf(v.z, P.dt, q0.z));  // 951
        uint32_t key = make_key(P, x.x, x.y, x.z);
        if (P.clean_boundary && reference_deletes(P, x, v)) key = (uint32_t)(P.ntiles_total + SPECIAL_DEAD);
        if (!isfinite(x.x + x.y + x.z)) key = (uint32_t)(P.ntiles_total + SPECIAL_DEAD);
        // one contiguous, cell-sorted run per tile in the other buffer
        store_particle<STORE_B>(V.qn, o, x, key == (uint32_t)tile ? mass : -mass, v, A, F, ps, vol, tag, B);
        V.keys_next[o] = key;
        if (key == (uint32_t)tile) {
          my_stay++;
        } else if (key != (uint32_t)(P.ntiles_total + SPECIAL_DEAD)) {
          const int m = atomicAdd(&V.cnt->n_movers_next, 1);  // ~0.5 % of the particles per substep
          if (!MPMB_CHK(V.cnt, m < V.cap_particles, 9, m, key, 0)) continue;
          V.mover_dst_n[m] = key;
          V.mover_idx_n[m] = (uint32_t)o;
          MPMB_COUNT_ARRIVAL(V, P, key);  // the next ordering's arrival count of that tile
        }
      }
      if (my_stay) atomicAdd(&s_stay[sp], my_stay);
    }
    __syncthreads();  // everyone is done with `buf` before the prefetch after next overwrites it
    if (nxt.slot != cur.slot) {
      if (tid == 0) { V.stay_next[cur.tm.tile] = s_stay[sp]; s_stay[sp] = 0; }
      sp ^= 1;
    }
    cur = nxt;
    buf ^= 1;
    vbuf = nvbuf;
  }
  cp_async_wait_all();
  // end of the substep: the lists and runs this kernel produced become the current ones — done by the last CTA to
  // finish (no separate 1-thread launch)
  if (commit) {
    __syncthreads();
    if (tid == 0) {
      __threadfence();
      if (atomicAdd(&V.cnt->g2p_done, 1) == (int)gridDim.x - 1) {
        __threadfence();
        Counters *c = V.cnt;
        c->n_movers = atomicAdd(&c->n_movers_next, 0);
        c->n_movers_next = 0;
        c->n_store = c->n_alive;  // G2P wrote one row per binned particle
        c->xstep += 1;
        c->g2p_done = 0;
      }
    }
  }
}

// ------------------------------------------------------------------------------ CPIC rigid coupling
// Replaces, for scenes with rigid bodies: update_rigid_page_map (src/mpm.cpp:1026-1076), rasterize_rigid_boundary and
// gather_cdf (src/rigid_transfer.cpp:18-113, 120-274) and the block_op_rigid branches of rasterize_optimized /
// resample_optimized (src/transfer.cpp:367-463, 706-835).  The rigid bodies themselves (RigidBody, its integrator, the
// rigid-rigid solver, articulation) belong to the reference's un-vendored core and stay on the host: the engine takes poses and
// velocities (mpmb_set_rigid_state), applies the impulses of both transfers as apply_tmp_impulse / apply_tmp_velocity would
// (assumptions stated in include/mpmb.h) and hands the velocities back (mpmb_get_rigid_state).
// Layout: RigidBoundaryParticles are NOT MPM particles here but a sample list (anchor offset + untransformed triangle + body);
// node colour data is two dense node arrays (a 64-bit key (distance bits, body) taken with atomicMin, the tags with atomicOr),
// cleared sample by sample; particle colour state is kept by particle id.  Tiles whose 4x4x8 page is a rigid page are
// skipped by k_g2p and processed by k_g2p_rigid; their arenas are rewritten by k_p2g_rigid after k_p2g (which still provides
// the output rows).  This is the reference's slow path too: simple kernels, shared-memory float atomics, no pipelining.
constexpr int RIGID_MAX = 12;                     // GridState::max_num_rigid_bodies (src/mpm_fwd.h:79)
constexpr uint32_t CDF_STATE_MASK = 0xAAAAAAAAu;  // state_mask (src/mpm.h:36)
constexpr uint32_t CDF_TAG_MASK = 0x00FFFFFFu;    // GridState::tag_mask
constexpr unsigned long long CDF_NO_KEY = ~0ull;

struct RigidDev {  // one body on the device
  float pos[3], rot[9], vel[3], ang[3], inv_mass, inv_inertia[9], fric[2];
  float acc[6];    // impulses of the running transfer: sum j, sum (p - pos) x j
};
struct RigidView {
  RigidDev *bodies;
  int n_bodies;
  int n_samples;
  const float *s_offset, *s_tri;
  const int *s_rigid;
  int *s_base;                     // packed base node of each sample's last rasterisation (-1: none), for the clear
  unsigned long long *node_key;    // (distance bits << 32) | (body + 1); CDF_NO_KEY = no rigid boundary near this node
  uint32_t *node_tags;             // GridState tag bits
  unsigned char *page;             // rigid_page_map over 4x4x8 blocks
  int nb[3];
  unsigned char *tile_flag;        // [slot]
  uint32_t *p_states;              // MPMParticle::states by particle id - id_base (persists)
  float4 *p_cdf;                   // (boundary_normal, boundary_distance) of this substep
  uint32_t *p_mark;                // (epoch << 1) | near_boundary_: p_cdf / near of a particle count only if its mark carries THIS substep's
                                   // epoch — gather_cdf then touches only the particles of rigid pages (no 8 M scattered zero-writes)
  uint32_t epoch;
  uint32_t id_base;
  int id_cap;
  float penalty, pushing_force;
};

__device__ __forceinline__ uint32_t cdf_word(const RigidView &R, size_t node) {  // GridState::states: tags | (body + 1) << 24
  const unsigned long long k = R.node_key[node];
  return (R.node_tags[node] & CDF_TAG_MASK) | (k == CDF_NO_KEY ? 0u : ((uint32_t)k & 0xffu) << 24);
}
__device__ __forceinline__ float cdf_dist_world(const RigidView &R, size_t node, float dx) {  // GridState::distance after line 77-78
  const unsigned long long k = R.node_key[node];
  return k == CDF_NO_KEY ? 0.f : __uint_as_float((uint32_t)(k >> 32)) * dx;
}
__device__ __forceinline__ bool cdf_compatible(uint32_t word, uint32_t pst) {  // src/transfer.cpp:416-420
  const uint32_t gs = word & CDF_TAG_MASK;
  const uint32_t mask = (gs & pst & CDF_STATE_MASK) >> 1;
  return (gs & mask) == (pst & mask);
}
__device__ __forceinline__ float3 rigid_to_world(const RigidDev &b, const float *l) {
  return make_float3(b.pos[0] + (b.rot[0] * l[0] + b.rot[3] * l[1] + b.rot[6] * l[2]), b.pos[1] + (b.rot[1] * l[0] + b.rot[4] * l[1] + b.rot[7] * l[2]),
                     b.pos[2] + (b.rot[2] * l[0] + b.rot[5] * l[1] + b.rot[8] * l[2]));
}
__device__ __forceinline__ float3 rigid_velocity_at(const RigidDev &b, float3 p) {  // velocity + angular_velocity x (p - position)
  const float dx = p.x - b.pos[0], dy = p.y - b.pos[1], dz = p.z - b.pos[2];
  return make_float3(b.vel[0] + (b.ang[1] * dz - b.ang[2] * dy), b.vel[1] + (b.ang[2] * dx - b.ang[0] * dz), b.vel[2] + (b.ang[0] * dy - b.ang[1] * dx));
}
// Per-thread impulse accumulator for ONE body at a time: every incompatible (particle, node) pair of a tile otherwise hits the
// same six shared floats (a 128-way contended CAS loop: 0.97 ms of k_p2g_rigid at config 3's size, profiles/r02g_rigid_cost.json);
// a thread flushes only when the body changes and at the end of its rows.
struct ImpulseAcc {
  float a[6];
  int id;
  __device__ __forceinline__ void init() { id = -1; a[0] = a[1] = a[2] = a[3] = a[4] = a[5] = 0.f; }
  __device__ __forceinline__ void flush(float (*s_acc)[6]) {
    if (id >= 0) {
#pragma unroll
      for (int k = 0; k < 6; k++) if (a[k] != 0.f) atomicAdd(&s_acc[id][k], a[k]);
    }
    a[0] = a[1] = a[2] = a[3] = a[4] = a[5] = 0.f;
  }
  __device__ __forceinline__ void add(float (*s_acc)[6], int rid, const RigidDev &b, float3 j, float3 p) {
    if (rid != id) { flush(s_acc); id = rid; }
    const float dx = p.x - b.pos[0], dy = p.y - b.pos[1], dz = p.z - b.pos[2];
    a[0] += j.x; a[1] += j.y; a[2] += j.z;
    a[3] += dy * j.z - dz * j.y; a[4] += dz * j.x - dx * j.z; a[5] += dx * j.y - dy * j.x;
  }
};
// friction_project (src/mpm_fwd.h:25-57) against a moving base
__device__ __forceinline__ float3 friction_project_rel(float3 v, float3 base, float3 n, float friction) {
  const float3 r = friction_project0(make_float3(v.x - base.x, v.y - base.y, v.z - base.z), n, friction);
  return make_float3(r.x + base.x, r.y + base.y, r.z + base.z);
}
__device__ __forceinline__ int pack_base(int bx, int by, int bz) { return (bx << 20) | (by << 10) | bz; }   // node coordinates < 1024

// nodes of the previous rasterisation back to "no boundary"
__global__ void k_cdf_clear(RigidView R, Params P) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= R.n_samples) return;
  const int pb = R.s_base[s];
  if (pb < 0) return;
  const int bx = pb >> 20, by = (pb >> 10) & 1023, bz = pb & 1023;
  for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) for (int c = 0; c < 3; c++) {
    const int i = bx + a, j = by + b, k = bz + c;
    if (i >= P.nnode[0] || j >= P.nnode[1] || k >= P.nnode[2]) continue;
    const size_t node = ((size_t)i * P.nnode[1] + j) * P.nnode[2] + k;
    R.node_key[node] = CDF_NO_KEY;
    R.node_tags[node] = 0u;
  }
}

// update_rigid_page_map + rasterize_rigid_boundary, one thread per RigidBoundaryParticle
__global__ void k_cdf_raster(RigidView R, Params P) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= R.n_samples) return;
  const int id = R.s_rigid[s];
  const RigidDev &B = R.bodies[id];
  const float3 pos = rigid_to_world(B, R.s_offset + 3 * s);   // align_with_rigid_body (src/boundary_particle.h:48-53)
  int bx, by, bz;
  float rr;
  base_rel(pos.x, P.inv_dx, bx, rr);
  base_rel(pos.y, P.inv_dx, by, rr);
  base_rel(pos.z, P.inv_dx, bz, rr);
  if (bx < 0 || by < 0 || bz < 0 || bx + 2 >= P.nnode[0] || by + 2 >= P.nnode[1] || bz + 2 >= P.nnode[2]) { R.s_base[s] = -1; return; }
  R.s_base[s] = pack_base(bx, by, bz);
  // page of the block the particle is sorted into, and (src/mpm.cpp:1062: the range test is on the OFFSET) its {0,1}^3 upper neighbours
  for (int i = 0; i < 2; i++) for (int j = 0; j < 2; j++) for (int k = 0; k < 2; k++) {
    const int x = (bx >> 2) + i, y = (by >> 2) + j, z = (bz >> 3) + k;
    if (x < R.nb[0] && y < R.nb[1] && z < R.nb[2]) R.page[((size_t)x * R.nb[1] + y) * R.nb[2] + z] = 1;
  }
  const float3 v0 = rigid_to_world(B, R.s_tri + 9 * s), v1 = rigid_to_world(B, R.s_tri + 9 * s + 3), v2 = rigid_to_world(B, R.s_tri + 9 * s + 6);
  // world_to_element: [v1 - v0, v2 - v0, n]^-1
  Mat3 M, Mi;
  const float e1[3] = {v1.x - v0.x, v1.y - v0.y, v1.z - v0.z}, e2[3] = {v2.x - v0.x, v2.y - v0.y, v2.z - v0.z};
  float n[3] = {e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2], e1[0] * e2[1] - e1[1] * e2[0]};
  const float nl = sqrtf(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
  for (int k = 0; k < 3; k++) { n[k] /= nl; M(k, 0) = e1[k]; M(k, 1) = e2[k]; M(k, 2) = n[k]; }
  {
    const float d = det3(M), idet = 1.0f / d;
    Mi(0, 0) = (M(1, 1) * M(2, 2) - M(1, 2) * M(2, 1)) * idet; Mi(0, 1) = (M(0, 2) * M(2, 1) - M(0, 1) * M(2, 2)) * idet; Mi(0, 2) = (M(0, 1) * M(1, 2) - M(0, 2) * M(1, 1)) * idet;
    Mi(1, 0) = (M(1, 2) * M(2, 0) - M(1, 0) * M(2, 2)) * idet; Mi(1, 1) = (M(0, 0) * M(2, 2) - M(0, 2) * M(2, 0)) * idet; Mi(1, 2) = (M(0, 2) * M(1, 0) - M(0, 0) * M(1, 2)) * idet;
    Mi(2, 0) = (M(1, 0) * M(2, 1) - M(1, 1) * M(2, 0)) * idet; Mi(2, 1) = (M(0, 1) * M(2, 0) - M(0, 0) * M(2, 1)) * idet; Mi(2, 2) = (M(0, 0) * M(1, 1) - M(0, 1) * M(1, 0)) * idet;
  }
  for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) for (int c = 0; c < 3; c++) {
    const int i = bx + a, j = by + b, k = bz + c;
    const float d0 = (float)i * P.dx - v0.x, d1 = (float)j * P.dx - v0.y, d2 = (float)k * P.dx - v0.z;
    const float c0 = Mi(0, 0) * d0 + Mi(0, 1) * d1 + Mi(0, 2) * d2, c1 = Mi(1, 0) * d0 + Mi(1, 1) * d1 + Mi(1, 2) * d2, c2 = Mi(2, 0) * d0 + Mi(2, 1) * d1 + Mi(2, 2) * d2;
    if (!(0.f <= c0 && 0.f <= c1 && c0 + c1 <= 1.f)) continue;                          // src/rigid_transfer.cpp:50-53
    const float dist = fabsf(c2) * P.inv_dx;                                              // 43, 60
    const size_t node = ((size_t)i * P.nnode[1] + j) * P.nnode[2] + k;
    // 65-69 under the node's lock: the nearest body wins (a tie goes to the smaller id here, to whoever came first there)
    atomicMin(&R.node_key[node], ((unsigned long long)__float_as_uint(dist) << 32) | (unsigned long long)(id + 1));
    atomicOr(&R.node_tags[node], (uint32_t)(2 + (c2 < 0.f ? 1 : 0)) << (id * 2));     // 73-74
  }
}

__global__ void k_rigid_tile_flags(View V, Params P, RigidView R) {
  const int slot = blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= V.cnt->n_tiles) return;
  const TileMeta tm = V.meta[slot];
  MPMB_TILE_XYZ(P, tm, tx, ty, tz);
  R.tile_flag[slot] = R.page[((size_t)tx * R.nb[1] + ty) * R.nb[2] + (tz >> 1)];   // tile (4x4x4 nodes) -> page (4x4x8 nodes)
}

__device__ __forceinline__ float det4(const float *m) {  // m[c*4+r]
  float det = 0.f;
#pragma unroll
  for (int c = 0; c < 4; c++) {
    int cc[3], k = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) if (j != c) cc[k++] = j;
#define A4(r, c_) m[(c_) * 4 + (r)]
    const float minor = A4(1, cc[0]) * (A4(2, cc[1]) * A4(3, cc[2]) - A4(2, cc[2]) * A4(3, cc[1])) - A4(1, cc[1]) * (A4(2, cc[0]) * A4(3, cc[2]) - A4(2, cc[2]) * A4(3, cc[0])) +
                        A4(1, cc[2]) * (A4(2, cc[0]) * A4(3, cc[1]) - A4(2, cc[1]) * A4(3, cc[0]));
    det += ((c & 1) ? -1.f : 1.f) * A4(0, c) * minor;
#undef A4
  }
  return det;
}

// gather_cdf (src/rigid_transfer.cpp:120-274): one thread per storage row that holds a live particle
__global__ void __launch_bounds__(128) k_gather_cdf(View V, Params P, RigidView R) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= V.cnt->n_store) return;
  if (V.keys[row] >= (uint32_t)P.ntiles_total) return;   // dead or migrating
  const float4 q0 = V.q[0][row];
  const float X = __fmul_rn(q0.x, P.inv_dx), Y = __fmul_rn(q0.y, P.inv_dx), Z = __fmul_rn(q0.z, P.inv_dx);
  {  // 142-146: the page of the particle's CELL.  Outside: boundary_distance = 0, normal = 0, near_boundary_ = false (138-140)
     // and the colours untouched — expressed by NOT renewing the particle's mark
    const int cx = (int)X >> 2, cy = (int)Y >> 2, cz = (int)Z >> 3;
    if (cx < 0 || cy < 0 || cz < 0 || cx >= R.nb[0] || cy >= R.nb[1] || cz >= R.nb[2] || !R.page[((size_t)cx * R.nb[1] + cy) * R.nb[2] + cz]) return;
  }
  const uint32_t id = (__float_as_uint(V.q[6][row].w) & TAG_ID_MASK) - R.id_base;
  if (id >= (uint32_t)R.id_cap) { atomicOr(&V.cnt->error, DEVERR_BAD_INPUT); return; }
  R.p_cdf[id] = make_float4(0.f, 0.f, 0.f, 0.f);
  R.p_mark[id] = R.epoch << 1;
  int bx, by, bz;
  float rx, ry, rz;
  base_rel(q0.x, P.inv_dx, bx, rx);
  base_rel(q0.y, P.inv_dx, by, ry);
  base_rel(q0.z, P.inv_dx, bz, rz);
  float wx[3], wy[3], wz[3];
  bspline_weights(rx, wx);
  bspline_weights(ry, wy);
  bspline_weights(rz, wz);
  uint32_t pst = R.p_states[id];
  uint32_t words[27];
  uint32_t all_b = 0;
  for (int n = 0; n < 27; n++) {
    const size_t node = ((size_t)(bx + n / 9) * P.nnode[1] + (by + (n / 3) % 3)) * P.nnode[2] + (bz + n % 3);
    words[n] = cdf_word(R, node);
    all_b |= words[n] & CDF_TAG_MASK & CDF_STATE_MASK;                                   // 155-159
  }
  pst &= (all_b + (all_b >> 1));                                                         // 162
  uint32_t to_add = all_b & ~pst;                                                        // 164
  while (to_add) {
    const uint32_t bit = to_add & (0u - to_add);
    to_add ^= bit;
    float wd0 = 0.f, wd1 = 0.f;
    for (int n = 0; n < 27; n++) {
      if ((words[n] >> 24) == 0u) continue;                                              // 182-184
      const size_t node = ((size_t)(bx + n / 9) * P.nnode[1] + (by + (n / 3) % 3)) * P.nnode[2] + (bz + n % 3);
      const float d = cdf_dist_world(R, node, P.dx) * P.inv_dx;
      const float w = (wx[n / 9] * wy[(n / 3) % 3]) * wz[n % 3];
      if (words[n] & CDF_TAG_MASK & bit) { if (words[n] & (bit >> 1)) wd1 += d * w; else wd0 += d * w; }   // 195-198
    }
    if (wd0 + wd1 > 1e-7f) pst |= bit | ((bit >> 1) * (uint32_t)(wd0 < wd1));           // 200-205
  }
  R.p_states[id] = pst;
  if (pst == 0u) return;
  float XtX[16], XtY[4];
  for (int k = 0; k < 16; k++) XtX[k] = 0.f;
  for (int k = 0; k < 4; k++) XtY[k] = 0.f;
  for (int n = 0; n < 27; n++) {
    if ((words[n] >> 24) == 0u) continue;                                                // 217-219
    const uint32_t gs = words[n] & CDF_TAG_MASK;
    if (gs == 0u) continue;
    const uint32_t mask = (gs & pst & CDF_STATE_MASK) >> 1;
    float sgn;
    if ((gs & mask) == (pst & mask)) sgn = 1.f;                                          // 232-236: same colour
    else {
      const uint32_t diff = (gs & mask) ^ (pst & mask);                                  // 239-243: exactly one colour differs
      if (diff > 0u && (diff & (diff - 1u)) == 0u) sgn = -1.f; else continue;
    }
    const int a = n / 9, b = (n / 3) % 3, c = n % 3;
    const size_t node = ((size_t)(bx + a) * P.nnode[1] + (by + b)) * P.nnode[2] + (bz + c);
    const float d = cdf_dist_world(R, node, P.dx) * P.inv_dx;
    const float w = (wx[a] * wy[b]) * wz[c];
    const float xp[4] = {-(rx - (float)a), -(ry - (float)b), -(rz - (float)c), 1.f};   // (-dpos, 1)
    for (int i = 0; i < 4; i++) {
      for (int j = 0; j < 4; j++) XtX[j * 4 + i] += xp[i] * xp[j] * w;
      XtY[i] += sgn * (d * xp[i]) * w;                                                   // (-d dpos, d) resp. its negative
    }
  }
  if (fabsf(det4(XtX)) > 1e-4f) {                                                        // 251: mpm_reconstruction_guard<3>
    // inversed(XtX) * XtY by Gaussian elimination with partial pivoting
    float A[4][5];
    for (int r = 0; r < 4; r++) { for (int c = 0; c < 4; c++) A[r][c] = XtX[c * 4 + r]; A[r][4] = XtY[r]; }
    for (int c = 0; c < 4; c++) {
      int pv = c;
      for (int r = c + 1; r < 4; r++) if (fabsf(A[r][c]) > fabsf(A[pv][c])) pv = r;
      for (int k = 0; k < 5; k++) { const float t = A[c][k]; A[c][k] = A[pv][k]; A[pv][k] = t; }
      for (int r = c + 1; r < 4; r++) { const float f = A[r][c] / A[c][c]; for (int k = c; k < 5; k++) A[r][k] -= f * A[c][k]; }
    }
    float sol[4];
    for (int r = 3; r >= 0; r--) { float sum = A[r][4]; for (int c = r + 1; c < 4; c++) sum -= A[r][c] * sol[c]; sol[r] = sum / A[r][r]; }
    R.p_mark[id] = (R.epoch << 1) | 1u;
    const float l2 = sol[0] * sol[0] + sol[1] * sol[1] + sol[2] * sol[2];
    float4 o = make_float4(0.f, 0.f, 0.f, sol[3] * P.dx);
    if (l2 > 1e-4f) { const float il = 1.0f / sqrtf(l2); o.x = sol[0] * il; o.y = sol[1] * il; o.z = sol[2] * il; }
    R.p_cdf[id] = o;
  }
}

// Node colour words of a tile's 6x6x6 stencil block into shared memory
__device__ __forceinline__ void load_tile_words(const RigidView &R, const Params &P, int tx, int ty, int tz, uint32_t *s_word) {
  for (int n = threadIdx.x; n < ARENA; n += blockDim.x) {
    const int i = tx * 4 + n / 36, j = ty * 4 + (n / 6) % 6, k = tz * 4 + n % 6;
    s_word[n] = (i < P.nnode[0] && j < P.nnode[1] && k < P.nnode[2]) ? cdf_word(R, ((size_t)i * P.nnode[1] + j) * P.nnode[2] + k) : 0u;
  }
}

// block_op_rigid of rasterize_optimized (src/transfer.cpp:367-463) for the tiles of rigid pages: rewrites the tile's arena
// (k_p2g ran on it first and provided the output rows).  One CTA per tile, one thread per row, shared float atomics.
__global__ void __launch_bounds__(128) k_p2g_rigid(View V, Params P, RigidView R) {
  __shared__ float s_arena[4][ARENA];
  __shared__ uint32_t s_word[ARENA];
  __shared__ float s_acc[RIGID_MAX][6];
  const int tid = threadIdx.x;
  const int n_tiles = V.cnt->n_tiles;
    for (int slot = blockIdx.x; slot < n_tiles; slot += gridDim.x) {
    if (!R.tile_flag[slot]) continue;
    const TileMeta tm = V.meta[slot];
    MPMB_TILE_XYZ(P, tm, tx, ty, tz);
    for (int n = tid; n < ARENA; n += blockDim.x) { s_arena[0][n] = 0.f; s_arena[1][n] = 0.f; s_arena[2][n] = 0.f; s_arena[3][n] = 0.f; }
    for (int n = tid; n < RIGID_MAX * 6; n += blockDim.x) (&s_acc[0][0])[n] = 0.f;
    load_tile_words(R, P, tx, ty, tz, s_word);
    __syncthreads();
    const int nrow = tm.run_len + tm.arr_len;
    // A tile's run is cell-sorted, so neighbouring rows scatter to the same nodes: lanes take rows a large odd stride apart
    // (a bijection of [0, nrow)), which cuts the retries of the shared float atomics (CAS loops) — measured at config 3's size
    // with a 13 740-sample paddle: k_p2g_rigid 1.73 ms with rows in order (profiles/r02f_rigid_cost.json)
    const int stride = (nrow % 61) ? 61 : ((nrow % 67) ? 67 : 1);
    ImpulseAcc imp;
    imp.init();
    for (int idx = tid; idx < nrow; idx += blockDim.x) {
      const int g = (int)(((long long)idx * stride) % nrow);
      const uint32_t row = g < tm.run_len ? (uint32_t)(tm.run_begin + g) : V.arrivals_sorted[tm.arr_off + (g - tm.run_len)];
      const float4 a0 = V.q[0][row];
      if (g < tm.run_len && !(a0.w > 0.f)) continue;   // hole
      const float4 a1 = V.q[1][row], a2 = V.q[2][row], a3 = V.q[3][row];
      const float4 b0 = V.q[7][row], b1 = V.q[8][row], b2 = V.q[9][row];
      const float mass = fabsf(a0.w);
      const uint32_t id = (__float_as_uint(V.q[6][row].w) & TAG_ID_MASK) - R.id_base;
      const uint32_t pst = id < (uint32_t)R.id_cap ? R.p_states[id] : 0u;
      const bool fresh = id < (uint32_t)R.id_cap && (R.p_mark[id] >> 1) == R.epoch;   // gather_cdf saw this particle in this substep
      const float4 pc = fresh ? R.p_cdf[id] : make_float4(0.f, 0.f, 0.f, 0.f);
      float3 v = make_float3(a1.x, a1.y, a1.z);
      if (P.particle_gravity) { v.x += P.gdt[0]; v.y += P.gdt[1]; v.z += P.gdt[2]; }
      int bx, by, bz;
      float rx, ry, rz;
      base_rel(a0.x, P.inv_dx, bx, rx);
      base_rel(a0.y, P.inv_dx, by, ry);
      base_rel(a0.z, P.inv_dx, bz, rz);
      const int lx = bx - tx * 4, ly = by - ty * 4, lz = bz - tz * 4;
      if ((unsigned)lx >= 4u || (unsigned)ly >= 4u || (unsigned)lz >= 4u) continue;
      float wx[3], wy[3], wz[3], gx[3], gy[3], gz[3];
      bspline_weights(rx, wx); bspline_weights(ry, wy); bspline_weights(rz, wz);
      // dw of MPMKernel<3,2> (src/kernel.h:133-134), in world units (shuffle(): * inv_delta_x)
      { const float f = rx - 0.5f; gx[0] = (f - 1.0f) * P.inv_dx; gx[1] = (-2.0f * f + 1.0f) * P.inv_dx; gx[2] = f * P.inv_dx; }
      { const float f = ry - 0.5f; gy[0] = (f - 1.0f) * P.inv_dx; gy[1] = (-2.0f * f + 1.0f) * P.inv_dx; gy[2] = f * P.inv_dx; }
      { const float f = rz - 0.5f; gz[0] = (f - 1.0f) * P.inv_dx; gz[1] = (-2.0f * f + 1.0f) * P.inv_dx; gz[2] = f * P.inv_dx; }
      const float A[9] = {a1.w, a2.x, a2.y, a2.z, a2.w, a3.x, a3.y, a3.z, a3.w};
      const float bb[9] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w, b2.x};
      // dt * calculate_force() back out of the cached affine matrix: A = force * S + apic_b * 4 m, S = -4 dt / dx
      float tf[9];
      for (int k = 0; k < 9; k++) tf[k] = (A[k] - bb[k] * (4.0f * mass)) * (-0.25f * P.dx);
      for (int n = 0; n < 27; n++) {
        const int a = n / 9, b = (n / 3) % 3, c = n % 3;
        const int ln = ((lx + a) * 6 + (ly + b)) * 6 + (lz + c);
        const float w = (wx[a] * wy[b]) * wz[c];
        const float dpx = rx - (float)a, dpy = ry - (float)b, dpz = rz - (float)c;
        const uint32_t word = s_word[ln];
        if (!cdf_compatible(word, pst)) {   // different colour: project against the body instead of writing to the grid (420-446)
          const int rid = (int)(word >> 24) - 1;
          if (rid < 0 || rid >= R.n_bodies) continue;
          const RigidDev &B = R.bodies[rid];
          const float3 gp = make_float3(P.dx * (float)(bx + a), P.dx * (float)(by + b), P.dx * (float)(bz + c));
          const float3 rv = rigid_velocity_at(B, gp);
          const float3 pr = friction_project_rel(v, rv, make_float3(pc.x, pc.y, pc.z), B.fric[(pst >> (2 * rid)) & 1u]);
          const float dwx = (gx[a] * wy[b]) * wz[c], dwy = (wx[a] * gy[b]) * wz[c], dwz = (wx[a] * wy[b]) * gz[c];
          const float3 j = make_float3(mass * w * (v.x - pr.x) + (tf[0] * dwx + tf[3] * dwy + tf[6] * dwz), mass * w * (v.y - pr.y) + (tf[1] * dwx + tf[4] * dwy + tf[7] * dwz),
                                       mass * w * (v.z - pr.z) + (tf[2] * dwx + tf[5] * dwy + tf[8] * dwz));
          imp.add(s_acc, rid, B, j, gp);
          continue;
        }
        atomicAdd(&s_arena[0][ln], w * (mass * v.x + (A[0] * dpx + A[3] * dpy + A[6] * dpz)));   // 451-459
        atomicAdd(&s_arena[1][ln], w * (mass * v.y + (A[1] * dpx + A[4] * dpy + A[7] * dpz)));
        atomicAdd(&s_arena[2][ln], w * (mass * v.z + (A[2] * dpx + A[5] * dpy + A[8] * dpz)));
        atomicAdd(&s_arena[3][ln], w * mass);
      }
    }
    imp.flush(s_acc);
    __syncthreads();
    float4 *out = V.arena + (size_t)slot * ARENA;
    for (int n = tid; n < ARENA; n += blockDim.x) out[n] = make_float4(s_arena[0][n], s_arena[1][n], s_arena[2][n], s_arena[3][n]);
    for (int n = tid; n < RIGID_MAX * 6; n += blockDim.x) {
      const float a = (&s_acc[0][0])[n];
      if (a != 0.f) atomicAdd(&R.bodies[n / 6].acc[n % 6], a);
    }
    __syncthreads();
  }
}

// apply_tmp_velocity (src/transfer.cpp:578-580, 967-969): velocity += inv_mass * sum j, angular_velocity += Iw^-1 * sum (p - c) x j
__global__ void k_rigid_apply(RigidView R) {
  const int b = threadIdx.x;
  if (b >= R.n_bodies) return;
  RigidDev &B = R.bodies[b];
  for (int k = 0; k < 3; k++) {
    B.vel[k] += B.inv_mass * B.acc[k];
    B.ang[k] += B.inv_inertia[k] * B.acc[3] + B.inv_inertia[3 + k] * B.acc[4] + B.inv_inertia[6 + k] * B.acc[5];
  }
  for (int k = 0; k < 6; k++) B.acc[k] = 0.f;
}

// block_op_rigid of resample_optimized (src/transfer.cpp:706-835) + what k_g2p does for the ordering (output rows, keys,
// movers, stay counts), for the tiles of rigid pages.  Launched BEFORE k_g2p, which commits the substep.
__global__ void __launch_bounds__(128) k_g2p_rigid(View V, Params P, RigidView R, const float4 *vel) {
  __shared__ float4 s_vel[ARENA];
  __shared__ uint32_t s_word[ARENA];
  __shared__ float s_acc[RIGID_MAX][6];
  __shared__ int s_stay;
  const int tid = threadIdx.x;
  const int n_tiles = V.cnt->n_tiles;
  const float scale = -4.0f * P.inv_dx * P.dt;
  for (int slot = blockIdx.x; slot < n_tiles; slot += gridDim.x) {
    if (!R.tile_flag[slot]) continue;
    const TileMeta tm = V.meta[slot];
    const int tile = tm.tile;
    MPMB_TILE_XYZ(P, tm, tx, ty, tz);
    for (int n = tid; n < ARENA; n += blockDim.x) s_vel[n] = vel[(size_t)slot * ARENA + n];
    for (int n = tid; n < RIGID_MAX * 6; n += blockDim.x) (&s_acc[0][0])[n] = 0.f;
    load_tile_words(R, P, tx, ty, tz, s_word);
    if (tid == 0) s_stay = 0;
    __syncthreads();
    const int nrow = tm.run_len + tm.arr_len;
    int my_stay = 0;
    ImpulseAcc imp;
    imp.init();
    for (int g = tid; g < nrow; g += blockDim.x) {
      const uint32_t row = g < tm.run_len ? (uint32_t)(tm.run_begin + g) : V.arrivals_sorted[tm.arr_off + (g - tm.run_len)];
      const float4 q0 = V.q[0][row];
      if (g < tm.run_len && !(q0.w > 0.f)) continue;   // hole
      const float mass = fabsf(q0.w);
      const float4 q1 = V.q[1][row], q4 = V.q[4][row], q5 = V.q[5][row], q6 = V.q[6][row];
      const size_t o = V.outpos[row];
      if (o >= (size_t)V.cap_particles) { atomicOr(&V.cnt->error, DEVERR_PARTICLE_CAPACITY); continue; }
      const uint32_t tag = __float_as_uint(q6.w);
      const uint32_t id = (tag & TAG_ID_MASK) - R.id_base;
      const bool known = id < (uint32_t)R.id_cap;
      const uint32_t pst = known ? R.p_states[id] : 0u;
      const uint32_t mark = known ? R.p_mark[id] : 0u;
      const bool fresh = known && (mark >> 1) == R.epoch;   // gather_cdf saw this particle in this substep
      const float4 pc = fresh ? R.p_cdf[id] : make_float4(0.f, 0.f, 0.f, 0.f);
      const bool near = fresh && (mark & 1u);
      const float3 bn = make_float3(pc.x, pc.y, pc.z);
      // p.get_velocity() as rasterize left it: the gravity kick is stored back there (src/transfer.cpp:383-385)
      float3 pv = make_float3(q1.x, q1.y, q1.z);
      if (P.particle_gravity) { pv.x += P.gdt[0]; pv.y += P.gdt[1]; pv.z += P.gdt[2]; }
      int bx, by, bz;
      float rx, ry, rz;
      base_rel(q0.x, P.inv_dx, bx, rx);
      base_rel(q0.y, P.inv_dx, by, ry);
      base_rel(q0.z, P.inv_dx, bz, rz);
      const int lx = bx - tx * 4, ly = by - ty * 4, lz = bz - tz * 4;
      if ((unsigned)lx >= 4u || (unsigned)ly >= 4u || (unsigned)lz >= 4u) continue;
      float wx[3], wy[3], wz[3];
      bspline_weights(rx, wx); bspline_weights(ry, wy); bspline_weights(rz, wz);
      float3 v = make_float3(0.f, 0.f, 0.f);
      Mat3 B;
      for (int k = 0; k < 9; k++) B.m[k] = 0.f;
      int rigid_id = -1;
      for (int n = 0; n < 27; n++) {
        const int a = n / 9, b = (n / 3) % 3, c = n % 3;
        const int ln = ((lx + a) * 6 + (ly + b)) * 6 + (lz + c);
        const float w = (wx[a] * wy[b]) * wz[c];
        const float dp[3] = {rx - (float)a, ry - (float)b, rz - (float)c};
        float3 gv = make_float3(s_vel[ln].x, s_vel[ln].y, s_vel[ln].z);
        const uint32_t word = s_word[ln];
        if (!cdf_compatible(word, pst)) {   // different colour (761-785)
          float3 fake = pv, vg = make_float3(0.f, 0.f, 0.f);
          float friction = 0.f;
          const int rid = (int)(word >> 24) - 1;
          if (rid >= 0 && rid < R.n_bodies) {
            const RigidDev &Bd = R.bodies[rid];
            vg = rigid_velocity_at(Bd, make_float3((float)(bx + a) * P.dx, (float)(by + b) * P.dx, (float)(bz + c) * P.dx));
            rigid_id = rid;
            friction = Bd.fric[(pst >> (2 * rid)) & 1u];
          }
          if (near) {
            fake = friction_project_rel(pv, vg, bn, friction);
            const float push = P.dt * P.dx * R.pushing_force;
            fake.x += bn.x * push; fake.y += bn.y * push; fake.z += bn.z * push;
          }
          gv = fake;
        }
        v.x = fmaf(gv.x, w, v.x); v.y = fmaf(gv.y, w, v.y); v.z = fmaf(gv.z, w, v.z);                  // 788
        const float wgx = w * gv.x, wgy = w * gv.y, wgz = w * gv.z;
        for (int cc = 0; cc < 3; cc++) { B(0, cc) = fmaf(wgx, dp[cc], B(0, cc)); B(1, cc) = fmaf(wgy, dp[cc], B(1, cc)); B(2, cc) = fmaf(wgz, dp[cc], B(2, cc)); }   // 793-795
      }
      Mat3 cdg;  // 810-815
      for (int k = 0; k < 9; k++) cdg.m[k] = fmaf(scale, B.m[k], (k % 4 == 0) ? 1.f : 0.f);
      Mat3 Bst = B;  // p.apic_b: zero next to a boundary (800-804)
      if (near) for (int k = 0; k < 9; k++) Bst.m[k] = 0.f;
      Mat3 F;
      F.m[0] = q4.x; F.m[1] = q4.y; F.m[2] = q4.z; F.m[3] = q4.w; F.m[4] = q5.x; F.m[5] = q5.y; F.m[6] = q5.z; F.m[7] = q5.w; F.m[8] = q6.x;
      float ps = q6.y;
      const float vol = q6.z;
      Mat3 force, A;
      material_step<true>(P.mats[tag >> TAG_ID_BITS], cdg, F, ps, vol, force);
      make_affine(force, Bst, mass, scale, A);
      float3 x = make_float3(fmaf(v.x, P.dt, q0.x), fmaf(v.y, P.dt, q0.y), fmaf(v.z, P.dt, q0.z));   // 819
      if (near && pc.w < -0.05f * P.dx && pc.w > -P.dx * 0.3f) {   // 823-832: position correction
        const float3 dv = make_float3(pc.w * bn.x * R.penalty, pc.w * bn.y * R.penalty, pc.w * bn.z * R.penalty);
        v.x -= dv.x; v.y -= dv.y; v.z -= dv.z;
        if (rigid_id != -1) imp.add(s_acc, rigid_id, R.bodies[rigid_id], make_float3(dv.x * mass, dv.y * mass, dv.z * mass), x);
      }
      uint32_t key = make_key(P, x.x, x.y, x.z);
      if (P.clean_boundary && reference_deletes(P, x, v)) key = (uint32_t)(P.ntiles_total + SPECIAL_DEAD);
      if (!isfinite(x.x + x.y + x.z)) key = (uint32_t)(P.ntiles_total + SPECIAL_DEAD);
      store_particle<true>(V.qn, o, x, key == (uint32_t)tile ? mass : -mass, v, A, F, ps, vol, tag, Bst);
      V.keys_next[o] = key;
      if (key == (uint32_t)tile) {
        my_stay++;
      } else if (key != (uint32_t)(P.ntiles_total + SPECIAL_DEAD)) {
        const int m = atomicAdd(&V.cnt->n_movers_next, 1);
        if (m >= V.cap_particles) { atomicOr(&V.cnt->error, DEVERR_PARTICLE_CAPACITY); continue; }
        V.mover_dst_n[m] = key;
        V.mover_idx_n[m] = (uint32_t)o;
        MPMB_COUNT_ARRIVAL(V, P, key);
      }
    }
    if (my_stay) atomicAdd(&s_stay, my_stay);
    imp.flush(s_acc);
    __syncthreads();
    if (tid == 0) V.stay_next[tile] = s_stay;
    for (int n = tid; n < RIGID_MAX * 6; n += blockDim.x) {
      const float a = (&s_acc[0][0])[n];
      if (a != 0.f) atomicAdd(&R.bodies[n / 6].acc[n % 6], a);
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------ debug grid
__global__ void k_dense_grid(View V, Params P, int which, float4 *dense) {
  size_t n = (size_t)P.nnode[0] * P.nnode[1] * P.nnode[2];
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    int gz = (int)(i % P.nnode[2]), gy = (int)((i / P.nnode[2]) % P.nnode[1]), gx = (int)(i / ((size_t)P.nnode[2] * P.nnode[1]));
    int tx = gx >> 2, ty = gy >> 2, tz = gz >> 2;
    float4 g = gather_node(V.arena, V.cap_tiles, gx & 3, gy & 3, gz & 3, [&](int ox, int oy, int oz) {
      int x = tx - ox, y = ty - oy, z = tz - oz - P.tz_off;  // z: local layer
      if (x < 0 || y < 0 || z < 0 || x >= P.nt[0] || y >= P.nt[1] || z >= P.nt[2]) return -1;
      return V.slot_map[(x * P.nt[1] + y) * P.nt[2] + z];
    });
    if (which == 1) g = node_update(P, V.sdf4, g, gx, gy, gz);
    dense[i] = g;
  }
}

__global__ void k_planes_to_sdf(Params P, int n_planes, const float4 *planes, float4 *sdf4) {
  size_t n = (size_t)P.nnode[0] * P.nnode[1] * P.nnode[2];
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    int gz = (int)(i % P.nnode[2]), gy = (int)((i / P.nnode[2]) % P.nnode[1]), gx = (int)(i / ((size_t)P.nnode[2] * P.nnode[1]));
    float best = 1e30f;
    float4 out = make_float4(1.f, 0.f, 0.f, 1e30f);
    for (int k = 0; k < n_planes; k++) {
      float4 pl = planes[k];
      float phi = pl.x * gx + pl.y * gy + pl.z * gz + pl.w;
      if (phi < best) { best = phi; out = make_float4(pl.x, pl.y, pl.z, phi); }
    }
    sdf4[i] = out;
  }
}

// Level set rasterised on the device from analytic solids (the scripts build theirs with
// levelset.add_plane / add_cuboid / add_sphere, scripts/mls-cpic/sand_sweep.py:13-19, scripts/async/sand.py:35,
// scripts/mls-cpic/sand_stir.py:9): phi = min over the shapes of their signed distance (grid units, negative inside the
// obstacle), n = the unit gradient of the minimiser.  inside_out turns a solid into a container.
//   plane   p = (n_x, n_y, n_z, d):      phi = n.X + d
//   sphere  p = (c_x, c_y, c_z, r):      phi = |X - c| - r
//   cuboid  p = (lo_xyz, hi_xyz):        phi = exact box distance (outside: to the nearest face/edge/corner; inside: -depth)
__device__ __forceinline__ float4 shape_sdf(const MpmbShape &S, float x, float y, float z) {
  float4 o;  // (n, phi)
  if (S.kind == MPMB_SHAPE_PLANE) {
    o = make_float4(S.p[0], S.p[1], S.p[2], S.p[0] * x + S.p[1] * y + S.p[2] * z + S.p[3]);
  } else if (S.kind == MPMB_SHAPE_SPHERE) {
    const float dx = x - S.p[0], dy = y - S.p[1], dz = z - S.p[2];
    const float r = sqrtf(dx * dx + dy * dy + dz * dz), inv = r > 1e-20f ? 1.0f / r : 0.f;
    o = make_float4(dx * inv, dy * inv, dz * inv, r - S.p[3]);
    if (!(r > 1e-20f)) o.x = 1.f;
  } else {
    const float X[3] = {x, y, z};
    float q[3], out2 = 0.f, depth = 1e30f;
    int amax = 0;
    float sgn[3];
    for (int a = 0; a < 3; a++) {
      const float c = 0.5f * (S.p[a] + S.p[3 + a]), hw = 0.5f * (S.p[3 + a] - S.p[a]);
      const float d = X[a] - c;
      sgn[a] = d < 0.f ? -1.f : 1.f;
      q[a] = fabsf(d) - hw;                  // > 0 outside along this axis
      if (q[a] > 0.f) out2 += q[a] * q[a];
      if (-q[a] < depth) { depth = -q[a]; amax = a; }
    }
    if (out2 > 0.f) {
      const float r = sqrtf(out2), inv = 1.0f / r;
      o = make_float4(q[0] > 0.f ? sgn[0] * q[0] * inv : 0.f, q[1] > 0.f ? sgn[1] * q[1] * inv : 0.f, q[2] > 0.f ? sgn[2] * q[2] * inv : 0.f, r);
    } else {
      o = make_float4(amax == 0 ? sgn[0] : 0.f, amax == 1 ? sgn[1] : 0.f, amax == 2 ? sgn[2] : 0.f, -depth);
    }
  }
  if (S.inside_out) o = make_float4(-o.x, -o.y, -o.z, -o.w);
  return o;
}
__global__ void k_shapes_to_sdf(Params P, int n_shapes, const MpmbShape *shapes, float4 *sdf4) {
  size_t n = (size_t)P.nnode[0] * P.nnode[1] * P.nnode[2];
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    int gz = (int)(i % P.nnode[2]), gy = (int)((i / P.nnode[2]) % P.nnode[1]), gx = (int)(i / ((size_t)P.nnode[2] * P.nnode[1]));
    float4 best = make_float4(1.f, 0.f, 0.f, 1e30f);
    for (int k = 0; k < n_shapes; k++) {
      const float4 c = shape_sdf(shapes[k], (float)gx, (float)gy, (float)gz);
      if (c.w < best.w) best = c;
    }
    sdf4[i] = best;
  }
}

// ------------------------------------------------------------------------------ z-slab exchange
// Halo message: int4 header {count,0,0,0} | int tile_xy[cap_xy] | float4 arena[cap_xy][216].
// Packs the arenas of the owned tiles of one tile layer (the partial sums of (p,m) the neighbour
// rank's nodes need); the neighbour registers them as ghost tiles, so its grid update sums them in
// the same fixed order as a single-GPU run would.
// layer_z: LOCAL tile layer (global layer - P.tz_off)
__global__ void k_halo_pack(View V, Params P, int layer_z, int cap_xy, int *hdr, int *tile_xy, float4 *arenas) {
  __shared__ int s_idx;
  const int n_tiles = V.cnt->n_tiles;
  for (int slot = blockIdx.x; slot < n_tiles; slot += gridDim.x) {
    const int tile = V.meta[slot].tile;
    if (tile % P.nt[2] != layer_z) continue;  // uniform per CTA
    if (threadIdx.x == 0) {
      int idx = atomicAdd(&hdr[0], 1);
      if (idx >= cap_xy) { atomicOr(&V.cnt->error, DEVERR_TILE_CAPACITY); idx = -1; }
      else tile_xy[idx] = tile / P.nt[2];
      s_idx = idx;
    }
    __syncthreads();
    const int idx = s_idx;
    if (idx >= 0)
      for (int n = threadIdx.x; n < ARENA; n += blockDim.x) arenas[(size_t)idx * ARENA + n] = V.arena[(size_t)slot * ARENA + n];
    __syncthreads();
  }
}

__global__ void k_halo_unpack(View V, Params P, int layer_z, int cap_xy, const int *hdr, const int *tile_xy, const float4 *arenas) {
  __shared__ int s_slot;
  const int count = min(hdr[0], cap_xy);
  for (int e = blockIdx.x; e < count; e += gridDim.x) {
    if (threadIdx.x == 0) {
      int slot = V.cnt->n_tiles + atomicAdd(&V.cnt->n_ghost, 1);
      if (slot >= V.cap_tiles) { atomicOr(&V.cnt->error, DEVERR_TILE_CAPACITY); slot = -1; }
      else if (MPMB_CHK(V.cnt, (unsigned)tile_xy[e] < (unsigned)(P.nt[0] * P.nt[1]), 10, e, tile_xy[e], count)) V.slot_map[tile_xy[e] * P.nt[2] + layer_z] = slot;  // rewritten densely by the next ordering
      s_slot = slot;
    }
    __syncthreads();
    const int slot = s_slot;
    if (slot >= 0)
      for (int n = threadIdx.x; n < ARENA; n += blockDim.x) V.arena[(size_t)slot * ARENA + n] = arenas[(size_t)e * ARENA + n];
    __syncthreads();
  }
}

// Migration message: int4 header {count,0,0,0} | float4 record[cap][N_Q].  Emigrants are the movers
// whose destination is the face's special key; their row in the (new) current storage is a hole
// for everybody else, the key is set dead so that downloads skip it.
__global__ void k_migrate_pack(View V, uint32_t key_face, uint32_t key_dead, int cap, int *hdr, float4 *rec) {
  const int n = V.cnt->n_movers;
  for (int m = blockIdx.x * blockDim.x + threadIdx.x; m < n; m += gridDim.x * blockDim.x) {
    if (V.mover_dst[m] != key_face) continue;
    const uint32_t i = V.mover_idx[m];
    const int idx = atomicAdd(&hdr[0], 1);
    V.keys[i] = key_dead;
    V.mover_dst[m] = key_dead;
    if (idx >= cap) { atomicOr(&V.cnt->error, DEVERR_MIGRATE_CAPACITY); continue; }
#pragma unroll
    for (int k = 0; k < N_Q; k++) rec[(size_t)idx * N_Q + k] = V.q[k][i];
  }
}

// Immigrants are appended after the last row of the current storage and enter the mover list, so the
// next ordering bins them like any other particle that changed tile.
__global__ void k_migrate_unpack(View V, Params P, int cap, const int *hdr, const float4 *rec) {
  const int count = min(hdr[0], cap);
  const int base = V.cnt->n_store, mbase = V.cnt->n_movers;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < count; e += gridDim.x * blockDim.x) {
    const int dst = base + e;
    if (dst >= V.cap_particles) { atomicOr(&V.cnt->error, DEVERR_PARTICLE_CAPACITY); continue; }
    float4 r0 = rec[(size_t)e * N_Q];
#pragma unroll
    for (int k = 0; k < N_Q; k++) V.q[k][dst] = rec[(size_t)e * N_Q + k];
    const uint32_t key = make_key(P, r0.x, r0.y, r0.z);
    V.keys[dst] = key;
    V.mover_dst[mbase + e] = key;
    V.mover_idx[mbase + e] = (uint32_t)dst;
    MPMB_COUNT_ARRIVAL(V, P, key);
  }
}
// ---- peer-memory exchange (no NCCL on the data path): the sending kernels store the payload straight into the neighbour
// GPU's receive buffer over NVLink and release {count, seq} system-wide; the receiving kernels acquire seq before they
// read.  One substep = one seq value; a bounded spin turns a lost peer into an error flag instead of a hung GPU.
// What mpmb_substep launches on a z-slab rank is fused:  one kernel per direction handles BOTH faces,
// resets its own counters and publishes from its last CTA; the receiving kernel waits for the neighbours' sequence
// numbers itself.  4 launches per substep instead of 22 (round 1: memset + pack + publish and wait + unpack (+ commit)
// per face and per message kind, each a 2-3 us launch on a stream whose real work is ~100 us at 8 ranks).
struct XFace {     // one face of one message kind, as the kernels see it
  char *tx;        // the neighbour's receive buffer (peer memory) — header | payload
  const char *rx;  // my receive buffer for this face
  int layer;       // halo: LOCAL tile layer packed (send) / registered as ghosts (recv)
  uint32_t key;    // migration: the special key of this face
};
__device__ __forceinline__ void xchg_spin(const int *hdr, int seq, Counters *cnt) {
  const long long t0 = clock64();
  for (;;) {
    int v;
    asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(hdr + 1) : "memory");
    if (v >= seq) return;
    if (clock64() - t0 > 60000000000ll) {  // ~30 s: a lost peer becomes an error flag, not a hung GPU
      atomicOr(&cnt->error, DEVERR_PEER_TIMEOUT);
      return;
    }
    __nanosleep(200);
  }
}
// last CTA of a sending kernel: header {count, seq} of every face, counters back to zero for the next substep
__device__ __forceinline__ void xchg_publish_last(int *xcount, int *done, XFace f0, XFace f1, int mask, int seq) {
  __threadfence_system();  // my payload stores are visible system-wide before the CTA is counted as done
  __syncthreads();
  if (threadIdx.x == 0) {
    if (atomicAdd(done, 1) == (int)gridDim.x - 1) {
      __threadfence_system();
#pragma unroll
      for (int f = 0; f < 2; f++) {
        if (!((mask >> f) & 1)) continue;
        int *hdr = (int *)(f == 0 ? f0.tx : f1.tx);
        hdr[0] = atomicAdd(&xcount[4 * f], 0);
        __threadfence_system();
        asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(hdr + 1), "r"(seq) : "memory");
        xcount[4 * f] = 0;
      }
      *done = 0;
    }
  }
}

__global__ void __launch_bounds__(128) k_halo_send2(View V, Params P, int cap_xy, int idx_bytes, XFace f0, XFace f1, int mask, int *xcount, int *done) {
  __shared__ int s_idx[2];
  const int seq = V.cnt->xstep + 1;  // this substep's halo
  const int n_tiles = V.cnt->n_tiles;
  for (int slot = blockIdx.x; slot < n_tiles; slot += gridDim.x) {
    const int tile = V.meta[slot].tile, lz = tile % P.nt[2];
    const bool to0 = (mask & 1) && lz == f0.layer, to1 = (mask & 2) && lz == f1.layer;  // a one-layer slab sends the tile both ways
    if (!to0 && !to1) continue;                                                          // uniform per CTA
    if (threadIdx.x < 2) {
      int idx = -1;
      if (threadIdx.x == 0 ? to0 : to1) {
        idx = atomicAdd(&xcount[4 * threadIdx.x], 1);
        if (idx >= cap_xy) { atomicOr(&V.cnt->error, DEVERR_TILE_CAPACITY); idx = -1; }
        else ((int *)((threadIdx.x == 0 ? f0.tx : f1.tx) + 16))[idx] = tile / P.nt[2];
      }
      s_idx[threadIdx.x] = idx;
    }
    __syncthreads();
    const float4 *src = V.arena + (size_t)slot * ARENA;
#pragma unroll
    for (int f = 0; f < 2; f++) {
      const int idx = s_idx[f];
      if (idx < 0) continue;
      float4 *dst = (float4 *)((f == 0 ? f0.tx : f1.tx) + 16 + idx_bytes) + (size_t)idx * ARENA;
      for (int n = threadIdx.x; n < ARENA; n += blockDim.x) dst[n] = src[n];
    }
    __syncthreads();
  }
  xchg_publish_last(xcount, done, f0, f1, mask, seq);
}

__global__ void __launch_bounds__(128) k_halo_recv2(View V, Params P, int cap_xy, int idx_bytes, XFace f0, XFace f1, int mask) {
  __shared__ int s_slot;
  const int seq = V.cnt->xstep + 1;
  if (threadIdx.x == 0) {
    if (mask & 1) xchg_spin((const int *)f0.rx, seq, V.cnt);
    if (mask & 2) xchg_spin((const int *)f1.rx, seq, V.cnt);
  }
  __syncthreads();
  const int c0 = (mask & 1) ? min(((const int *)f0.rx)[0], cap_xy) : 0, c1 = (mask & 2) ? min(((const int *)f1.rx)[0], cap_xy) : 0;
  for (int e = blockIdx.x; e < c0 + c1; e += gridDim.x) {
    const bool lo = e < c0;
    const char *b = lo ? f0.rx : f1.rx;
    const int k = lo ? e : e - c0, layer = lo ? f0.layer : f1.layer;
    if (threadIdx.x == 0) {
      int slot = V.cnt->n_tiles + atomicAdd(&V.cnt->n_ghost, 1);
      if (slot >= V.cap_tiles) { atomicOr(&V.cnt->error, DEVERR_TILE_CAPACITY); slot = -1; }
      else if (MPMB_CHK(V.cnt, (unsigned)((const int *)(b + 16))[k] < (unsigned)(P.nt[0] * P.nt[1]), 11, k, ((const int *)(b + 16))[k], c0 + c1))
        V.slot_map[((const int *)(b + 16))[k] * P.nt[2] + layer] = slot;  // rewritten densely by the next ordering
      s_slot = slot;
    }
    __syncthreads();
    const int slot = s_slot;
    const float4 *src = (const float4 *)(b + 16 + idx_bytes) + (size_t)k * ARENA;
    if (slot >= 0)
      for (int n = threadIdx.x; n < ARENA; n += blockDim.x) V.arena[(size_t)slot * ARENA + n] = src[n];
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256) k_migrate_send2(View V, uint32_t key_dead, int cap, XFace f0, XFace f1, int mask, int *xcount, int *done) {
  const int n = V.cnt->n_movers;
  const int seq = V.cnt->xstep;  // k_g2p has committed this substep
  for (int m = blockIdx.x * blockDim.x + threadIdx.x; m < n; m += gridDim.x * blockDim.x) {
    const uint32_t d = V.mover_dst[m];
    const int f = ((mask & 1) && d == f0.key) ? 0 : (((mask & 2) && d == f1.key) ? 1 : -1);
    if (f < 0) continue;
    const uint32_t i = V.mover_idx[m];
    const int idx = atomicAdd(&xcount[4 * f], 1);
    V.keys[i] = key_dead;
    V.mover_dst[m] = key_dead;
    if (idx >= cap) { atomicOr(&V.cnt->error, DEVERR_MIGRATE_CAPACITY); continue; }
    float4 *rec = (float4 *)((f == 0 ? f0.tx : f1.tx) + 16);
#pragma unroll
    for (int k = 0; k < N_Q; k++) rec[(size_t)idx * N_Q + k] = V.q[k][i];
  }
  xchg_publish_last(xcount, done, f0, f1, mask, seq);
}

// one CTA: waits for both neighbours, appends the immigrants of both faces after the last storage row, enters them in the
// mover list and commits the counts
__global__ void __launch_bounds__(1024) k_migrate_recv2(View V, Params P, int cap, XFace f0, XFace f1, int mask) {
  __shared__ int s_n[2];
  const int seq = V.cnt->xstep;
  if (threadIdx.x == 0) {
    if (mask & 1) xchg_spin((const int *)f0.rx, seq, V.cnt);
    if (mask & 2) xchg_spin((const int *)f1.rx, seq, V.cnt);
    int c0 = (mask & 1) ? min(((const int *)f0.rx)[0], cap) : 0, c1 = (mask & 2) ? min(((const int *)f1.rx)[0], cap) : 0;
    if (c0 > cap || c1 > cap) atomicOr(&V.cnt->error, DEVERR_MIGRATE_CAPACITY);
    const int room = max(V.cap_particles - V.cnt->n_store, 0);
    if (c0 + c1 > room) { atomicOr(&V.cnt->error, DEVERR_PARTICLE_CAPACITY); c0 = min(c0, room); c1 = min(c1, room - c0); }
    s_n[0] = c0; s_n[1] = c1;
  }
  __syncthreads();
  const int c0 = s_n[0], c1 = s_n[1];
  const int base = V.cnt->n_store, mbase = V.cnt->n_movers;
  for (int e = threadIdx.x; e < c0 + c1; e += blockDim.x) {
    const float4 *rec = (const float4 *)((e < c0 ? f0.rx : f1.rx) + 16) + (size_t)(e < c0 ? e : e - c0) * N_Q;
    const int dst = base + e;
    const float4 r0 = rec[0];
#pragma unroll
    for (int k = 0; k < N_Q; k++) V.q[k][dst] = rec[k];
    const uint32_t key = make_key(P, r0.x, r0.y, r0.z);
    V.keys[dst] = key;
    V.mover_dst[mbase + e] = key;
    V.mover_idx[mbase + e] = (uint32_t)dst;
    MPMB_COUNT_ARRIVAL(V, P, key);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    V.cnt->n_store = base + c0 + c1;
    V.cnt->n_movers = mbase + c0 + c1;
  }
}

__global__ void k_migrate_commit(Counters *c, const int *hdr, int cap, int cap_particles) {
  const int n = min(min(hdr[0], cap), max(cap_particles - c->n_store, 0));
  c->n_store += n;
  c->n_movers += n;
}

}  // namespace mpmb

// =====================================================================================
// Host side
// =====================================================================================
using namespace mpmb;

struct MpmbEngine {
  MpmbConfig cfg{};
  Params P{};
  cudaStream_t stream = nullptr;
  std::string err;
  bool sticky_cuda = false;

  int64_t cap = 0;       // particle rows allocated per buffer
  int cur = 0;           // which q/keys buffer is current
  int ord = 0;           // which (run,stay) set is current
  int mov = 0;           // which mover list the next ordering consumes
  bool fresh = false;    // storage was just uploaded: arrival lists come from the radix sort
  float4 *q[2][N_Q] = {};
  uint32_t *keys[2] = {};       // tile of every row (cur / next)
  uint32_t *outpos = nullptr;
  uint32_t *mover_dst[2] = {}, *mover_idx[2] = {};
  uint32_t *arrivals = nullptr, *arrivals_sorted = nullptr;
  uint32_t *keys_sorted = nullptr, *iota = nullptr;  // upload-time radix sort scratch
  uint32_t special_min = 0, key_dead = 0;
  void *cub_temp = nullptr;
  size_t cub_bytes = 0;
  int key_bits = 32;

  int cap_tiles = 0;
  int ntot = 0;           // tiles of the dense tile grid
  int ord_blocks = 0;
  int *run_begin[2] = {}, *run_len[2] = {}, *stay[2] = {};  // [ord]: current / next
  int *arr_cnt = nullptr, *arr_off = nullptr, *arr_len = nullptr, *arr_cur = nullptr, *slot_map = nullptr, *blocksum = nullptr;
  TileMeta *meta = nullptr;
  float4 *arena = nullptr;
  float4 *vel = nullptr;  // node velocities per active tile (k_grid -> k_g2p)
  float4 *sdf4 = nullptr;
  Counters *cnt = nullptr;

  int stage = 0;  // 0 idle/after resample, 1 after sort, 2 after rasterize
  int64_t mig_cap = 0;
  // peer-memory exchange: rx = my receive buffers [kind 0 halo / 1 migration][face], tx = the
  // neighbour's receive buffer I write into through face f (mapped by IPC or same-process pointer)
  char *rx[2][2] = {}, *tx[2][2] = {};
  bool tx_ipc[2][2] = {};
  int *xcount = nullptr;  // local pack counters (int4 per [kind][face])
  int xstep = 0;          // substeps completed since the peers were connected
  bool skip_b = false;        // intermediate substeps of mpmb_substep do not store apic_b
  void *stage_buf = nullptr;  // cached staging buffer of the host<->device marshalling
  size_t stage_bytes = 0;
  // device image of the reference's AoS pool + index vector, kept from mpmb_upload_aos to mpmb_download_aos
  unsigned char *aos_pool = nullptr;
  uint32_t *aos_idx = nullptr;
  size_t aos_pool_bytes = 0, aos_idx_cap = 0;
  const void *aos_host_pool = nullptr;  // identity of the image: host pool pointer, slots and particle count
  int64_t aos_slots = 0, aos_n = 0;
  int *scan_a = nullptr, *scan_b = nullptr;  // flags / prefix scratch (downloads, seeding)
  size_t scan_cap = 0;
  void *scan_tmp = nullptr;
  size_t scan_tmp_bytes = 0;
  uint32_t id_base = 0;
  int num_sms = 148;
  int nt2_global = 0;  // tile layers of the whole domain (P.nt[2] is slab-local)
  int grid_p2g = 148 * 4, grid_g2p = 148 * 4;  // persistent grids = SMs x resident CTAs (queried)
  int64_t launches = 0;

  // two substeps captured as one CUDA graph per buffer parity: mpmb_substep replays it instead of re-launching
  // 16 (1 GPU) / 24 (z-slab rank) kernels from the host; invalidated by anything that changes a kernel argument
  cudaGraphExec_t graph_exec[2] = {nullptr, nullptr};
  int graph_launches = 0;
  cudaStream_t cap_stream = nullptr;
  bool use_graph = true;

  // CPIC rigid coupling (enabled by mpmb_set_rigid_samples with at least one sample)
  RigidView R{};
  bool rigid_on = false;
  float *rs_offset = nullptr, *rs_tri = nullptr;
  int *rs_rigid = nullptr;
  size_t rigid_nodes = 0;
  int64_t id_span_pending = 0;
  int64_t id_span = 0;   // ids of the resident particles lie in [id_base, id_base + id_span): uploads number 0..n-1, the seeded
                         // lattice numbers by lattice index (gaps where the boundary band is skipped)

  bool profiling = false;
  struct Ev { cudaEvent_t a, b; int stage; };
  std::vector<Ev> events;
  size_t events_used = 0;
  double prof_ms[MPMB_N_STAGES] = {};
  int64_t prof_launches[MPMB_N_STAGES] = {};
};

static thread_local std::string g_create_error;

static int fail(MpmbEngine *h, int code, const char *fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (h) h->err = buf;
  else g_create_error = buf;
  return code;
}

#define CUDA_TRY(h, expr)                                                                      \
  do {                                                                                         \
    cudaError_t e_ = (expr);                                                                   \
    if (e_ != cudaSuccess) {                                                                   \
      (h)->sticky_cuda = true;                                                                 \
      return fail((h), MPMB_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
    }                                                                                          \
  } while (0)

#define CHECK_HANDLE(h)                                                       \
  do {                                                                        \
    if (!(h)) return fail(nullptr, MPMB_ERR_INVALID, "null handle");          \
    if ((h)->sticky_cuda) return MPMB_ERR_CUDA;                               \
    cudaError_t e0_ = cudaSetDevice((h)->cfg.device);                         \
    if (e0_ != cudaSuccess) return fail((h), MPMB_ERR_CUDA, "cudaSetDevice"); \
  } while (0)

static void graph_reset(MpmbEngine *h) {
  for (int p = 0; p < 2; p++)
    if (h->graph_exec[p]) { cudaGraphExecDestroy(h->graph_exec[p]); h->graph_exec[p] = nullptr; }
}

static View make_view(MpmbEngine *h) {
  View V{};
  for (int k = 0; k < N_Q; k++) {
    V.q[k] = h->q[h->cur][k];
    V.qn[k] = h->q[h->cur ^ 1][k];
  }
  V.keys = h->keys[h->cur];
  V.keys_next = h->keys[h->cur ^ 1];
  V.outpos = h->outpos;
  V.run_begin = h->run_begin[h->ord];
  V.run_len = h->run_len[h->ord];
  V.out_begin = h->run_begin[h->ord ^ 1];
  V.total = h->run_len[h->ord ^ 1];
  V.stay_cnt = h->stay[h->ord];
  V.stay_next = h->stay[h->ord ^ 1];
  V.arr_cnt = h->arr_cnt;
  V.arr_off = h->arr_off;
  V.arr_len = h->arr_len;
  V.arr_cur = h->arr_cur;
  V.slot_map = h->slot_map;
  V.meta = h->meta;
  V.mover_dst = h->mover_dst[h->mov];
  V.mover_idx = h->mover_idx[h->mov];
  V.mover_dst_n = h->mover_dst[h->mov ^ 1];
  V.mover_idx_n = h->mover_idx[h->mov ^ 1];
  V.arrivals = h->arrivals;
  V.arrivals_sorted = h->arrivals_sorted;
  V.blocksum = h->blocksum;
  V.arena = h->arena;
  V.sdf4 = h->sdf4;
  V.cnt = h->cnt;
  V.cap_tiles = h->cap_tiles;
  V.cap_particles = (int)h->cap;
  V.rigid_flag = h->rigid_on ? h->R.tile_flag : nullptr;
  return V;
}

static void prof_begin(MpmbEngine *h, int stage) {
  if (!h->profiling) return;
  if (h->events_used == h->events.size()) {
    MpmbEngine::Ev e{};
    cudaEventCreate(&e.a);
    cudaEventCreate(&e.b);
    h->events.push_back(e);
  }
  h->events[h->events_used].stage = stage;
  cudaEventRecord(h->events[h->events_used].a, h->stream);
}
static void prof_end(MpmbEngine *h, int n_launches) {
  if (!h->profiling) return;
  cudaEventRecord(h->events[h->events_used].b, h->stream);
  h->prof_launches[h->events[h->events_used].stage] += n_launches;
  h->events_used++;
}
static void prof_collect(MpmbEngine *h) {
  for (size_t i = 0; i < h->events_used; i++) {
    float ms = 0.f;
    cudaEventSynchronize(h->events[i].b);
    cudaEventElapsedTime(&ms, h->events[i].a, h->events[i].b);
    h->prof_ms[h->events[i].stage] += ms;
  }
  h->events_used = 0;
}

static int free_particles(MpmbEngine *h) {
  graph_reset(h);
  for (int b = 0; b < 2; b++) {
    for (int k = 0; k < N_Q; k++) { cudaFree(h->q[b][k]); h->q[b][k] = nullptr; }
    cudaFree(h->keys[b]); h->keys[b] = nullptr;
    cudaFree(h->mover_dst[b]); cudaFree(h->mover_idx[b]);
    h->mover_dst[b] = h->mover_idx[b] = nullptr;
  }
  cudaFree(h->outpos); cudaFree(h->arrivals); cudaFree(h->arrivals_sorted); cudaFree(h->keys_sorted); cudaFree(h->iota); cudaFree(h->cub_temp);
  h->outpos = h->arrivals = h->arrivals_sorted = h->keys_sorted = h->iota = nullptr;
  h->cub_temp = nullptr;
  h->cap = 0;
  return 0;
}

static int alloc_particles(MpmbEngine *h, int64_t cap) {
  free_particles(h);
  if (cap >= (1ll << 30)) return fail(h, MPMB_ERR_CAPACITY, "capacity %lld exceeds 2^30 particle rows per GPU", (long long)cap);
  if (cap < 1) cap = 1;
  for (int b = 0; b < 2; b++) {
    for (int k = 0; k < N_Q; k++) CUDA_TRY(h, cudaMalloc(&h->q[b][k], sizeof(float4) * cap));
    CUDA_TRY(h, cudaMalloc(&h->keys[b], sizeof(uint32_t) * cap));
    CUDA_TRY(h, cudaMalloc(&h->mover_dst[b], sizeof(uint32_t) * cap));
    CUDA_TRY(h, cudaMalloc(&h->mover_idx[b], sizeof(uint32_t) * cap));
  }
  CUDA_TRY(h, cudaMalloc(&h->outpos, sizeof(uint32_t) * (cap + 8)));  // the bulk copy of a run's words is widened to 16-byte bounds
  CUDA_TRY(h, cudaMalloc(&h->arrivals, sizeof(uint32_t) * cap));
  CUDA_TRY(h, cudaMalloc(&h->arrivals_sorted, sizeof(uint32_t) * cap));
  CUDA_TRY(h, cudaMalloc(&h->keys_sorted, sizeof(uint32_t) * cap));
  CUDA_TRY(h, cudaMalloc(&h->iota, sizeof(uint32_t) * cap));
  k_iota<<<(unsigned)((cap + 255) / 256), 256, 0, h->stream>>>(h->iota, (int)cap);
  h->cub_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, h->cub_bytes, h->keys[0], h->keys_sorted, h->iota, h->arrivals_sorted, (int)cap, 0, 32, h->stream);
  CUDA_TRY(h, cudaMalloc(&h->cub_temp, h->cub_bytes));
  h->cap = cap;
  return MPMB_OK;
}

extern "C" {

static int read_n_store(MpmbEngine *h, int *n_store);

int mpmb_version(void) { return MPMB_VERSION; }

const char *mpmb_last_error(MpmbHandle h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int mpmb_create(const MpmbConfig *cfg, MpmbHandle *out) {
  if (!cfg || !out) return fail(nullptr, MPMB_ERR_INVALID, "null argument");
  *out = nullptr;
  for (int d = 0; d < 3; d++)
    if (cfg->res[d] < 16 || cfg->res[d] > 4096) return fail(nullptr, MPMB_ERR_INVALID, "res[%d]=%d out of range [16,4096]", d, cfg->res[d]);
  if (!(cfg->dx > 0.f) || !(cfg->dt > 0.f)) return fail(nullptr, MPMB_ERR_INVALID, "dx and dt must be positive");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) return fail(nullptr, MPMB_ERR_CUDA, "no CUDA device: %s", cudaGetErrorString(e));
  if (cfg->device < 0 || cfg->device >= ndev) return fail(nullptr, MPMB_ERR_INVALID, "device %d not in [0,%d)", cfg->device, ndev);
  MpmbEngine *h = new MpmbEngine();
  h->cfg = *cfg;
  if (h->cfg.world <= 0) h->cfg.world = 1;
  Params &P = h->P;
  size_t ntot = 1;
  for (int d = 0; d < 3; d++) {
    P.res[d] = cfg->res[d];
    P.nnode[d] = cfg->res[d] + 1;
    P.nt[d] = (P.nnode[d] + 3) / 4 + 1;
    P.gdt[d] = cfg->gravity[d] * cfg->dt;
  }
  h->nt2_global = P.nt[2];
  P.tz_off = 0;
  if (h->cfg.world > 1) {
    if (cfg->tile_z0 < 0 || cfg->tile_z1 > P.nt[2] || cfg->tile_z0 >= cfg->tile_z1) {
      const int n2 = P.nt[2];
      delete h;
      return fail(nullptr, MPMB_ERR_INVALID, "bad slab [%d,%d) for %d tile layers", cfg->tile_z0, cfg->tile_z1, n2);
    }
    P.tz_off = cfg->tile_z0 - 1;                   // one ghost layer below ...
    P.nt[2] = cfg->tile_z1 - cfg->tile_z0 + 2;     // ... and one above the owned layers
  }
  for (int d = 0; d < 3; d++) ntot *= (size_t)P.nt[d];
  P.ntiles_total = (int)ntot;
  P.dx = cfg->dx;
  P.inv_dx = 1.0f / cfg->dx;
  P.dt = cfg->dt;
  P.particle_gravity = cfg->particle_gravity;
  P.clean_boundary = cfg->clean_boundary;
  P.friction = 0.f;
  P.has_sdf = 0;
  P.world = h->cfg.world;
  P.tile_z0 = h->cfg.world > 1 ? cfg->tile_z0 : 0;
  P.tile_z1 = h->cfg.world > 1 ? cfg->tile_z1 : P.nt[2];
  for (int g = 0; g < MPMB_MAX_GROUPS; g++) {
    P.mats[g].kind = MAT_JELLY;  // JellyParticle defaults E=1e5, nu=0.3 (src/particles.cpp:383-389)
    for (int k = 0; k < 8; k++) P.mats[g].p[k] = 0.f;
    P.mats[g].p[0] = 1e5f / (2.f * 1.3f);
    P.mats[g].p[1] = 1e5f * 0.3f / (1.3f * 0.4f);
  }
  h->use_graph = cfg->no_graph == 0;
  if (const char *g = getenv("MPMB_GRAPH")) h->use_graph = g[0] != '0';   // MPMB_GRAPH=0: launch every kernel from the host (A/B, debugging)
  h->special_min = (uint32_t)ntot;
  h->key_dead = (uint32_t)ntot + SPECIAL_DEAD;
  h->key_bits = 1;
  while ((1ull << h->key_bits) <= (unsigned long long)ntot + SPECIAL_DEAD) h->key_bits++;
  if (h->key_bits > 31) { delete h; return fail(nullptr, MPMB_ERR_INVALID, "tile grid too large for 32-bit keys"); }
  if (cudaSetDevice(cfg->device) != cudaSuccess) { delete h; return fail(nullptr, MPMB_ERR_CUDA, "cudaSetDevice failed"); }
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, cfg->device);
  h->num_sms = prop.multiProcessorCount;
  {
    int occ = 0;
    cudaFuncSetAttribute(k_p2g, cudaFuncAttributeMaxDynamicSharedMemorySize, P2G_DYN_BYTES);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_p2g, P2G_T, P2G_DYN_BYTES) == cudaSuccess && occ > 0) h->grid_p2g = h->num_sms * occ;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_g2p<128, true, true>, 128, 0) == cudaSuccess && occ > 0) h->grid_g2p = h->num_sms * occ;
  }
  // tiles: every tile of the (slab of the) domain can be active
  int64_t cap_tiles = (int64_t)ntot;
  h->cap_tiles = (int)cap_tiles;
  auto bail = [&](const char *what) {
    std::string msg = std::string(what) + ": " + cudaGetErrorString(cudaGetLastError());
    mpmb_destroy(h);
    return fail(nullptr, MPMB_ERR_CUDA, "%s", msg.c_str());
  };
  h->ntot = (int)ntot;
  h->ord_blocks = (int)((ntot + ORD_TILE - 1) / ORD_TILE);
  {
    int **dense[] = {&h->run_begin[0], &h->run_begin[1], &h->run_len[0], &h->run_len[1], &h->stay[0], &h->stay[1],
                     &h->arr_cnt, &h->arr_off, &h->arr_len, &h->arr_cur, &h->slot_map};
    for (int **pp : dense) {
      if (cudaMalloc(pp, sizeof(int) * ntot) != cudaSuccess) return bail("cudaMalloc dense tile array");
      cudaMemset(*pp, 0, sizeof(int) * ntot);
    }
  }
  if (cudaMalloc(&h->blocksum, sizeof(int) * 3 * (h->ord_blocks + 1)) != cudaSuccess) return bail("cudaMalloc blocksum");
  if (cudaMalloc(&h->meta, sizeof(TileMeta) * cap_tiles) != cudaSuccess) return bail("cudaMalloc meta");
  if (cudaMalloc(&h->arena, sizeof(float4) * ARENA * (cap_tiles + 1)) != cudaSuccess) return bail("cudaMalloc arena");
  cudaMemset(h->arena + (size_t)ARENA * cap_tiles, 0, sizeof(float4) * ARENA);  // the all-zero arena read for absent neighbours
  if (cudaMalloc(&h->vel, sizeof(float4) * ARENA * cap_tiles) != cudaSuccess) return bail("cudaMalloc vel");
  if (cudaMalloc(&h->cnt, sizeof(Counters)) != cudaSuccess) return bail("cudaMalloc counters");
  cudaMemset(h->slot_map, 0xFF, sizeof(int) * ntot);
  cudaMemset(h->cnt, 0, sizeof(Counters));
  h->mig_cap = h->cfg.world > 1 ? (cfg->migrate_capacity > 0 ? cfg->migrate_capacity : 65536) : 0;
  if (h->cfg.world > 1) {
    const int64_t bytes[2] = {mpmb_halo_bytes(h), mpmb_migrate_bytes(h)};
    for (int k = 0; k < 2; k++)
      for (int f = 0; f < 2; f++) {
        if (cudaMalloc(&h->rx[k][f], bytes[k]) != cudaSuccess) return bail("cudaMalloc exchange buffer");
        cudaMemset(h->rx[k][f], 0, 16);
      }
    if (cudaMalloc(&h->xcount, sizeof(int) * 32) != cudaSuccess) return bail("cudaMalloc exchange counters");
    cudaMemset(h->xcount, 0, sizeof(int) * 32);  // [0],[4] halo counts, [8],[12] migration counts, [16],[20] finished-CTA counters
  }
  if (cfg->capacity > 0) {
    int rc = alloc_particles(h, cfg->capacity);
    if (rc != MPMB_OK) { g_create_error = h->err; mpmb_destroy(h); return rc; }
  }
  *out = h;
  return MPMB_OK;
}

int mpmb_destroy(MpmbHandle h) {
  if (!h) return MPMB_OK;
  cudaSetDevice(h->cfg.device);
  cudaDeviceSynchronize();
  free_particles(h);
  for (int b = 0; b < 2; b++) { cudaFree(h->run_begin[b]); cudaFree(h->run_len[b]); cudaFree(h->stay[b]); }
  cudaFree(h->arr_cnt); cudaFree(h->arr_off); cudaFree(h->arr_len); cudaFree(h->arr_cur); cudaFree(h->slot_map); cudaFree(h->blocksum);
  cudaFree(h->meta);
  cudaFree(h->arena); cudaFree(h->vel); cudaFree(h->sdf4); cudaFree(h->cnt); cudaFree(h->stage_buf); cudaFree(h->xcount);
  cudaFree(h->aos_pool); cudaFree(h->aos_idx); cudaFree(h->scan_a); cudaFree(h->scan_b); cudaFree(h->scan_tmp);
  for (int k = 0; k < 2; k++)
    for (int f = 0; f < 2; f++) {
      if (h->tx_ipc[k][f] && h->tx[k][f]) cudaIpcCloseMemHandle(h->tx[k][f]);
      cudaFree(h->rx[k][f]);
    }
  for (auto &e : h->events) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); }
  graph_reset(h);
  if (h->cap_stream) cudaStreamDestroy(h->cap_stream);
  cudaFree(h->R.bodies); cudaFree(h->rs_offset); cudaFree(h->rs_tri); cudaFree(h->rs_rigid); cudaFree(h->R.s_base); cudaFree(h->R.node_key);
  cudaFree(h->R.node_tags); cudaFree(h->R.page); cudaFree(h->R.tile_flag); cudaFree(h->R.p_states); cudaFree(h->R.p_cdf); cudaFree(h->R.p_mark);

  delete h;
  return MPMB_OK;
}

int mpmb_set_stream(MpmbHandle h, void *s) {
  CHECK_HANDLE(h);
  h->stream = (cudaStream_t)s;
  return MPMB_OK;
}

int mpmb_synchronize(MpmbHandle h) {
  CHECK_HANDLE(h);
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  Counters c;
  CUDA_TRY(h, cudaMemcpy(&c, h->cnt, sizeof(c), cudaMemcpyDeviceToHost));
  if (c.error & DEVERR_TILE_CAPACITY) return fail(h, MPMB_ERR_CAPACITY, "active tile capacity exceeded");
  if (c.error & DEVERR_MIGRATE_CAPACITY) return fail(h, MPMB_ERR_CAPACITY, "migration buffer capacity exceeded");
  if (c.error & DEVERR_PARTICLE_CAPACITY) return fail(h, MPMB_ERR_CAPACITY, "particle capacity exceeded");
  if (c.error & DEVERR_PEER_TIMEOUT) return fail(h, MPMB_ERR_STATE, "peer exchange timed out (neighbour rank not stepping)");
  if (c.error & DEVERR_BAD_INPUT) return fail(h, MPMB_ERR_INVALID, "uploaded particles with group outside [0,%d) or mass <= 0", MPMB_MAX_GROUPS);
  return MPMB_OK;
}

int mpmb_set_material(MpmbHandle h, int32_t group, int32_t kind, const float *params, int32_t n_params) {
  CHECK_HANDLE(h);
  if (group < 0 || group >= MPMB_MAX_GROUPS) return fail(h, MPMB_ERR_INVALID, "group %d out of range", group);
  if (kind < MPMB_MAT_LINEAR || kind > MPMB_MAT_VISCO) return fail(h, MPMB_ERR_INVALID, "unknown material kind %d", kind);
  if (n_params < 0 || n_params > MPMB_MAT_PARAMS || (n_params > 0 && !params)) return fail(h, MPMB_ERR_INVALID, "bad parameter vector");
  h->P.mats[group].kind = kind;
  for (int k = 0; k < 8; k++) h->P.mats[group].p[k] = k < n_params ? params[k] : 0.f;
  graph_reset(h);
  // particles of this group already resident: their cached affine matrix belongs to the old material
  if (h->cap > 0 && h->stage == 0) {
    int ns = 0;
    int rc = read_n_store(h, &ns);
    if (rc != MPMB_OK) return rc;
    if (ns > 0) {
      View V = make_view(h);
      k_refresh_affine<<<(ns + 127) / 128, 128, 0, h->stream>>>(V, h->P, h->keys[h->cur], ns, h->special_min, group);
      h->launches++;
      CUDA_TRY(h, cudaGetLastError());
    }
  }
  return MPMB_OK;
}

int mpmb_set_delta_t(MpmbHandle h, float dt) {
  CHECK_HANDLE(h);
  if (!(dt > 0.f) || !std::isfinite(dt)) return fail(h, MPMB_ERR_INVALID, "delta_t must be positive and finite");
  if (h->stage != 0) return fail(h, MPMB_ERR_STATE, "delta_t changes between substeps, not inside one");
  if (dt == h->P.dt) return MPMB_OK;
  h->cfg.dt = dt;
  h->P.dt = dt;
  for (int d = 0; d < 3; d++) h->P.gdt[d] = h->cfg.gravity[d] * dt;
  graph_reset(h);
  // resident particles: the affine matrix cached for the next rasterize carries the old dt (stress * (-4 dt / dx))
  if (h->cap > 0) {
    int ns = 0;
    int rc = read_n_store(h, &ns);
    if (rc != MPMB_OK) return rc;
    if (ns > 0) {
      View V = make_view(h);
      for (int g = 0; g < MPMB_MAX_GROUPS; g++) {
        k_refresh_affine<<<(ns + 127) / 128, 128, 0, h->stream>>>(V, h->P, h->keys[h->cur], ns, h->special_min, g);
        h->launches++;
      }
      CUDA_TRY(h, cudaGetLastError());
    }
  }
  return MPMB_OK;
}

int mpmb_set_id_base(MpmbHandle h, int64_t base) {
  CHECK_HANDLE(h);
  if (base < 0 || base >= (1ll << TAG_ID_BITS)) return fail(h, MPMB_ERR_INVALID, "id base %lld outside [0, 2^%d)", (long long)base, TAG_ID_BITS);
  h->id_base = (uint32_t)base;
  return MPMB_OK;
}

int mpmb_set_sdf(MpmbHandle h, const float *sdf4, float friction) {
  CHECK_HANDLE(h);
  size_t n = (size_t)h->P.nnode[0] * h->P.nnode[1] * h->P.nnode[2];
  if (!sdf4) {
    h->P.has_sdf = 0;
    graph_reset(h);
    return MPMB_OK;
  }
  if (!h->sdf4) CUDA_TRY(h, cudaMalloc(&h->sdf4, sizeof(float4) * n));
  CUDA_TRY(h, cudaMemcpyAsync(h->sdf4, sdf4, sizeof(float4) * n, cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  h->P.has_sdf = 1;
  h->P.friction = friction;
  graph_reset(h);
  return MPMB_OK;
}

int mpmb_set_planes(MpmbHandle h, int32_t n_planes, const float *planes4, float friction) {
  CHECK_HANDLE(h);
  if (n_planes <= 0 || n_planes > 64 || !planes4) return fail(h, MPMB_ERR_INVALID, "need 1..64 planes");
  size_t n = (size_t)h->P.nnode[0] * h->P.nnode[1] * h->P.nnode[2];
  if (!h->sdf4) CUDA_TRY(h, cudaMalloc(&h->sdf4, sizeof(float4) * n));
  float4 *d_planes = nullptr;
  CUDA_TRY(h, cudaMalloc(&d_planes, sizeof(float4) * n_planes));
  CUDA_TRY(h, cudaMemcpyAsync(d_planes, planes4, sizeof(float4) * n_planes, cudaMemcpyHostToDevice, h->stream));
  k_planes_to_sdf<<<h->num_sms * 8, 256, 0, h->stream>>>(h->P, n_planes, d_planes, h->sdf4);
  h->launches++;
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  cudaFree(d_planes);
  h->P.has_sdf = 1;
  h->P.friction = friction;
  graph_reset(h);
  return MPMB_OK;
}

int mpmb_set_levelset_shapes(MpmbHandle h, int32_t n_shapes, const MpmbShape *shapes, float friction) {
  CHECK_HANDLE(h);
  if (n_shapes <= 0 || n_shapes > 64 || !shapes) return fail(h, MPMB_ERR_INVALID, "need 1..64 shapes");
  for (int k = 0; k < n_shapes; k++)
    if (shapes[k].kind < MPMB_SHAPE_PLANE || shapes[k].kind > MPMB_SHAPE_CUBOID) return fail(h, MPMB_ERR_INVALID, "unknown shape kind %d", shapes[k].kind);
  size_t n = (size_t)h->P.nnode[0] * h->P.nnode[1] * h->P.nnode[2];
  if (!h->sdf4) CUDA_TRY(h, cudaMalloc(&h->sdf4, sizeof(float4) * n));
  MpmbShape *d_shapes = nullptr;
  CUDA_TRY(h, cudaMalloc(&d_shapes, sizeof(MpmbShape) * n_shapes));
  CUDA_TRY(h, cudaMemcpyAsync(d_shapes, shapes, sizeof(MpmbShape) * n_shapes, cudaMemcpyHostToDevice, h->stream));
  k_shapes_to_sdf<<<h->num_sms * 8, 256, 0, h->stream>>>(h->P, n_shapes, d_shapes, h->sdf4);
  h->launches++;
  cudaError_t e = cudaStreamSynchronize(h->stream);
  cudaFree(d_shapes);
  CUDA_TRY(h, e);
  h->P.has_sdf = 1;
  h->P.friction = friction;
  graph_reset(h);
  return MPMB_OK;
}

static int ensure_capacity(MpmbEngine *h, int64_t n) {
  int64_t want = n;
  if (h->cfg.world > 1) want = n + n / 4 + 16 * h->mig_cap;
  if (h->cfg.capacity > 0) {
    if (n > h->cap) return fail(h, MPMB_ERR_CAPACITY, "%lld particles exceed the configured capacity %lld", (long long)n, (long long)h->cap);
    return MPMB_OK;
  }
  if (want > h->cap) return alloc_particles(h, want);
  return MPMB_OK;
}

static int rigid_reserve_ids(MpmbEngine *h, int64_t n_ids);
static int finish_upload(MpmbEngine *h, int64_t n) {
  // Fresh storage in upload order: no runs yet, every particle is an "arrival" of its tile.  The
  // one radix sort of the engine's life (per upload) groups the rows by tile; the ordinary
  // ordering scan then lays out the first runs.
  h->id_span = std::max<int64_t>(h->id_span_pending, n);
  h->id_span_pending = 0;
  if (h->rigid_on) {  // new particles start with MPMParticle::states = 0 (mpmb_set_particle_states overrides)
    int rc = rigid_reserve_ids(h, std::max<int64_t>(h->id_span, h->cap));
    if (rc != MPMB_OK) return rc;
    CUDA_TRY(h, cudaMemsetAsync(h->R.p_states, 0, sizeof(uint32_t) * h->R.id_cap, h->stream));
  }
  View V = make_view(h);
  const size_t dense = sizeof(int) * (size_t)h->ntot;
  CUDA_TRY(h, cudaMemsetAsync(h->run_begin[h->ord], 0, dense, h->stream));
  CUDA_TRY(h, cudaMemsetAsync(h->run_len[h->ord], 0, dense, h->stream));
  CUDA_TRY(h, cudaMemsetAsync(h->stay[h->ord], 0, dense, h->stream));
  CUDA_TRY(h, cudaMemsetAsync(h->arr_cnt, 0, dense, h->stream));
  Counters c{};
  c.n_store = (int)n;
  CUDA_TRY(h, cudaMemcpyAsync(h->cnt, &c, sizeof(int) * 3, cudaMemcpyHostToDevice, h->stream));          // n_store, n_tiles, n_ghost
  CUDA_TRY(h, cudaMemsetAsync(&h->cnt->n_movers, 0, sizeof(int) * 3, h->stream));                          // n_movers, n_movers_next, n_alive
  if (n > 0) {
    cub::DeviceRadixSort::SortPairs(h->cub_temp, h->cub_bytes, h->keys[h->cur], h->keys_sorted, h->iota, h->arrivals_sorted, (int)n, 0, h->key_bits, h->stream);
    k_count_sorted<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(V, h->keys_sorted, (int)n, h->ntot);
    h->launches++;
  }
  h->fresh = true;
  h->stage = 0;
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  CUDA_TRY(h, cudaGetLastError());
  return MPMB_OK;
}

int mpmb_upload_particles(MpmbHandle h, int64_t n, const float *x, const float *v, const float *F, const float *b,
                          const float *mass, const float *vol, const float *scalar, const int32_t *group) {
  CHECK_HANDLE(h);
  if (n < 0 || (n > 0 && (!x || !v || !mass || !vol))) return fail(h, MPMB_ERR_INVALID, "x, v, mass, vol are required");
  if ((int64_t)h->id_base + n > (1ll << TAG_ID_BITS)) return fail(h, MPMB_ERR_CAPACITY, "particle ids exceed 2^%d", TAG_ID_BITS);
  h->aos_host_pool = nullptr;
  int rc = ensure_capacity(h, n);
  if (rc != MPMB_OK) return rc;
  if (n == 0) return finish_upload(h, 0);
  // stage the field arrays on the device (one allocation)
  size_t fl = (size_t)n * (3 + 3 + (F ? 9 : 0) + (b ? 9 : 0) + 1 + 1 + (scalar ? 1 : 0)) + (group ? (size_t)n : 0);
  float *stage = nullptr;
  if (h->stage_bytes < fl * sizeof(float)) {
    cudaFree(h->stage_buf);
    h->stage_buf = nullptr;
    h->stage_bytes = 0;
    CUDA_TRY(h, cudaMalloc(&h->stage_buf, fl * sizeof(float)));
    h->stage_bytes = fl * sizeof(float);
  }
  stage = (float *)h->stage_buf;
  float *p = stage;
  auto put = [&](const void *src, size_t count) -> float * {
    if (!src) return nullptr;
    float *dst = p;
    cudaMemcpyAsync(dst, src, count * sizeof(float), cudaMemcpyHostToDevice, h->stream);
    p += count;
    return dst;
  };
  float *dx_ = put(x, 3 * n), *dv = put(v, 3 * n), *dF = put(F, 9 * n), *db = put(b, 9 * n), *dm = put(mass, n), *dvol = put(vol, n),
        *ds = put(scalar, n);
  int *dg = (int *)put(group, n);
  View V = make_view(h);
  k_pack_particles<<<(unsigned)((n + 127) / 128), 128, 0, h->stream>>>(V, h->P, (int)n, dx_, dv, dF, db, dm, dvol, ds, dg, h->keys[h->cur], h->id_base);
  h->launches++;
  cudaError_t e = cudaStreamSynchronize(h->stream);
  CUDA_TRY(h, e);
  CUDA_TRY(h, cudaGetLastError());
  return finish_upload(h, n);
}

// flags / exclusive prefix scratch for n items (grow-only) and the CUB scan over them
static int scan_reserve(MpmbEngine *h, size_t n) {
  if (n + 1 > h->scan_cap) {
    cudaFree(h->scan_a); cudaFree(h->scan_b);
    h->scan_a = h->scan_b = nullptr;
    h->scan_cap = 0;
    CUDA_TRY(h, cudaMalloc(&h->scan_a, sizeof(int) * (n + 1)));
    CUDA_TRY(h, cudaMalloc(&h->scan_b, sizeof(int) * (n + 1)));
    h->scan_cap = n + 1;
  }
  size_t need = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, need, h->scan_a, h->scan_b, (int)(n + 1), h->stream);
  if (need > h->scan_tmp_bytes) {
    cudaFree(h->scan_tmp);
    h->scan_tmp = nullptr;
    h->scan_tmp_bytes = 0;
    CUDA_TRY(h, cudaMalloc(&h->scan_tmp, need));
    h->scan_tmp_bytes = need;
  }
  return MPMB_OK;
}
// prefix = exclusive scan of flags over n (+1 so that prefix[n] is the total); returns the total
static int scan_flags(MpmbEngine *h, size_t n, int *total) {
  CUDA_TRY(h, cudaMemsetAsync(h->scan_a + n, 0, sizeof(int), h->stream));
  size_t bytes = h->scan_tmp_bytes;
  cub::DeviceScan::ExclusiveSum(h->scan_tmp, bytes, h->scan_a, h->scan_b, (int)(n + 1), h->stream);
  CUDA_TRY(h, cudaMemcpyAsync(total, h->scan_b + n, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  return MPMB_OK;
}

// byte range [lo, hi) of a slot that the layout's fields occupy: only that window of every slot crosses the bus (a 2-D
// copy with the slot stride as pitch) — 204 of the reference's 320 bytes
static void aos_window(const MpmbAosLayout *L, int *lo, int *hi) {
  const int mat = 2 * L->col_pitch + 12;  // three padded columns, the last one 3 floats long
  int a = std::min(std::min(L->off_pos, L->off_v_and_m), std::min(L->off_dg_e, L->off_apic_b));
  int b = std::max(std::max(L->off_pos + 12, L->off_v_and_m + 16), std::max(L->off_dg_e + mat, L->off_apic_b + mat));
  a = std::min(a, L->off_vol);
  b = std::max(b, L->off_vol + 4);
  if (L->off_scalar >= 0) { a = std::min(a, L->off_scalar); b = std::max(b, L->off_scalar + 4); }
  *lo = std::max(0, a & ~15);
  *hi = std::min(L->stride, (b + 15) & ~15);
}

static int aos_reserve(MpmbEngine *h, size_t pool_bytes, size_t n_idx) {
  if (pool_bytes > h->aos_pool_bytes) {
    cudaFree(h->aos_pool);
    h->aos_pool = nullptr;
    h->aos_pool_bytes = 0;
    CUDA_TRY(h, cudaMalloc(&h->aos_pool, pool_bytes));
    h->aos_pool_bytes = pool_bytes;
  }
  if (n_idx > h->aos_idx_cap) {
    cudaFree(h->aos_idx);
    h->aos_idx = nullptr;
    h->aos_idx_cap = 0;
    CUDA_TRY(h, cudaMalloc(&h->aos_idx, sizeof(uint32_t) * n_idx));
    h->aos_idx_cap = n_idx;
  }
  return MPMB_OK;
}

int mpmb_upload_aos(MpmbHandle h, int64_t n, const void *pool, int64_t pool_slots, const uint32_t *indices,
                    const MpmbAosLayout *L, const int32_t *group) {
  CHECK_HANDLE(h);
  if (n < 0 || !pool || !indices || !L || L->stride <= 0) return fail(h, MPMB_ERR_INVALID, "pool, indices and layout are required");
  if ((int64_t)h->id_base + n > (1ll << TAG_ID_BITS)) return fail(h, MPMB_ERR_CAPACITY, "particle ids exceed 2^%d", TAG_ID_BITS);
  int rc = ensure_capacity(h, n);
  if (rc != MPMB_OK) return rc;
  h->aos_host_pool = nullptr;
  if (n == 0) return finish_upload(h, 0);
  // the pool's device image and the index vector stay resident (grow-only buffers): mpmb_download_aos scatters the
  // results into the image and returns it with one contiguous copy
  const size_t pool_bytes = (size_t)pool_slots * L->stride;
  if ((rc = aos_reserve(h, pool_bytes, (size_t)n)) != MPMB_OK) return rc;
  int *d_grp = nullptr;
  if (group) {
    if (h->stage_bytes < sizeof(int) * (size_t)n) {
      cudaFree(h->stage_buf);
      h->stage_buf = nullptr;
      h->stage_bytes = 0;
      CUDA_TRY(h, cudaMalloc(&h->stage_buf, sizeof(int) * (size_t)n));
      h->stage_bytes = sizeof(int) * (size_t)n;
    }
    d_grp = (int *)h->stage_buf;
    CUDA_TRY(h, cudaMemcpyAsync(d_grp, group, sizeof(int) * n, cudaMemcpyHostToDevice, h->stream));
  }
  CUDA_TRY(h, cudaMemcpyAsync(h->aos_idx, indices, sizeof(uint32_t) * n, cudaMemcpyHostToDevice, h->stream));
  {
    int lo, hi;
    aos_window(L, &lo, &hi);
    CUDA_TRY(h, cudaMemcpy2DAsync(h->aos_pool + lo, (size_t)L->stride, (const char *)pool + lo, (size_t)L->stride, (size_t)(hi - lo), (size_t)pool_slots,
                                  cudaMemcpyHostToDevice, h->stream));
  }
  View V = make_view(h);
  k_pack_aos<<<(unsigned)((n + 127) / 128), 128, 0, h->stream>>>(V, h->P, (int)n, h->aos_pool, h->aos_idx, *L, d_grp, h->keys[h->cur], h->id_base);
  h->launches++;
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  CUDA_TRY(h, cudaGetLastError());
  h->aos_host_pool = pool;
  h->aos_slots = pool_slots;
  h->aos_n = n;
  return finish_upload(h, n);
}

int mpmb_seed_lattice(MpmbHandle h, const int32_t lo_cell[3], const int32_t hi_cell[3], float vol, float mass, float jitter, uint32_t seed,
                      int32_t group, const float v0[3], int64_t *n_seeded) {
  CHECK_HANDLE(h);
  if (!lo_cell || !hi_cell) return fail(h, MPMB_ERR_INVALID, "null argument");
  if (group < 0 || group >= MPMB_MAX_GROUPS || !(mass > 0.f) || !(vol > 0.f)) return fail(h, MPMB_ERR_INVALID, "bad group, mass or volume");
  SeedBox B{};
  size_t cells = 1;
  for (int d = 0; d < 3; d++) {
    B.lo[d] = lo_cell[d];
    B.n[d] = hi_cell[d] - lo_cell[d];
    if (B.n[d] <= 0 || lo_cell[d] < 0 || hi_cell[d] > h->P.res[d]) return fail(h, MPMB_ERR_INVALID, "empty or out-of-domain cell block");
    cells *= (size_t)B.n[d];
    B.v0[d] = v0 ? v0[d] : 0.f;
  }
  if (cells * 8 + (size_t)h->id_base > (size_t)(1ll << TAG_ID_BITS)) return fail(h, MPMB_ERR_CAPACITY, "lattice ids exceed 2^%d", TAG_ID_BITS);
  // candidate cell layers: a particle of cell layer kz has its base node in layer kz-1 or kz (|jitter| < 0.25)
  B.z_first = 0;
  B.z_count = B.n[2];
  if (h->cfg.world > 1) {
    const int c0 = std::max(B.lo[2], 4 * h->P.tile_z0 - 1), c1 = std::min(B.lo[2] + B.n[2], 4 * h->P.tile_z1 + 2);
    B.z_first = c0 - B.lo[2];
    B.z_count = std::max(0, c1 - c0);
  }
  B.jitter = jitter; B.vol = vol; B.mass = mass; B.seed = seed; B.id_base = h->id_base; B.group = group;
  const int kind = h->P.mats[group].kind;
  B.ps = (kind == MAT_SNOW || kind == MAT_WATER) ? 1.0f : 0.0f;
  const size_t n_cand = (size_t)B.n[0] * B.n[1] * (size_t)B.z_count * 8;
  h->aos_host_pool = nullptr;
  if (n_cand >= (size_t)0x7fffffff) return fail(h, MPMB_ERR_CAPACITY, "too many candidate particles for one rank");
  int total = 0, rc;
  if (n_cand > 0) {
    if ((rc = scan_reserve(h, n_cand)) != MPMB_OK) return rc;
    k_seed_flags<<<h->num_sms * 8, 256, 0, h->stream>>>(h->P, B, n_cand, h->scan_a);
    if ((rc = scan_flags(h, n_cand, &total)) != MPMB_OK) return rc;
  }
  if ((rc = ensure_capacity(h, total)) != MPMB_OK) return rc;
  if (total > 0) {
    View V = make_view(h);
    k_seed_write<<<h->num_sms * 8, 256, 0, h->stream>>>(V, h->P, B, n_cand, h->scan_a, h->scan_b, h->keys[h->cur]);
    h->launches += 2;
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    CUDA_TRY(h, cudaGetLastError());
  }
  if (n_seeded) *n_seeded = total;
  h->id_span_pending = (int64_t)n_cand;   // ids = lattice index
  return finish_upload(h, total);
}

static int read_n_store(MpmbEngine *h, int *n_store) {
  Counters c;
  CUDA_TRY(h, cudaMemcpy(&c, h->cnt, sizeof(c), cudaMemcpyDeviceToHost));
  *n_store = c.n_store;
  return MPMB_OK;
}

int mpmb_num_particles(MpmbHandle h, int64_t *n) {
  CHECK_HANDLE(h);
  if (!n) return fail(h, MPMB_ERR_INVALID, "null argument");
  *n = 0;
  if (h->cap == 0) return MPMB_OK;
  CUDA_TRY(h, cudaMemsetAsync(&h->cnt->n_live, 0, sizeof(int), h->stream));
  k_count_live<<<h->num_sms * 4, 256, 0, h->stream>>>(h->keys[h->cur], h->special_min, h->cnt);
  h->launches++;
  int rc = mpmb_synchronize(h);
  if (rc != MPMB_OK) return rc;
  int live = 0;
  CUDA_TRY(h, cudaMemcpy(&live, &h->cnt->n_live, sizeof(int), cudaMemcpyDeviceToHost));
  *n = live;
  return MPMB_OK;
}

int mpmb_get_update_count(MpmbHandle h, int64_t *updates) {
  CHECK_HANDLE(h);
  if (!updates) return fail(h, MPMB_ERR_INVALID, "null argument");
  int rc = mpmb_synchronize(h);
  if (rc != MPMB_OK) return rc;
  unsigned long long u = 0;
  CUDA_TRY(h, cudaMemcpy(&u, &h->cnt->updates, sizeof(u), cudaMemcpyDeviceToHost));
  *updates = (int64_t)u;
  return MPMB_OK;
}

int mpmb_download_particles(MpmbHandle h, int64_t cap, int64_t *n_out, uint32_t *id, float *x, float *v, float *F, float *b,
                            float *mass, float *vol, float *scalar, int32_t *group) {
  CHECK_HANDLE(h);
  if (!n_out) return fail(h, MPMB_ERR_INVALID, "n_out is required");
  int rc = mpmb_synchronize(h);
  if (rc != MPMB_OK) return rc;
  int n = 0;
  if ((rc = read_n_store(h, &n)) != MPMB_OK) return rc;
  *n_out = 0;
  if (n == 0 || h->cap == 0) return MPMB_OK;
  // exclusive prefix of the alive flags (CUB scan)
  int *flags = nullptr, *prefix = nullptr;
  CUDA_TRY(h, cudaMalloc(&flags, sizeof(int) * n));
  CUDA_TRY(h, cudaMalloc(&prefix, sizeof(int) * (n + 1)));
  k_alive_flags<<<(n + 255) / 256, 256, 0, h->stream>>>(h->keys[h->cur], n, h->special_min, flags);
  void *tmp = nullptr;
  size_t tmp_bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, flags, prefix, n, h->stream);
  CUDA_TRY(h, cudaMalloc(&tmp, tmp_bytes));
  cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, flags, prefix, n, h->stream);
  int last_flag = 0, last_prefix = 0;
  CUDA_TRY(h, cudaMemcpyAsync(&last_flag, flags + n - 1, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(h, cudaMemcpyAsync(&last_prefix, prefix + n - 1, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  const int64_t alive = (int64_t)last_flag + last_prefix;
  if (alive > cap) {
    cudaFree(flags); cudaFree(prefix); cudaFree(tmp);
    return fail(h, MPMB_ERR_CAPACITY, "%lld live particles do not fit in the %lld rows provided", (long long)alive, (long long)cap);
  }
  size_t fl = (size_t)alive * ((x ? 3 : 0) + (v ? 3 : 0) + (F ? 9 : 0) + (b ? 9 : 0) + (mass ? 1 : 0) + (vol ? 1 : 0) + (scalar ? 1 : 0) +
                               (id ? 1 : 0) + (group ? 1 : 0));
  float *stage = nullptr;
  if (h->stage_bytes < fl * sizeof(float)) {
    cudaFree(h->stage_buf);
    h->stage_buf = nullptr;
    h->stage_bytes = 0;
    CUDA_TRY(h, cudaMalloc(&h->stage_buf, fl * sizeof(float)));
    h->stage_bytes = fl * sizeof(float);
  }
  stage = (float *)h->stage_buf;
  float *p = stage;
  auto take = [&](bool want, size_t per) -> float * {
    if (!want) return nullptr;
    float *r = p;
    p += per * alive;
    return r;
  };
  float *dx_ = take(x, 3), *dv = take(v, 3), *dF = take(F, 9), *db = take(b, 9), *dm = take(mass, 1), *dvol = take(vol, 1), *ds = take(scalar, 1);
  uint32_t *did = (uint32_t *)take(id, 1);
  int *dg = (int *)take(group, 1);
  View V = make_view(h);
  k_unpack_particles<<<(n + 127) / 128, 128, 0, h->stream>>>(V, h->keys[h->cur], n, h->special_min, did, dx_, dv, dF, db, dm, dvol, ds, dg, prefix);
  h->launches += 2;
  auto get = [&](void *dst, const void *src, size_t per) {
    if (dst) cudaMemcpyAsync(dst, src, per * alive * sizeof(float), cudaMemcpyDeviceToHost, h->stream);
  };
  get(x, dx_, 3); get(v, dv, 3); get(F, dF, 9); get(b, db, 9); get(mass, dm, 1); get(vol, dvol, 1); get(scalar, ds, 1);
  get(id, did, 1); get(group, dg, 1);
  cudaError_t e = cudaStreamSynchronize(h->stream);
  cudaFree(flags); cudaFree(prefix); cudaFree(tmp);
  CUDA_TRY(h, e);
  CUDA_TRY(h, cudaGetLastError());
  *n_out = alive;
  return MPMB_OK;
}

int mpmb_download_aos(MpmbHandle h, void *pool, int64_t pool_slots, uint32_t *indices, int64_t n_indices, const MpmbAosLayout *L,
                      int64_t *n_alive) {
  CHECK_HANDLE(h);
  if (!pool || !indices || !L || !n_alive || L->stride <= 0 || n_indices < 0) return fail(h, MPMB_ERR_INVALID, "null argument");
  int ns = 0, rc;
  if ((rc = mpmb_synchronize(h)) != MPMB_OK) return rc;
  if ((rc = read_n_store(h, &ns)) != MPMB_OK) return rc;
  *n_alive = 0;
  if (n_indices == 0) return MPMB_OK;
  const size_t pool_bytes = (size_t)pool_slots * L->stride;
  // the device image of the pool: still resident from mpmb_upload_aos of the same pool, else brought in now
  const bool resident = h->aos_host_pool == pool && h->aos_slots == pool_slots && h->aos_n == n_indices && h->aos_pool;
  if (!resident) {
    if ((rc = aos_reserve(h, pool_bytes, (size_t)n_indices)) != MPMB_OK) return rc;
    CUDA_TRY(h, cudaMemcpyAsync(h->aos_idx, indices, sizeof(uint32_t) * n_indices, cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(h, cudaMemcpyAsync(h->aos_pool, pool, pool_bytes, cudaMemcpyHostToDevice, h->stream));
  }
  (void)pool_bytes;
  if ((rc = scan_reserve(h, (size_t)n_indices)) != MPMB_OK) return rc;
  CUDA_TRY(h, cudaMemsetAsync(h->scan_a, 0, sizeof(int) * (size_t)n_indices, h->stream));
  if (ns > 0 && h->cap > 0) {
    View V = make_view(h);
    k_aos_scatter<<<(ns + 127) / 128, 128, 0, h->stream>>>(V, h->keys[h->cur], ns, h->special_min, h->aos_pool, h->aos_idx, n_indices, pool_slots, *L,
                                                           h->id_base, h->scan_a, h->cnt);
    h->launches++;
  }
  int alive = 0;
  if ((rc = scan_flags(h, (size_t)n_indices, &alive)) != MPMB_OK) return rc;
  // survivors' slots in id order == what clear_boundary_particles leaves in MPM::particles (src/mpm.cpp:618-622)
  uint32_t *d_surv = (uint32_t *)h->keys_sorted;  // upload-time scratch, free between uploads; cap >= n rows
  if ((int64_t)h->cap < n_indices) return fail(h, MPMB_ERR_CAPACITY, "index vector longer than the particle capacity");
  k_compact_survivors<<<(unsigned)((n_indices + 255) / 256), 256, 0, h->stream>>>(h->aos_idx, h->scan_a, h->scan_b, n_indices, d_surv);
  h->launches++;
  {  // only the window of every slot that holds the layout's fields comes back; the rest of the host pool is untouched
    int lo, hi;
    aos_window(L, &lo, &hi);
    CUDA_TRY(h, cudaMemcpy2DAsync((char *)pool + lo, (size_t)L->stride, h->aos_pool + lo, (size_t)L->stride, (size_t)(hi - lo), (size_t)pool_slots,
                                  cudaMemcpyDeviceToHost, h->stream));
  }
  if (alive > 0) CUDA_TRY(h, cudaMemcpyAsync(indices, d_surv, sizeof(uint32_t) * (size_t)alive, cudaMemcpyDeviceToHost, h->stream));
  if ((rc = mpmb_synchronize(h)) != MPMB_OK) return rc;   // also surfaces ids / slots outside the vectors
  CUDA_TRY(h, cudaGetLastError());
  h->aos_host_pool = nullptr;   // the host owns the pool again: the next upload refreshes the image
  *n_alive = alive;
  return MPMB_OK;
}

int mpmb_download_bgeo_points(MpmbHandle h, int64_t id_range, void *records, int64_t cap_records, int64_t *n_out) {
  CHECK_HANDLE(h);
  if (!records || !n_out || id_range <= 0) return fail(h, MPMB_ERR_INVALID, "bad argument");
  int ns = 0, rc;
  if ((rc = mpmb_synchronize(h)) != MPMB_OK) return rc;
  if ((rc = read_n_store(h, &ns)) != MPMB_OK) return rc;
  *n_out = 0;
  if (ns == 0 || h->cap == 0) return MPMB_OK;
  if ((rc = scan_reserve(h, (size_t)id_range)) != MPMB_OK) return rc;
  CUDA_TRY(h, cudaMemsetAsync(h->scan_a, 0, sizeof(int) * (size_t)id_range, h->stream));
  View V = make_view(h);
  k_id_flags<<<(ns + 255) / 256, 256, 0, h->stream>>>(V, h->keys[h->cur], ns, h->special_min, h->id_base, id_range, h->scan_a, h->cnt);
  int alive = 0;
  if ((rc = scan_flags(h, (size_t)id_range, &alive)) != MPMB_OK) return rc;
  if (alive > cap_records) return fail(h, MPMB_ERR_CAPACITY, "%d live particles do not fit in the %lld records provided", alive, (long long)cap_records);
  const size_t bytes = (size_t)alive * 48;
  if (h->stage_bytes < bytes) {
    cudaFree(h->stage_buf);
    h->stage_buf = nullptr;
    h->stage_bytes = 0;
    CUDA_TRY(h, cudaMalloc(&h->stage_buf, bytes));
    h->stage_bytes = bytes;
  }
  k_bgeo_points<<<(ns + 127) / 128, 128, 0, h->stream>>>(V, h->keys[h->cur], ns, h->special_min, h->id_base, id_range, h->scan_b, (uint32_t *)h->stage_buf);
  h->launches += 2;
  CUDA_TRY(h, cudaMemcpyAsync(records, h->stage_buf, bytes, cudaMemcpyDeviceToHost, h->stream));
  if ((rc = mpmb_synchronize(h)) != MPMB_OK) return rc;
  CUDA_TRY(h, cudaGetLastError());
  *n_out = alive;
  return MPMB_OK;
}

// ------------------------------------------------------------------------------ CPIC host side
static int rigid_reserve_ids(MpmbEngine *h, int64_t n_ids) {
  if (n_ids <= h->R.id_cap && h->R.p_states) return MPMB_OK;
  const int64_t cap = std::max<int64_t>(n_ids, 1024);
  uint32_t *st = nullptr;
  float4 *cdf = nullptr;
  uint32_t *nr = nullptr;
  CUDA_TRY(h, cudaMalloc(&st, sizeof(uint32_t) * cap));
  CUDA_TRY(h, cudaMalloc(&cdf, sizeof(float4) * cap));
  CUDA_TRY(h, cudaMalloc(&nr, sizeof(uint32_t) * cap));
  CUDA_TRY(h, cudaMemsetAsync(st, 0, sizeof(uint32_t) * cap, h->stream));
  CUDA_TRY(h, cudaMemsetAsync(cdf, 0, sizeof(float4) * cap, h->stream));
  CUDA_TRY(h, cudaMemsetAsync(nr, 0, sizeof(uint32_t) * cap, h->stream));   // epoch 0 is never current
  if (h->R.p_states && h->R.id_cap > 0) CUDA_TRY(h, cudaMemcpyAsync(st, h->R.p_states, sizeof(uint32_t) * h->R.id_cap, cudaMemcpyDeviceToDevice, h->stream));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  cudaFree(h->R.p_states); cudaFree(h->R.p_cdf); cudaFree(h->R.p_mark);
  h->R.p_states = st; h->R.p_cdf = cdf; h->R.p_mark = nr;
  h->R.id_cap = (int)cap;
  return MPMB_OK;
}

// per substep, after the ordering: pages + colour field of this pose, tile flags, particle colours
static int rigid_prepare(MpmbEngine *h, const View &V, int *nl) {
  RigidView &R = h->R;
  R.id_base = h->id_base;
  if (!R.p_states || R.id_cap < h->id_span) { int rc = rigid_reserve_ids(h, std::max<int64_t>(h->id_span, h->cap)); if (rc != MPMB_OK) return rc; }
  R.epoch = (R.epoch + 1u) & 0x7fffffffu;
  if (R.epoch == 0u) R.epoch = 1u;
  const int ns = R.n_samples, sb = (ns + 127) / 128;
  k_cdf_clear<<<sb, 128, 0, h->stream>>>(R, h->P);
  CUDA_TRY(h, cudaMemsetAsync(R.page, 0, (size_t)R.nb[0] * R.nb[1] * R.nb[2], h->stream));
  k_cdf_raster<<<sb, 128, 0, h->stream>>>(R, h->P);
  k_rigid_tile_flags<<<(h->cap_tiles + 127) / 128, 128, 0, h->stream>>>(V, h->P, R);
  k_gather_cdf<<<(int)((h->cap + 127) / 128), 128, 0, h->stream>>>(V, h->P, R);
  *nl += 4;
  return MPMB_OK;
}

int mpmb_set_rigid_samples(MpmbHandle h, int32_t n_bodies, int64_t n_samples, const float *offset3, const float *tri9, const int32_t *rigid_id) {
  CHECK_HANDLE(h);
  if (h->cfg.world > 1) return fail(h, MPMB_ERR_STATE, "rigid coupling is implemented for single-GPU engines (world == 1)");
  if (n_samples == 0) { h->rigid_on = false; graph_reset(h); return MPMB_OK; }
  if (n_bodies < 2 || n_bodies > RIGID_MAX || n_samples < 0 || n_samples > (1 << 24) || !offset3 || !tri9 || !rigid_id)
    return fail(h, MPMB_ERR_INVALID, "need 2..%d bodies (entry 0 is the background body) and their boundary samples", RIGID_MAX);
  if (h->P.nnode[0] > 1023 || h->P.nnode[1] > 1023 || h->P.nnode[2] > 1023) return fail(h, MPMB_ERR_INVALID, "rigid coupling supports grids up to 1022 cells per axis");
  for (int64_t s = 0; s < n_samples; s++)
    if (rigid_id[s] < 1 || rigid_id[s] >= n_bodies) return fail(h, MPMB_ERR_INVALID, "sample %lld names body %d (valid: 1..%d)", (long long)s, rigid_id[s], n_bodies - 1);
  RigidView &R = h->R;
  cudaFree(h->rs_offset); cudaFree(h->rs_tri); cudaFree(h->rs_rigid); cudaFree(R.s_base);
  h->rs_offset = h->rs_tri = nullptr; h->rs_rigid = nullptr; R.s_base = nullptr;
  CUDA_TRY(h, cudaMalloc(&h->rs_offset, sizeof(float) * 3 * n_samples));
  CUDA_TRY(h, cudaMalloc(&h->rs_tri, sizeof(float) * 9 * n_samples));
  CUDA_TRY(h, cudaMalloc(&h->rs_rigid, sizeof(int) * n_samples));
  CUDA_TRY(h, cudaMalloc(&R.s_base, sizeof(int) * n_samples));
  CUDA_TRY(h, cudaMemcpyAsync(h->rs_offset, offset3, sizeof(float) * 3 * n_samples, cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(h, cudaMemcpyAsync(h->rs_tri, tri9, sizeof(float) * 9 * n_samples, cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(h, cudaMemcpyAsync(h->rs_rigid, rigid_id, sizeof(int) * n_samples, cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(h, cudaMemsetAsync(R.s_base, 0xff, sizeof(int) * n_samples, h->stream));
  R.s_offset = h->rs_offset; R.s_tri = h->rs_tri; R.s_rigid = h->rs_rigid;
  R.n_samples = (int)n_samples;
  R.n_bodies = n_bodies;
  if (!R.bodies) {
    CUDA_TRY(h, cudaMalloc(&R.bodies, sizeof(RigidDev) * RIGID_MAX));
    CUDA_TRY(h, cudaMemsetAsync(R.bodies, 0, sizeof(RigidDev) * RIGID_MAX, h->stream));
  }
  const size_t nn = (size_t)h->P.nnode[0] * h->P.nnode[1] * h->P.nnode[2];
  if (!R.node_key) {
    CUDA_TRY(h, cudaMalloc(&R.node_key, sizeof(unsigned long long) * nn));
    CUDA_TRY(h, cudaMalloc(&R.node_tags, sizeof(uint32_t) * nn));
    R.nb[0] = (h->P.nnode[0] + 3) / 4 + 1; R.nb[1] = (h->P.nnode[1] + 3) / 4 + 1; R.nb[2] = (h->P.nnode[2] + 7) / 8 + 1;
    CUDA_TRY(h, cudaMalloc(&R.page, (size_t)R.nb[0] * R.nb[1] * R.nb[2]));
    CUDA_TRY(h, cudaMalloc(&R.tile_flag, (size_t)std::max(h->cap_tiles, 1)));
    CUDA_TRY(h, cudaMemsetAsync(R.tile_flag, 0, (size_t)std::max(h->cap_tiles, 1), h->stream));
    h->rigid_nodes = nn;
  }
  CUDA_TRY(h, cudaMemsetAsync(R.node_key, 0xff, sizeof(unsigned long long) * nn, h->stream));
  CUDA_TRY(h, cudaMemsetAsync(R.node_tags, 0, sizeof(uint32_t) * nn, h->stream));
  if (R.pushing_force == 0.f && R.penalty == 0.f) R.pushing_force = 20000.0f;   // src/mpm.cpp:35,40 defaults
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  h->rigid_on = true;
  // poses change between substeps: the host drives every substep, so mpmb_substep does not replay graphs while bodies are present
  // (use_graph itself keeps what the configuration asked for: switching the coupling off brings graph replay back)
  graph_reset(h);
  return MPMB_OK;
}

int mpmb_set_rigid_coupling(MpmbHandle h, float penalty, float pushing_force) {
  CHECK_HANDLE(h);
  h->R.penalty = penalty;
  h->R.pushing_force = pushing_force;
  return MPMB_OK;
}

int mpmb_set_rigid_state(MpmbHandle h, int32_t n_bodies, const MpmbRigidBody *bodies) {
  CHECK_HANDLE(h);
  if (!h->rigid_on) return fail(h, MPMB_ERR_STATE, "no rigid bodies: call mpmb_set_rigid_samples first");
  if (n_bodies != h->R.n_bodies || !bodies) return fail(h, MPMB_ERR_INVALID, "expected %d bodies", h->R.n_bodies);
  RigidDev dev[RIGID_MAX];
  memset(dev, 0, sizeof(dev));
  for (int b = 0; b < n_bodies; b++) {
    const MpmbRigidBody &s = bodies[b];
    RigidDev &d = dev[b];
    for (int k = 0; k < 3; k++) { d.pos[k] = s.position[k]; d.vel[k] = s.velocity[k]; d.ang[k] = s.angular_velocity[k]; }
    for (int k = 0; k < 9; k++) { d.rot[k] = s.rot[k]; d.inv_inertia[k] = s.inv_inertia[k]; }
    d.inv_mass = s.inv_mass;
    d.fric[0] = s.frictions[0]; d.fric[1] = s.frictions[1];
  }
  CUDA_TRY(h, cudaMemcpyAsync(h->R.bodies, dev, sizeof(RigidDev) * RIGID_MAX, cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));   // `dev` lives on this stack frame
  return MPMB_OK;
}

int mpmb_get_rigid_state(MpmbHandle h, int32_t n_bodies, MpmbRigidBody *bodies) {
  CHECK_HANDLE(h);
  if (!h->rigid_on) return fail(h, MPMB_ERR_STATE, "no rigid bodies");
  if (n_bodies != h->R.n_bodies || !bodies) return fail(h, MPMB_ERR_INVALID, "expected %d bodies", h->R.n_bodies);
  RigidDev dev[RIGID_MAX];
  CUDA_TRY(h, cudaMemcpyAsync(dev, h->R.bodies, sizeof(RigidDev) * RIGID_MAX, cudaMemcpyDeviceToHost, h->stream));
  int rc = mpmb_synchronize(h);
  if (rc != MPMB_OK) return rc;
  for (int b = 0; b < n_bodies; b++) {
    MpmbRigidBody &s = bodies[b];
    const RigidDev &d = dev[b];
    for (int k = 0; k < 3; k++) { s.position[k] = d.pos[k]; s.velocity[k] = d.vel[k]; s.angular_velocity[k] = d.ang[k]; }
    for (int k = 0; k < 9; k++) { s.rot[k] = d.rot[k]; s.inv_inertia[k] = d.inv_inertia[k]; }
    s.inv_mass = d.inv_mass;
    s.frictions[0] = d.fric[0]; s.frictions[1] = d.fric[1];
  }
  return MPMB_OK;
}

int mpmb_set_particle_states(MpmbHandle h, int64_t n, const uint32_t *states) {
  CHECK_HANDLE(h);
  if (!h->rigid_on) return fail(h, MPMB_ERR_STATE, "no rigid bodies");
  if (n < 0 || (n > 0 && !states)) return fail(h, MPMB_ERR_INVALID, "bad argument");
  int rc = rigid_reserve_ids(h, std::max<int64_t>(n, h->cap));
  if (rc != MPMB_OK) return rc;
  CUDA_TRY(h, cudaMemcpyAsync(h->R.p_states, states, sizeof(uint32_t) * n, cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  return MPMB_OK;
}

int mpmb_get_particle_cdf(MpmbHandle h, int64_t n, uint32_t *states, float *normal3, float *distance, uint8_t *near_boundary) {
  CHECK_HANDLE(h);
  if (!h->rigid_on || !h->R.p_states) return fail(h, MPMB_ERR_STATE, "no rigid bodies / no substep yet");
  if (n < 0 || n > h->R.id_cap) return fail(h, MPMB_ERR_INVALID, "n exceeds the %d particle ids tracked", h->R.id_cap);
  std::vector<float4> cdf((size_t)n);
  std::vector<uint32_t> mark((size_t)n);
  if (states) CUDA_TRY(h, cudaMemcpyAsync(states, h->R.p_states, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(h, cudaMemcpyAsync(mark.data(), h->R.p_mark, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(h, cudaMemcpyAsync(cdf.data(), h->R.p_cdf, sizeof(float4) * n, cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  for (int64_t i = 0; i < n; i++) {   // particles gather_cdf did not see in the last substep: distance 0, normal 0, not near (138-146)
    const bool fresh = (mark[i] >> 1) == h->R.epoch;
    if (normal3) { normal3[3 * i] = fresh ? cdf[i].x : 0.f; normal3[3 * i + 1] = fresh ? cdf[i].y : 0.f; normal3[3 * i + 2] = fresh ? cdf[i].z : 0.f; }
    if (distance) distance[i] = fresh ? cdf[i].w : 0.f;
    if (near_boundary) near_boundary[i] = (fresh && (mark[i] & 1u)) ? 1 : 0;
  }
  return MPMB_OK;
}

int mpmb_download_cdf(MpmbHandle h, uint32_t *node_states, float *node_distance) {
  CHECK_HANDLE(h);
  if (!h->rigid_on) return fail(h, MPMB_ERR_STATE, "no rigid bodies");
  const size_t nn = h->rigid_nodes;
  std::vector<unsigned long long> key(nn);
  std::vector<uint32_t> tags(nn);
  CUDA_TRY(h, cudaMemcpyAsync(key.data(), h->R.node_key, sizeof(unsigned long long) * nn, cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(h, cudaMemcpyAsync(tags.data(), h->R.node_tags, sizeof(uint32_t) * nn, cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  for (size_t i = 0; i < nn; i++) {
    const bool has = key[i] != CDF_NO_KEY;
    if (node_states) node_states[i] = (tags[i] & CDF_TAG_MASK) | (has ? ((uint32_t)key[i] & 0xffu) << 24 : 0u);
    if (node_distance) {
      uint32_t bits = (uint32_t)(key[i] >> 32);
      float d;
      memcpy(&d, &bits, 4);
      node_distance[i] = has ? d * h->P.dx : 0.f;
    }
  }
  return MPMB_OK;
}

// ------------------------------------------------------------------------------ stages
int mpmb_sort_particles_and_populate_grid(MpmbHandle h) {
  CHECK_HANDLE(h);
  if (h->stage != 0) return fail(h, MPMB_ERR_STATE, "sort must follow resample/upload");
  prof_begin(h, 0);
  int nl = 3;
  if (h->cap > 0) {
    View V = make_view(h);
    k_order_a<<<h->ord_blocks, ORD_B, 0, h->stream>>>(V, h->ntot);
    k_order_b<<<1, 1024, 0, h->stream>>>(V, h->ord_blocks);
    k_order_c<<<h->ord_blocks, ORD_B, 0, h->stream>>>(V, h->ntot, h->P.nt[1], h->P.nt[2], h->P.tz_off);
    if (!h->fresh) {
      k_mover_place<<<h->num_sms * 2, 256, 0, h->stream>>>(V, h->ntot);
      k_mover_rank<<<h->num_sms * 2, 256, 0, h->stream>>>(V, h->ntot);
      nl += 2;
    }
#ifdef MPMB_VALIDATE
    if (!h->fresh) k_validate_order<<<h->num_sms * 2, 256, 0, h->stream>>>(V, h->ntot);
#endif
    h->fresh = false;
    if (h->rigid_on) {
      int rc = rigid_prepare(h, V, &nl);
      if (rc != MPMB_OK) return rc;
    }
  }
  h->launches += nl;
  prof_end(h, nl);
  CUDA_TRY(h, cudaGetLastError());
  h->stage = 1;
  return MPMB_OK;
}

int mpmb_rasterize(MpmbHandle h) {
  CHECK_HANDLE(h);
  if (h->stage != 1) return fail(h, MPMB_ERR_STATE, "rasterize must follow sort_particles_and_populate_grid");
  prof_begin(h, 1);
  View V = make_view(h);
  if (h->cap > 0) k_p2g<<<h->grid_p2g, P2G_T, P2G_DYN_BYTES, h->stream>>>(V, h->P, 0);
  h->launches += 1;
  if (h->cap > 0 && h->rigid_on) {  // block_op_rigid for the tiles of rigid pages, then apply_tmp_velocity (src/transfer.cpp:578-580)
    k_p2g_rigid<<<h->num_sms * 4, 128, 0, h->stream>>>(V, h->P, h->R);
    k_rigid_apply<<<1, 32, 0, h->stream>>>(h->R);
    h->launches += 2;
  }
  prof_end(h, 1);
  CUDA_TRY(h, cudaGetLastError());
  h->stage = 2;
  return MPMB_OK;
}

// the k_g2p instantiation for this scene: apic_b stored or not, the three later material kinds compiled in or not
static void launch_g2p(MpmbEngine *h, const View &V, int part, int commit, bool store_b) {
  bool ext = false;
  for (int g = 0; g < MPMB_MAX_GROUPS; g++) ext |= h->P.mats[g].kind > MAT_SAND;
  if (ext) {
    if (store_b) k_g2p<128, true, true><<<h->grid_g2p, 128, 0, h->stream>>>(V, h->P, h->vel, part, commit);
    else k_g2p<128, false, true><<<h->grid_g2p, 128, 0, h->stream>>>(V, h->P, h->vel, part, commit);
  } else {
    if (store_b) k_g2p<128, true, false><<<h->grid_g2p, 128, 0, h->stream>>>(V, h->P, h->vel, part, commit);
    else k_g2p<128, false, false><<<h->grid_g2p, 128, 0, h->stream>>>(V, h->P, h->vel, part, commit);
  }
}

int mpmb_resample(MpmbHandle h) {
  CHECK_HANDLE(h);
  if (h->stage != 2) return fail(h, MPMB_ERR_STATE, "resample must follow rasterize");
  View V = make_view(h);
  prof_begin(h, 4);
  if (h->cap > 0) k_grid<<<h->num_sms * 8, 256, 0, h->stream>>>(V, h->P, h->vel, 0);
  prof_end(h, 1);
  prof_begin(h, 2);
  if (h->cap > 0) {
    if (h->rigid_on) {
      k_g2p_rigid<<<h->num_sms * 4, 128, 0, h->stream>>>(V, h->P, h->R, h->vel);
      h->launches += 1;
    }
    launch_g2p(h, V, 0, 1, !h->skip_b || h->rigid_on);   // rigid tiles read apic_b back (the impulse needs dt * force)
    if (h->rigid_on) {
      k_rigid_apply<<<1, 32, 0, h->stream>>>(h->R);       // src/transfer.cpp:967-969
      h->launches += 1;
    }
  }
  h->launches += 2;
  prof_end(h, 1);
  CUDA_TRY(h, cudaGetLastError());
  h->cur ^= 1;  // the buffer G2P wrote is the current storage ...
  h->ord ^= 1;  // ... its runs / stay counts are the current ones ...
  h->mov ^= 1;  // ... and its mover list feeds the next ordering
  h->xstep++;
  h->stage = 0;
  return MPMB_OK;
}

static bool peers_connected(MpmbEngine *h);
static int xchg_halo_fused(MpmbEngine *h);
static int xchg_migrate_fused(MpmbEngine *h);
int mpmb_halo_send(MpmbHandle h, int32_t face);
int mpmb_halo_recv(MpmbHandle h, int32_t face);
int mpmb_migrate_send(MpmbHandle h, int32_t face);
int mpmb_migrate_recv(MpmbHandle h, int32_t face);

// z-slab runs: the same two stages in two launches each, boundary-layer tiles (part 1) and the
// rest (part 2), so that the host can put the halo exchange between them (include/mpmb.h).
int mpmb_rasterize_part(MpmbHandle h, int32_t part) {
  CHECK_HANDLE(h);
  if (part != 1 && part != 2) return fail(h, MPMB_ERR_INVALID, "part must be 1 (boundary) or 2 (interior)");
  if (!((part == 1 && h->stage == 1) || (part == 2 && h->stage == 11))) return fail(h, MPMB_ERR_STATE, "rasterize_part: boundary first, then interior, after sort");
  prof_begin(h, 1);
  View V = make_view(h);
  if (h->cap > 0) k_p2g<<<h->grid_p2g, P2G_T, P2G_DYN_BYTES, h->stream>>>(V, h->P, part);
  h->launches += 1;
  prof_end(h, 1);
  CUDA_TRY(h, cudaGetLastError());
  h->stage = part == 1 ? 11 : 2;
  return MPMB_OK;
}

int mpmb_resample_part(MpmbHandle h, int32_t part) {
  CHECK_HANDLE(h);
  if (part != 1 && part != 2) return fail(h, MPMB_ERR_INVALID, "part must be 1 (boundary) or 2 (interior)");
  if (!((part == 2 && h->stage == 2) || (part == 1 && h->stage == 22))) return fail(h, MPMB_ERR_STATE, "resample_part: interior first, then boundary, after rasterize");
  prof_begin(h, 2);
  View V = make_view(h);
  int nl = 2;
  if (h->cap > 0) {
    k_grid<<<h->num_sms * 8, 256, 0, h->stream>>>(V, h->P, h->vel, part);
    launch_g2p(h, V, part, part == 1 ? 1 : 0, true);
  }
  h->launches += nl;
  prof_end(h, nl);
  CUDA_TRY(h, cudaGetLastError());
  if (part == 2) { h->stage = 22; return MPMB_OK; }
  h->cur ^= 1;
  h->ord ^= 1;
  h->mov ^= 1;
  h->xstep++;
  h->stage = 0;
  return MPMB_OK;
}

// one substep; skip_b: do not store apic_b (an intermediate substep of mpmb_substep)
static int substep_once(MpmbEngine *h, bool peers, bool skip_b) {
  int rc;
  if ((rc = mpmb_sort_particles_and_populate_grid(h)) != MPMB_OK) return rc;
  if ((rc = mpmb_rasterize(h)) != MPMB_OK) return rc;
  if (peers && (rc = xchg_halo_fused(h)) != MPMB_OK) return rc;  // boundary-layer arenas into the neighbours' memory, theirs in as ghosts
  h->skip_b = skip_b;
#ifdef MPMB_CHECKED
  { static const char *dbg = getenv("MPMB_DBG_SKIPB"); if (dbg) h->skip_b = dbg[0] == '1'; }   // debug: force one k_g2p instantiation
#endif
  rc = mpmb_resample(h);
  h->skip_b = false;
  if (rc != MPMB_OK) return rc;
  if (peers && (rc = xchg_migrate_fused(h)) != MPMB_OK) return rc;
  return MPMB_OK;
}

// Captures two substeps (both buffer parities) into a graph for the parity the engine is in, and leaves the engine's
// host-side state as if they had run; the caller launches the graph.  Any failure turns graphs off for this handle.
static int graph_capture_pair(MpmbEngine *h, bool peers) {
  const int par = h->cur;
  if (!h->cap_stream && cudaStreamCreateWithFlags(&h->cap_stream, cudaStreamNonBlocking) != cudaSuccess) { h->use_graph = false; cudaGetLastError(); return MPMB_OK; }
  cudaStream_t user = h->stream;
  const int64_t l0 = h->launches;
  const int cur0 = h->cur, ord0 = h->ord, mov0 = h->mov, x0 = h->xstep;
  h->stream = h->cap_stream;
  cudaGraph_t g = nullptr;
  cudaError_t e = cudaStreamBeginCapture(h->cap_stream, cudaStreamCaptureModeThreadLocal);
  int rc = MPMB_OK;
  if (e == cudaSuccess) {
    rc = substep_once(h, peers, true);
    if (rc == MPMB_OK) rc = substep_once(h, peers, true);
    e = cudaStreamEndCapture(h->cap_stream, &g);
  }
  h->stream = user;
  if (e == cudaSuccess && rc == MPMB_OK && g) e = cudaGraphInstantiate(&h->graph_exec[par], g, 0);
  if (g) cudaGraphDestroy(g);
  if (e != cudaSuccess || rc != MPMB_OK || !h->graph_exec[par]) {
    cudaGetLastError();
    h->graph_exec[par] = nullptr;
    h->use_graph = false;
    h->cur = cur0; h->ord = ord0; h->mov = mov0; h->xstep = x0; h->stage = 0; h->launches = l0;   // nothing has run
    h->sticky_cuda = false;
    return MPMB_OK;
  }
  h->graph_launches = (int)(h->launches - l0);
  h->cur = cur0; h->ord = ord0; h->mov = mov0; h->xstep = x0; h->launches = l0;   // the caller's launch accounts for them
  return MPMB_OK;
}

int mpmb_substep(MpmbHandle h, int32_t nsub) {
  CHECK_HANDLE(h);
  const bool peers = peers_connected(h);
  if (h->cfg.world > 1 && !peers) return fail(h, MPMB_ERR_STATE, "z-slab run: connect the peers (mpmb_xchg_connect) or drive the stages and move the buffers yourself");
  int s = 0, rc;
  // the first substep after an upload takes the radix-sorted arrival lists (a different ordering launch): never in a graph
  if (h->fresh && nsub > 0) {
    if ((rc = substep_once(h, peers, nsub > 1)) != MPMB_OK) return rc;
    s = 1;
  }
  // pairs of intermediate substeps through the captured graph; the last substep (it stores apic_b) always runs directly
  while (h->use_graph && !h->rigid_on && !h->profiling && h->cap > 0 && h->stage == 0 && nsub - s >= 3) {
    const int par = h->cur;
    if (!h->graph_exec[par]) {
      if ((rc = graph_capture_pair(h, peers)) != MPMB_OK) return rc;
      if (!h->graph_exec[par]) break;
    }
    CUDA_TRY(h, cudaGraphLaunch(h->graph_exec[par], h->stream));
    h->launches += h->graph_launches;
    h->xstep += 2;   // cur / ord / mov flip twice: unchanged
    s += 2;
  }
  for (; s < nsub; s++)
    if ((rc = substep_once(h, peers, s + 1 < nsub)) != MPMB_OK) return rc;
  return MPMB_OK;
}

int mpmb_download_grid(MpmbHandle h, int32_t which, float *dense4) {
  CHECK_HANDLE(h);
  if (h->stage != 2) return fail(h, MPMB_ERR_STATE, "the grid exists between rasterize and resample");
  if (!dense4 || which < 0 || which > 1) return fail(h, MPMB_ERR_INVALID, "bad argument");
  size_t n = (size_t)h->P.nnode[0] * h->P.nnode[1] * h->P.nnode[2];
  float4 *d = nullptr;
  CUDA_TRY(h, cudaMalloc(&d, sizeof(float4) * n));
  View V = make_view(h);
  k_dense_grid<<<h->num_sms * 8, 256, 0, h->stream>>>(V, h->P, which, d);
  h->launches++;
  cudaError_t e = cudaMemcpyAsync(dense4, d, sizeof(float4) * n, cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  cudaFree(d);
  CUDA_TRY(h, e);
  return MPMB_OK;
}

// ------------------------------------------------------------------------------ debug (not part of mpmb.h)
// Only in builds with -DMPMB_DEBUG_EXPORTS (python -m taichi_mpm_b200.build --define MPMB_DEBUG_EXPORTS --out ...):
// the product library exports exactly what include/mpmb.h declares.
#if defined(MPMB_CHECKED) || defined(MPMB_VALIDATE)
extern "C" int mpmb_debug_check(MpmbHandle h, int *out8) {
  CHECK_HANDLE(h);
  cudaStreamSynchronize(h->stream);
  Counters c;
  if (cudaMemcpy(&c, h->cnt, sizeof(c), cudaMemcpyDeviceToHost) != cudaSuccess) return MPMB_ERR_CUDA;
  for (int k = 0; k < 8; k++) out8[k] = c.chk[k];
  out8[4] = c.n_store; out8[5] = c.n_alive; out8[6] = c.n_movers; out8[7] = c.n_tiles;
  return MPMB_OK;
}
#endif
#ifdef MPMB_DEBUG_EXPORTS
// which: 0 run_begin 1 run_len 2 stay 3 arr_len 4 out_begin 5 total 6 arr_off   (dense, ntot ints)
extern "C" int mpmb_debug_dense(MpmbHandle h, int which, int *out) {
  CHECK_HANDLE(h);
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  View V = make_view(h);
  const int *src[] = {V.run_begin, V.run_len, V.stay_cnt, V.arr_len, V.out_begin, V.total, V.arr_off};
  CUDA_TRY(h, cudaMemcpy(out, src[which], sizeof(int) * h->ntot, cudaMemcpyDeviceToHost));
  return h->ntot;
}
// keys and signed masses of the first n rows of the current storage; counters
extern "C" int mpmb_debug_rows(MpmbHandle h, int n, uint32_t *keys, float *mass4, int *counters8) {
  CHECK_HANDLE(h);
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  if (keys) CUDA_TRY(h, cudaMemcpy(keys, h->keys[h->cur], sizeof(uint32_t) * n, cudaMemcpyDeviceToHost));
  if (mass4) CUDA_TRY(h, cudaMemcpy(mass4, h->q[h->cur][0], sizeof(float4) * n, cudaMemcpyDeviceToHost));
  if (counters8) CUDA_TRY(h, cudaMemcpy(counters8, h->cnt, sizeof(Counters), cudaMemcpyDeviceToHost));
  return MPMB_OK;
}
#endif  // MPMB_DEBUG_EXPORTS

// ------------------------------------------------------------------------------ profiling
int mpmb_set_profiling(MpmbHandle h, int32_t enabled) {
  CHECK_HANDLE(h);
  if (!enabled && h->profiling) prof_collect(h);
  h->profiling = enabled != 0;
  return MPMB_OK;
}

int mpmb_get_profile(MpmbHandle h, double ms[MPMB_N_STAGES], int64_t launches[MPMB_N_STAGES], int32_t reset) {
  CHECK_HANDLE(h);
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  prof_collect(h);
  for (int s = 0; s < MPMB_N_STAGES; s++) {
    if (ms) ms[s] = h->prof_ms[s];
    if (launches) launches[s] = h->prof_launches[s];
    if (reset) { h->prof_ms[s] = 0; h->prof_launches[s] = 0; }
  }
  return MPMB_OK;
}

int mpmb_get_counters(MpmbHandle h, int64_t *active_tiles, int64_t *alive, int64_t *kernel_launches) {
  CHECK_HANDLE(h);
  int rc = mpmb_synchronize(h);
  if (rc != MPMB_OK) return rc;
  Counters c;
  CUDA_TRY(h, cudaMemcpy(&c, h->cnt, sizeof(c), cudaMemcpyDeviceToHost));
  if (active_tiles) *active_tiles = c.n_tiles;
  if (alive) *alive = c.n_alive;
  if (kernel_launches) *kernel_launches = h->launches;
  return MPMB_OK;
}

// rows of the current storage, movers waiting for the next ordering (= holes of their old runs), ghost tiles
int mpmb_get_ordering_stats(MpmbHandle h, int64_t *rows, int64_t *movers, int64_t *ghost_tiles) {
  CHECK_HANDLE(h);
  int rc = mpmb_synchronize(h);
  if (rc != MPMB_OK) return rc;
  Counters c;
  CUDA_TRY(h, cudaMemcpy(&c, h->cnt, sizeof(c), cudaMemcpyDeviceToHost));
  if (rows) *rows = c.n_store;
  if (movers) *movers = c.n_movers;
  if (ghost_tiles) *ghost_tiles = c.n_ghost;
  return MPMB_OK;
}

// ------------------------------------------------------------------------------ multi-GPU (z slabs)
static int64_t halo_cap_xy(MpmbEngine *h) {
  int64_t full = (int64_t)h->P.nt[0] * h->P.nt[1];
  return (h->cfg.halo_capacity > 0 && h->cfg.halo_capacity < full) ? h->cfg.halo_capacity : full;
}
static int64_t halo_idx_bytes(MpmbEngine *h) { return (halo_cap_xy(h) * 4 + 15) / 16 * 16; }

int64_t mpmb_halo_bytes(MpmbHandle h) {
  if (!h || h->cfg.world <= 1) return 0;
  return 16 + halo_idx_bytes(h) + halo_cap_xy(h) * ARENA * (int64_t)sizeof(float4);
}

int mpmb_halo_pack(MpmbHandle h, int32_t face, void *dev_buf) {
  CHECK_HANDLE(h);
  if (h->cfg.world <= 1) return fail(h, MPMB_ERR_STATE, "halo exchange needs world > 1");
  if (h->stage != 2 && h->stage != 11) return fail(h, MPMB_ERR_STATE, "halo_pack must follow rasterize (or its boundary part)");
  if (!dev_buf || face < 0 || face > 1) return fail(h, MPMB_ERR_INVALID, "bad argument");
  prof_begin(h, 3);
  char *b = (char *)dev_buf;
  CUDA_TRY(h, cudaMemsetAsync(b, 0, 16, h->stream));
  View V = make_view(h);
  const int layer = (face == 0 ? h->P.tile_z0 : h->P.tile_z1 - 1) - h->P.tz_off;
  k_halo_pack<<<h->num_sms * 4, 128, 0, h->stream>>>(V, h->P, layer, (int)halo_cap_xy(h), (int *)b, (int *)(b + 16),
                                                     (float4 *)(b + 16 + halo_idx_bytes(h)));
  h->launches++;
  prof_end(h, 1);
  CUDA_TRY(h, cudaGetLastError());
  return MPMB_OK;
}

int mpmb_halo_unpack(MpmbHandle h, int32_t face, const void *dev_buf) {
  CHECK_HANDLE(h);
  if (h->cfg.world <= 1) return fail(h, MPMB_ERR_STATE, "halo exchange needs world > 1");
  if (h->stage != 2 && h->stage != 22) return fail(h, MPMB_ERR_STATE, "halo_unpack must follow rasterize (and may follow the interior resample)");
  if (!dev_buf || face < 0 || face > 1) return fail(h, MPMB_ERR_INVALID, "bad argument");
  const int glayer = face == 0 ? h->P.tile_z0 - 1 : h->P.tile_z1;  // the neighbour's boundary layer
  if (glayer < 0 || glayer >= h->nt2_global) return fail(h, MPMB_ERR_INVALID, "no neighbour through face %d", face);
  const int layer = glayer - h->P.tz_off;
  prof_begin(h, 3);
  const char *b = (const char *)dev_buf;
  View V = make_view(h);
  k_halo_unpack<<<h->num_sms * 4, 128, 0, h->stream>>>(V, h->P, layer, (int)halo_cap_xy(h), (const int *)b, (const int *)(b + 16),
                                                       (const float4 *)(b + 16 + halo_idx_bytes(h)));
  h->launches++;
  prof_end(h, 1);
  CUDA_TRY(h, cudaGetLastError());
  return MPMB_OK;
}

int64_t mpmb_migrate_bytes(MpmbHandle h) {
  if (!h || h->cfg.world <= 1) return 0;
  return 16 + h->mig_cap * N_Q * (int64_t)sizeof(float4);
}

int mpmb_migrate_pack(MpmbHandle h, int32_t face, void *dev_buf) {
  CHECK_HANDLE(h);
  if (h->cfg.world <= 1) return fail(h, MPMB_ERR_STATE, "migration needs world > 1");
  if (h->stage != 0) return fail(h, MPMB_ERR_STATE, "migrate_pack must follow resample");
  if (!dev_buf || face < 0 || face > 1) return fail(h, MPMB_ERR_INVALID, "bad argument");
  prof_begin(h, 3);
  char *b = (char *)dev_buf;
  CUDA_TRY(h, cudaMemsetAsync(b, 0, 16, h->stream));
  View V = make_view(h);
  const uint32_t key_face = h->special_min + (face == 0 ? SPECIAL_MIG_DOWN : SPECIAL_MIG_UP);
  k_migrate_pack<<<h->num_sms * 2, 256, 0, h->stream>>>(V, key_face, h->key_dead, (int)h->mig_cap, (int *)b, (float4 *)(b + 16));
  h->launches++;
  prof_end(h, 1);
  CUDA_TRY(h, cudaGetLastError());
  return MPMB_OK;
}

int mpmb_migrate_unpack(MpmbHandle h, int32_t face, const void *dev_buf) {
  CHECK_HANDLE(h);
  if (h->cfg.world <= 1) return fail(h, MPMB_ERR_STATE, "migration needs world > 1");
  if (h->stage != 0) return fail(h, MPMB_ERR_STATE, "migrate_unpack must follow resample");
  if (!dev_buf || face < 0 || face > 1) return fail(h, MPMB_ERR_INVALID, "bad argument");
  prof_begin(h, 3);
  const char *b = (const char *)dev_buf;
  View V = make_view(h);
  k_migrate_unpack<<<64, 256, 0, h->stream>>>(V, h->P, (int)h->mig_cap, (const int *)b, (const float4 *)(b + 16));
  k_migrate_commit<<<1, 1, 0, h->stream>>>(h->cnt, (const int *)b, (int)h->mig_cap, (int)h->cap);
  h->launches += 2;
  prof_end(h, 2);
  CUDA_TRY(h, cudaGetLastError());
  return MPMB_OK;
}


// ------------------------------------------------------------------------------ peer-memory exchange
int mpmb_xchg_buffer(MpmbHandle h, int32_t kind, int32_t face, void **dev_ptr) {
  CHECK_HANDLE(h);
  if (h->cfg.world <= 1 || kind < 0 || kind > 1 || face < 0 || face > 1 || !dev_ptr) return fail(h, MPMB_ERR_INVALID, "bad argument");
  *dev_ptr = h->rx[kind][face];
  return MPMB_OK;
}

int mpmb_xchg_ipc_handle(MpmbHandle h, int32_t kind, int32_t face, void *out64) {
  CHECK_HANDLE(h);
  if (h->cfg.world <= 1 || kind < 0 || kind > 1 || face < 0 || face > 1 || !out64) return fail(h, MPMB_ERR_INVALID, "bad argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  cudaIpcMemHandle_t hd;
  CUDA_TRY(h, cudaIpcGetMemHandle(&hd, h->rx[kind][face]));
  memcpy(out64, &hd, 64);
  return MPMB_OK;
}

// Through face f this rank writes into the neighbour's receive buffer of the OPPOSITE face.  Give
// either that buffer's IPC handle (another process) or its device pointer (same process).
int mpmb_xchg_connect(MpmbHandle h, int32_t kind, int32_t face, const void *handle64, void *same_process_ptr) {
  CHECK_HANDLE(h);
  if (h->cfg.world <= 1 || kind < 0 || kind > 1 || face < 0 || face > 1) return fail(h, MPMB_ERR_INVALID, "bad argument");
  if (same_process_ptr) {
    h->tx[kind][face] = (char *)same_process_ptr;
    h->tx_ipc[kind][face] = false;
  } else {
    if (!handle64) return fail(h, MPMB_ERR_INVALID, "handle or pointer required");
    cudaIpcMemHandle_t hd;
    memcpy(&hd, handle64, 64);
    void *p = nullptr;
    CUDA_TRY(h, cudaIpcOpenMemHandle(&p, hd, cudaIpcMemLazyEnablePeerAccess));
    h->tx[kind][face] = (char *)p;
    h->tx_ipc[kind][face] = true;
  }
  // a (re)connection restarts the sequence: my own receive headers of this kind must not hold an old, larger seq
  for (int f = 0; f < 2; f++)
    if (h->rx[kind][f]) CUDA_TRY(h, cudaMemset(h->rx[kind][f], 0, 16));
  CUDA_TRY(h, cudaMemset(&h->cnt->xstep, 0, sizeof(int)));
  h->xstep = 0;
  graph_reset(h);
  return MPMB_OK;
}

static bool peers_connected(MpmbEngine *h) {
  if (h->cfg.world <= 1) return false;
  const bool lo = h->P.tile_z0 > 0 && h->cfg.rank > 0, hi = h->cfg.rank + 1 < h->cfg.world;
  for (int k = 0; k < 2; k++) {
    if (lo && !h->tx[k][0]) return false;
    if (hi && !h->tx[k][1]) return false;
  }
  return lo || hi;
}

// both faces of one message kind as the fused kernels take them
static void xfaces(MpmbEngine *h, int kind, XFace f[2], int *mask) {
  *mask = 0;
  for (int k = 0; k < 2; k++) {
    f[k].tx = h->tx[kind][k];
    f[k].rx = h->rx[kind][k];
    f[k].layer = 0;
    f[k].key = h->special_min + (k == 0 ? SPECIAL_MIG_DOWN : SPECIAL_MIG_UP);
    if (h->tx[kind][k]) *mask |= 1 << k;
  }
}

static int xchg_halo_launch(MpmbEngine *h, int mask_send, int mask_recv) {
  XFace s[2], r[2];
  int mask;
  xfaces(h, 0, s, &mask);
  xfaces(h, 0, r, &mask);
  s[0].layer = h->P.tile_z0 - h->P.tz_off;      s[1].layer = h->P.tile_z1 - 1 - h->P.tz_off;  // my boundary layers
  r[0].layer = h->P.tile_z0 - 1 - h->P.tz_off;  r[1].layer = h->P.tile_z1 - h->P.tz_off;      // the neighbours'
  prof_begin(h, 3);
  View V = make_view(h);
  int nl = 0;
  if (mask_send & mask) {
    k_halo_send2<<<h->num_sms * 4, 128, 0, h->stream>>>(V, h->P, (int)halo_cap_xy(h), (int)halo_idx_bytes(h), s[0], s[1], mask_send & mask, h->xcount,
                                                         h->xcount + 16);
    nl++;
  }
  if (mask_recv & mask) {
    k_halo_recv2<<<h->num_sms * 4, 128, 0, h->stream>>>(V, h->P, (int)halo_cap_xy(h), (int)halo_idx_bytes(h), r[0], r[1], mask_recv & mask);
    nl++;
  }
  h->launches += nl;
  prof_end(h, nl);
  CUDA_TRY(h, cudaGetLastError());
  return MPMB_OK;
}

static int xchg_migrate_launch(MpmbEngine *h, int mask_send, int mask_recv) {
  XFace f[2];
  int mask;
  xfaces(h, 1, f, &mask);
  prof_begin(h, 3);
  View V = make_view(h);
  int nl = 0;
  if (mask_send & mask) {
    k_migrate_send2<<<h->num_sms * 2, 256, 0, h->stream>>>(V, h->key_dead, (int)h->mig_cap, f[0], f[1], mask_send & mask, h->xcount + 8, h->xcount + 20);
    nl++;
  }
  if (mask_recv & mask) {
    k_migrate_recv2<<<1, 1024, 0, h->stream>>>(V, h->P, (int)h->mig_cap, f[0], f[1], mask_recv & mask);
    nl++;
  }
  h->launches += nl;
  prof_end(h, nl);
  CUDA_TRY(h, cudaGetLastError());
  return MPMB_OK;
}

// what mpmb_substep runs on a z-slab rank: both faces per launch
static int xchg_halo_fused(MpmbEngine *h) {
  if (h->stage != 2) return fail(h, MPMB_ERR_STATE, "halo exchange must follow rasterize");
  return xchg_halo_launch(h, 3, 3);
}
static int xchg_migrate_fused(MpmbEngine *h) {
  if (h->stage != 0) return fail(h, MPMB_ERR_STATE, "migration must follow resample");
  return xchg_migrate_launch(h, 3, 3);
}

// the same steps one face at a time, for hosts that drive the stages themselves (same kernels, one-face masks)
int mpmb_halo_send(MpmbHandle h, int32_t face) {
  CHECK_HANDLE(h);
  if (h->cfg.world <= 1 || face < 0 || face > 1 || !h->tx[0][face]) return fail(h, MPMB_ERR_STATE, "no connected peer through face %d", face);
  if (h->stage != 2) return fail(h, MPMB_ERR_STATE, "halo_send must follow rasterize");
  return xchg_halo_launch(h, 1 << face, 0);
}

int mpmb_halo_recv(MpmbHandle h, int32_t face) {
  CHECK_HANDLE(h);
  if (h->cfg.world <= 1 || face < 0 || face > 1 || !h->tx[0][face]) return fail(h, MPMB_ERR_STATE, "no connected peer through face %d", face);
  if (h->stage != 2) return fail(h, MPMB_ERR_STATE, "halo_recv must follow rasterize");
  return xchg_halo_launch(h, 0, 1 << face);
}

int mpmb_migrate_send(MpmbHandle h, int32_t face) {
  CHECK_HANDLE(h);
  if (h->cfg.world <= 1 || face < 0 || face > 1 || !h->tx[1][face]) return fail(h, MPMB_ERR_STATE, "no connected peer through face %d", face);
  if (h->stage != 0) return fail(h, MPMB_ERR_STATE, "migrate_send must follow resample");
  return xchg_migrate_launch(h, 1 << face, 0);
}

int mpmb_migrate_recv(MpmbHandle h, int32_t face) {
  CHECK_HANDLE(h);
  if (h->cfg.world <= 1 || face < 0 || face > 1 || !h->tx[1][face]) return fail(h, MPMB_ERR_STATE, "no connected peer through face %d", face);
  if (h->stage != 0) return fail(h, MPMB_ERR_STATE, "migrate_recv must follow resample");
  return xchg_migrate_launch(h, 0, 1 << face);
}

}  // extern "C"
