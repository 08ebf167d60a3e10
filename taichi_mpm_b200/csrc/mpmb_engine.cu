// libmpmb.so — B200-native MLS-MPM substep engine (sm_100a).  C-ABI in include/mpmb.h.
//
// Data layout in HBM:
//   * particles: SoA of seven float4 streams q0..q6 (112 B/particle), double buffered.
//       q0=(x,y,z,scalar) q1=(F0..F3) q2=(F4..F7) q3=(F8,b0,b1,b2) q4=(b3..b6) q5=(b7,b8,vx,vy)
//       q6=(vz,mass,vol,tag)   tag = group<<26 | id
//   * order: u32 key = tile<<6 | cell per particle (tile = 4x4x4 nodes, z fastest), radix-sorted
//     every substep; G2P writes its output at the sorted position, so storage order tracks the
//     sorted order and the permutation read by the next substep is near-identity (coalesced).
//   * grid: no dense grid.  P2G leaves one 6x6x6 float4 "arena" (tile + the +2 stencil halo,
//     the reference's GridCache footprint src/transfer.cpp:59-63) per active tile; G2P rebuilds
//     each node as the fixed-order sum of the <=8 arenas covering it, normalises, applies the
//     level-set boundary and keeps the result in shared memory.  No global float atomics.
#include <cuda_runtime.h>
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

constexpr uint32_t KEY_DEAD = 0xFFFFFFFFu;
constexpr uint32_t KEY_MIG_UP = 0xFFFFFFFEu;    // left through the +z face of the slab
constexpr uint32_t KEY_MIG_DOWN = 0xFFFFFFFDu;  // left through the -z face
constexpr uint32_t KEY_SPECIAL_MIN = 0xFFFFFFFDu;
constexpr int ARENA = 216;  // 6*6*6 nodes
constexpr int N_Q = 7;

struct Params {
  int res[3];
  int nnode[3];
  int nt[3];
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
  int n_alive;
  int n_tiles;
  int n_ghost;      // ghost tiles appended after the owned ones (world>1)
  int error;        // sticky device-side error flags
  int mig_count[2]; // particles packed for face 0 / 1
  int pad[2];
};
enum { DEVERR_TILE_CAPACITY = 1, DEVERR_MIGRATE_CAPACITY = 2, DEVERR_PARTICLE_CAPACITY = 4 };

struct View {  // raw pointers handed to kernels
  float4 *q[N_Q];        // current (read) buffer
  float4 *qn[N_Q];       // next (written by G2P)
  const uint32_t *keys_sorted;
  const uint32_t *perm;
  uint32_t *keys_next;
  int *tile_id, *tile_begin, *tile_end;
  int *slot_map;
  float4 *arena;
  const float4 *sdf4;
  Counters *cnt;
  int cap_tiles;
};

// ------------------------------------------------------------------------------ helpers
__device__ __forceinline__ void base_rel(float x, float inv_dx, int &base, float &rel) {
  // pos_ = p.pos * inv_delta_x (src/transfer.cpp:490); base = int(x - 0.5f) (src/kernel.h:119-121).
  // Explicit round-to-nearest ops so that no FMA contraction changes the cell assignment.
  float X = __fmul_rn(x, inv_dx);
  base = (int)__fsub_rn(X, 0.5f);
  rel = __fsub_rn(X, (float)base);
}

__device__ __forceinline__ uint32_t make_key(const Params &P, float x, float y, float z, bool &in_domain) {
  int bx, by, bz;
  float r;
  base_rel(x, P.inv_dx, bx, r);
  base_rel(y, P.inv_dx, by, r);
  base_rel(z, P.inv_dx, bz, r);
  // the 27-node stencil [base, base+2] must stay on the node grid
  in_domain = (bx >= 0) && (by >= 0) && (bz >= 0) && (bx + 2 < P.nnode[0]) && (by + 2 < P.nnode[1]) && (bz + 2 < P.nnode[2]);
  if (!in_domain) return KEY_DEAD;
  int tx = bx >> 2, ty = by >> 2, tz = bz >> 2;
  if (P.world > 1) {
    if (tz < P.tile_z0) return KEY_MIG_DOWN;
    if (tz >= P.tile_z1) return KEY_MIG_UP;
  }
  uint32_t tile = (uint32_t)((tx * P.nt[1] + ty) * P.nt[2] + tz);
  uint32_t cell = (uint32_t)(((bx & 3) << 4) | ((by & 3) << 2) | (bz & 3));
  return (tile << 6) | cell;
}

// near_boundary + abnormal (src/mpm.h:269-276, src/mpm.cpp:595-598)
__device__ __forceinline__ bool reference_deletes(const Params &P, float3 x, float3 v) {
  float X = x.x * P.inv_dx, Y = x.y * P.inv_dx, Z = x.z * P.inv_dx;
  float mn = fminf(X, fminf(Y, Z));
  float mx = fmaxf(X - (float)P.res[0], fmaxf(Y - (float)P.res[1], Z - (float)P.res[2]));
  bool bad = (mn < 7.0f) || (mx > -7.0f);
  bad |= !(isfinite(x.x) && isfinite(x.y) && isfinite(x.z) && isfinite(v.x) && isfinite(v.y) && isfinite(v.z));
  return bad;
}

// ------------------------------------------------------------------------------ upload kernels
__global__ void k_iota(uint32_t *a, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) a[i] = (uint32_t)i;
}

// Field-wise host arrays (staged on the device) -> q streams.
__global__ void k_pack_particles(View V, Params P, int n, const float *x, const float *v, const float *F, const float *b,
                                 const float *mass, const float *vol, const float *scalar, const int *group, uint32_t *keys) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float f[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, bb[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  if (F)
    for (int k = 0; k < 9; k++) f[k] = F[9 * (size_t)i + k];
  if (b)
    for (int k = 0; k < 9; k++) bb[k] = b[9 * (size_t)i + k];
  int g = group ? group[i] : 0;
  float ps;
  if (scalar) ps = scalar[i];
  else {
    int kind = P.mats[g].kind;
    ps = (kind == MAT_SNOW || kind == MAT_WATER) ? 1.0f : 0.0f;  // Jp=1 (205), j=1 (461), logJp=0 (595)
  }
  float3 xx = make_float3(x[3 * (size_t)i], x[3 * (size_t)i + 1], x[3 * (size_t)i + 2]);
  float3 vv = make_float3(v[3 * (size_t)i], v[3 * (size_t)i + 1], v[3 * (size_t)i + 2]);
  uint32_t tag = ((uint32_t)g << 26) | (uint32_t)i;
  V.q[0][i] = make_float4(xx.x, xx.y, xx.z, ps);
  V.q[1][i] = make_float4(f[0], f[1], f[2], f[3]);
  V.q[2][i] = make_float4(f[4], f[5], f[6], f[7]);
  V.q[3][i] = make_float4(f[8], bb[0], bb[1], bb[2]);
  V.q[4][i] = make_float4(bb[3], bb[4], bb[5], bb[6]);
  V.q[5][i] = make_float4(bb[7], bb[8], vv.x, vv.y);
  V.q[6][i] = make_float4(vv.z, mass[i], vol[i], __uint_as_float(tag));
  bool in_dom;
  keys[i] = make_key(P, xx.x, xx.y, xx.z, in_dom);
}

// Reference AoS slots (staged on the device) -> q streams.
__global__ void k_pack_aos(View V, Params P, int n, const unsigned char *pool, const uint32_t *indices, MpmbAosLayout L,
                           const int *group, uint32_t *keys) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned char *s = pool + (size_t)indices[i] * L.stride;
  const float *pos = (const float *)(s + L.off_pos);
  const float *vm = (const float *)(s + L.off_v_and_m);
  float f[9], bb[9];
  for (int c = 0; c < 3; c++) {
    const float *fc = (const float *)(s + L.off_dg_e + c * L.col_pitch);
    const float *bc = (const float *)(s + L.off_apic_b + c * L.col_pitch);
    for (int r = 0; r < 3; r++) {
      f[c * 3 + r] = fc[r];
      bb[c * 3 + r] = bc[r];
    }
  }
  int g = group ? group[i] : 0;
  float vol = *(const float *)(s + L.off_vol);
  float ps = L.off_scalar >= 0 ? *(const float *)(s + L.off_scalar) : 0.f;
  uint32_t tag = ((uint32_t)g << 26) | (uint32_t)i;
  V.q[0][i] = make_float4(pos[0], pos[1], pos[2], ps);
  V.q[1][i] = make_float4(f[0], f[1], f[2], f[3]);
  V.q[2][i] = make_float4(f[4], f[5], f[6], f[7]);
  V.q[3][i] = make_float4(f[8], bb[0], bb[1], bb[2]);
  V.q[4][i] = make_float4(bb[3], bb[4], bb[5], bb[6]);
  V.q[5][i] = make_float4(bb[7], bb[8], vm[0], vm[1]);
  V.q[6][i] = make_float4(vm[2], vm[3], vol, __uint_as_float(tag));
  bool in_dom;
  keys[i] = make_key(P, pos[0], pos[1], pos[2], in_dom);
}

// q streams -> field-wise arrays, compacting live particles (storage order).
__global__ void k_unpack_particles(View V, const uint32_t *keys, int n, int *count, uint32_t *id, float *x, float *v, float *F,
                                   float *b, float *mass, float *vol, float *scalar, int *group, const int *prefix) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (keys[i] >= KEY_SPECIAL_MIN) return;
  int o = prefix[i];
  float4 q0 = V.q[0][i], q1 = V.q[1][i], q2 = V.q[2][i], q3 = V.q[3][i], q4 = V.q[4][i], q5 = V.q[5][i], q6 = V.q[6][i];
  uint32_t tag = __float_as_uint(q6.w);
  if (id) id[o] = tag & 0x3FFFFFFu;
  if (group) group[o] = (int)(tag >> 26);
  if (x) { x[3 * (size_t)o] = q0.x; x[3 * (size_t)o + 1] = q0.y; x[3 * (size_t)o + 2] = q0.z; }
  if (scalar) scalar[o] = q0.w;
  if (F) {
    float *f = F + 9 * (size_t)o;
    f[0] = q1.x; f[1] = q1.y; f[2] = q1.z; f[3] = q1.w; f[4] = q2.x; f[5] = q2.y; f[6] = q2.z; f[7] = q2.w; f[8] = q3.x;
  }
  if (b) {
    float *p = b + 9 * (size_t)o;
    p[0] = q3.y; p[1] = q3.z; p[2] = q3.w; p[3] = q4.x; p[4] = q4.y; p[5] = q4.z; p[6] = q4.w; p[7] = q5.x; p[8] = q5.y;
  }
  if (v) { v[3 * (size_t)o] = q5.z; v[3 * (size_t)o + 1] = q5.w; v[3 * (size_t)o + 2] = q6.x; }
  if (mass) mass[o] = q6.y;
  if (vol) vol[o] = q6.z;
  (void)count;
}

__global__ void k_alive_flags(const uint32_t *keys, int n, int *flags) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) flags[i] = keys[i] < KEY_SPECIAL_MIN ? 1 : 0;
}

// ------------------------------------------------------------------------------ tile list
__global__ void k_clear_tiles(View V) {
  int n = V.cnt->n_tiles + V.cnt->n_ghost;
  for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < n; s += gridDim.x * blockDim.x) V.slot_map[V.tile_id[s]] = -1;
}
__global__ void k_reset_counters(Counters *c) {
  c->n_tiles = 0;
  c->n_ghost = 0;
  c->n_alive = 0;
  c->mig_count[0] = 0;
  c->mig_count[1] = 0;
}

// Active-tile list = run heads of the sorted keys.  Replaces page_map / block_meta construction
// (src/mpm.cpp:817-826,876-889) — on the device, no host page map.
__global__ void k_build_tiles(View V, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t k = V.keys_sorted[i];
  if (k >= KEY_SPECIAL_MIN) {
    if (i == 0 || V.keys_sorted[i - 1] < KEY_SPECIAL_MIN) V.cnt->n_alive = i;
    return;
  }
  if (i == n - 1) V.cnt->n_alive = n;
  uint32_t t = k >> 6;
  if (i > 0 && (V.keys_sorted[i - 1] >> 6) == t) return;
  int slot = atomicAdd(&V.cnt->n_tiles, 1);
  if (slot >= V.cap_tiles) {
    atomicOr(&V.cnt->error, DEVERR_TILE_CAPACITY);
    return;
  }
  // end of the run: first index whose tile differs (specials sort last)
  int lo = i + 1, hi = n;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if ((V.keys_sorted[mid] >> 6) == t && V.keys_sorted[mid] < KEY_SPECIAL_MIN) lo = mid + 1;
    else hi = mid;
  }
  V.tile_id[slot] = (int)t;
  V.tile_begin[slot] = i;
  V.tile_end[slot] = lo;
  V.slot_map[t] = slot;
}

// ------------------------------------------------------------------------------ P2G
// Replaces MPM<3>::rasterize_optimized / block_op_normal (src/transfer.cpp:467-569).
// One CTA per active tile (persistent round-robin), one thread per particle of the tile's run.
// Contributions are accumulated into a shared-memory 6x6x6 arena; the node visiting order of each
// lane is rotated by its lane index so that the particles of one cell (adjacent lanes after the
// sort) hit 27 different nodes at any instant instead of one.
template <int BLOCK>
__global__ void __launch_bounds__(BLOCK) k_p2g(View V, Params P) {
  __shared__ float s_arena[4][ARENA + 8];
  const int tid = threadIdx.x;
  const int n_tiles = V.cnt->n_tiles;
  const float S = -4.0f * P.inv_dx * P.dt;  // src/transfer.cpp:465
  for (int slot = blockIdx.x; slot < n_tiles; slot += gridDim.x) {
    for (int n = tid; n < ARENA; n += BLOCK) {
      s_arena[0][n] = 0.f; s_arena[1][n] = 0.f; s_arena[2][n] = 0.f; s_arena[3][n] = 0.f;
    }
    __syncthreads();
    const int begin = V.tile_begin[slot], end = V.tile_end[slot];
    const int tile = V.tile_id[slot];
    const int tz = tile % P.nt[2], ty = (tile / P.nt[2]) % P.nt[1], tx = tile / (P.nt[2] * P.nt[1]);
    for (int j = begin + tid; j < end; j += BLOCK) {
      const uint32_t p = V.perm[j];
      const float4 q0 = V.q[0][p], q1 = V.q[1][p], q2 = V.q[2][p], q3 = V.q[3][p], q4 = V.q[4][p], q5 = V.q[5][p], q6 = V.q[6][p];
      const float mass = q6.y, vol = q6.z;
      const uint32_t tag = __float_as_uint(q6.w);
      const Material &mat = P.mats[tag >> 26];
      float3 v = make_float3(q5.z, q5.w, q6.x);
      if (P.particle_gravity) {  // src/transfer.cpp:485-487
        v.x += P.gdt[0]; v.y += P.gdt[1]; v.z += P.gdt[2];
      }
      int bx, by, bz;
      float rx, ry, rz;
      base_rel(q0.x, P.inv_dx, bx, rx);
      base_rel(q0.y, P.inv_dx, by, ry);
      base_rel(q0.z, P.inv_dx, bz, rz);
      bx -= tx * 4; by -= ty * 4; bz -= tz * 4;
      float wx[3], wy[3], wz[3];
      bspline_weights(rx, wx);
      bspline_weights(ry, wy);
      bspline_weights(rz, wz);
      Mat3 F;
      F.m[0] = q1.x; F.m[1] = q1.y; F.m[2] = q1.z; F.m[3] = q1.w; F.m[4] = q2.x; F.m[5] = q2.y; F.m[6] = q2.z; F.m[7] = q2.w; F.m[8] = q3.x;
      Mat3 A;  // affine = stress * S + apic_b * (4 m)   (src/transfer.cpp:503,521-522)
      calculate_force(mat, F, q0.w, vol, A);
      const float bm = 4.0f * mass;
      A.m[0] = fmaf(A.m[0], S, q3.y * bm); A.m[1] = fmaf(A.m[1], S, q3.z * bm); A.m[2] = fmaf(A.m[2], S, q3.w * bm);
      A.m[3] = fmaf(A.m[3], S, q4.x * bm); A.m[4] = fmaf(A.m[4], S, q4.y * bm); A.m[5] = fmaf(A.m[5], S, q4.z * bm);
      A.m[6] = fmaf(A.m[6], S, q4.w * bm); A.m[7] = fmaf(A.m[7], S, q5.x * bm); A.m[8] = fmaf(A.m[8], S, q5.y * bm);
      const float mvx = mass * v.x, mvy = mass * v.y, mvz = mass * v.z;
      // lane-dependent rotation of the stencil visiting order
      const int lane = tid & 31;
      const int sx = lane % 3, sy = (lane / 3) % 3, sz = (lane / 9) % 3;
      float wxr[3], wyr[3], wzr[3];
      int nxr[3], nyr[3], nzr[3];
#pragma unroll
      for (int i = 0; i < 3; i++) {
        int ix = i + sx; ix -= (ix >= 3) ? 3 : 0;
        int iy = i + sy; iy -= (iy >= 3) ? 3 : 0;
        int iz = i + sz; iz -= (iz >= 3) ? 3 : 0;
        nxr[i] = ix; nyr[i] = iy; nzr[i] = iz;
        wxr[i] = ix == 0 ? wx[0] : (ix == 1 ? wx[1] : wx[2]);
        wyr[i] = iy == 0 ? wy[0] : (iy == 1 ? wy[1] : wy[2]);
        wzr[i] = iz == 0 ? wz[0] : (iz == 1 ? wz[1] : wz[2]);
      }
#pragma unroll
      for (int i = 0; i < 3; i++) {
        const float d0 = rx - (float)nxr[i];  // particle - node, grid units (src/transfer.cpp:528)
        const float ax = fmaf(A.m[0], d0, mvx), ay = fmaf(A.m[1], d0, mvy), az = fmaf(A.m[2], d0, mvz);
#pragma unroll
        for (int jn = 0; jn < 3; jn++) {
          const float d1 = ry - (float)nyr[jn];
          const float bxv = fmaf(A.m[3], d1, ax), byv = fmaf(A.m[4], d1, ay), bzv = fmaf(A.m[5], d1, az);
          const float wij = wxr[i] * wyr[jn];
          const int row = ((bx + nxr[i]) * 6 + (by + nyr[jn])) * 6 + bz;
#pragma unroll
          for (int k = 0; k < 3; k++) {
            const float d2 = rz - (float)nzr[k];
            const float w = wij * wzr[k];
            const int node = row + nzr[k];
            atomicAdd(&s_arena[0][node], w * fmaf(A.m[6], d2, bxv));
            atomicAdd(&s_arena[1][node], w * fmaf(A.m[7], d2, byv));
            atomicAdd(&s_arena[2][node], w * fmaf(A.m[8], d2, bzv));
            atomicAdd(&s_arena[3][node], w * mass);
          }
        }
      }
    }
    __syncthreads();
    float4 *out = V.arena + (size_t)slot * ARENA;
    for (int n = tid; n < ARENA; n += BLOCK) out[n] = make_float4(s_arena[0][n], s_arena[1][n], s_arena[2][n], s_arena[3][n]);
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------ grid node
// Momentum/mass of global node g = fixed-order sum of the arenas that cover it:
// owner tile T=(g>>2) holds it at local l=g&3; tile T-o (o in {0,1}^3) holds it at l+4o (needs l<=1).
template <class SlotOf>
__device__ __forceinline__ float4 gather_node(const float4 *arena, int lx, int ly, int lz, SlotOf slot_of) {
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int ox = 0; ox < 2; ox++) {
    if (ox && lx > 1) continue;
#pragma unroll
    for (int oy = 0; oy < 2; oy++) {
      if (oy && ly > 1) continue;
#pragma unroll
      for (int oz = 0; oz < 2; oz++) {
        if (oz && lz > 1) continue;
        int slot = slot_of(ox, oy, oz);
        if (slot < 0) continue;
        float4 a = arena[(size_t)slot * ARENA + ((lx + 4 * ox) * 6 + (ly + 4 * oy)) * 6 + (lz + 4 * oz)];
        acc.x += a.x; acc.y += a.y; acc.z += a.z; acc.w += a.w;
      }
    }
  }
  return acc;
}

// normalize_grid_and_apply_external_force (src/mpm.cpp:277-294) + apply_grid_boundary_conditions
// (src/mpm.cpp:296-372, static level set) for one node.
__device__ __forceinline__ float4 node_update(const Params &P, const float4 *sdf4, float4 g, int gx, int gy, int gz) {
  float m = g.w;
  if (m > 0.f) {
    float inv = 1.0f / m;
    float ix = P.particle_gravity ? 0.f : P.gdt[0], iy = P.particle_gravity ? 0.f : P.gdt[1], iz = P.particle_gravity ? 0.f : P.gdt[2];
    g.x = fmaf(g.x, inv, ix);
    g.y = fmaf(g.y, inv, iy);
    g.z = fmaf(g.z, inv, iz);
  }
  if (P.has_sdf && m != 0.f && gx < P.nnode[0] && gy < P.nnode[1] && gz < P.nnode[2]) {
    float4 s = sdf4[((size_t)gx * P.nnode[1] + gy) * P.nnode[2] + gz];
    if (!(s.w < -3.0f || 0.0f < s.w)) {
      float3 v = friction_project0(make_float3(g.x, g.y, g.z), make_float3(s.x, s.y, s.z), P.friction);
      g.x = v.x; g.y = v.y; g.z = v.z;
    }
  }
  return g;
}

// ------------------------------------------------------------------------------ G2P
// Replaces normalize_grid_and_apply_external_force + apply_grid_boundary_conditions +
// MPM<3>::resample_optimized / block_op_normal (src/transfer.cpp:837-954) + Particle::plasticity +
// clear_boundary_particles (src/mpm.cpp:583-633), and emits the next substep's sort key.
template <int BLOCK>
__global__ void __launch_bounds__(BLOCK) k_g2p(View V, Params P) {
  __shared__ float4 s_vel[ARENA];
  __shared__ int s_nb[27];
  const int tid = threadIdx.x;
  const int n_tiles = V.cnt->n_tiles;
  const float scale = -4.0f * P.inv_dx * P.dt;  // src/transfer.cpp:938
  for (int slot = blockIdx.x; slot < n_tiles; slot += gridDim.x) {
    const int tile = V.tile_id[slot];
    const int tz = tile % P.nt[2], ty = (tile / P.nt[2]) % P.nt[1], tx = tile / (P.nt[2] * P.nt[1]);
    if (tid < 27) {
      int ox = tid / 9 - 1, oy = (tid / 3) % 3 - 1, oz = tid % 3 - 1;
      int x = tx + ox, y = ty + oy, z = tz + oz;
      int s = -1;
      if (x >= 0 && y >= 0 && z >= 0 && x < P.nt[0] && y < P.nt[1] && z < P.nt[2]) s = V.slot_map[(x * P.nt[1] + y) * P.nt[2] + z];
      s_nb[tid] = s;
    }
    __syncthreads();
    for (int n = tid; n < ARENA; n += BLOCK) {
      int a = n / 36, b = (n / 6) % 6, c = n % 6;
      int wx_ = a >> 2, wy_ = b >> 2, wz_ = c >> 2;  // owner tile offset (0/1)
      int lx = a & 3, ly = b & 3, lz = c & 3;
      float4 g = gather_node(V.arena, lx, ly, lz, [&](int ox, int oy, int oz) {
        return s_nb[(wx_ - ox + 1) * 9 + (wy_ - oy + 1) * 3 + (wz_ - oz + 1)];
      });
      s_vel[n] = node_update(P, V.sdf4, g, tx * 4 + a, ty * 4 + b, tz * 4 + c);
    }
    __syncthreads();
    const int begin = V.tile_begin[slot], end = V.tile_end[slot];
    for (int j = begin + tid; j < end; j += BLOCK) {
      const uint32_t p = V.perm[j];
      const float4 q0 = V.q[0][p], q1 = V.q[1][p], q2 = V.q[2][p], q3 = V.q[3][p], q6 = V.q[6][p];
      const uint32_t tag = __float_as_uint(q6.w);
      const Material &mat = P.mats[tag >> 26];
      int bx, by, bz;
      float rx, ry, rz;
      base_rel(q0.x, P.inv_dx, bx, rx);
      base_rel(q0.y, P.inv_dx, by, ry);
      base_rel(q0.z, P.inv_dx, bz, rz);
      bx -= tx * 4; by -= ty * 4; bz -= tz * 4;
      float wx[3], wy[3], wz[3];
      bspline_weights(rx, wx);
      bspline_weights(ry, wy);
      bspline_weights(rz, wz);
      float3 v = make_float3(0.f, 0.f, 0.f);
      Mat3 B;
#pragma unroll
      for (int k = 0; k < 9; k++) B.m[k] = 0.f;
#pragma unroll
      for (int i = 0; i < 3; i++)
#pragma unroll
        for (int jn = 0; jn < 3; jn++) {
          const float wij = wx[i] * wy[jn];
          const int row = ((bx + i) * 6 + (by + jn)) * 6 + bz;
#pragma unroll
          for (int k = 0; k < 3; k++) {
            const float w = wij * wz[k];
            const float4 g = s_vel[row + k];
            const float d0 = rx - (float)i, d1 = ry - (float)jn, d2 = rz - (float)k;
            v.x = fmaf(g.x, w, v.x); v.y = fmaf(g.y, w, v.y); v.z = fmaf(g.z, w, v.z);
            const float wgx = w * g.x, wgy = w * g.y, wgz = w * g.z;  // b_[r] += w v_i d_r (900-903)
            B.m[0] = fmaf(wgx, d0, B.m[0]); B.m[1] = fmaf(wgy, d0, B.m[1]); B.m[2] = fmaf(wgz, d0, B.m[2]);
            B.m[3] = fmaf(wgx, d1, B.m[3]); B.m[4] = fmaf(wgy, d1, B.m[4]); B.m[5] = fmaf(wgz, d1, B.m[5]);
            B.m[6] = fmaf(wgx, d2, B.m[6]); B.m[7] = fmaf(wgy, d2, B.m[7]); B.m[8] = fmaf(wgz, d2, B.m[8]);
          }
        }
      Mat3 cdg;  // cdg = I + (-4 inv_dx dt) b   (src/transfer.cpp:938-942)
#pragma unroll
      for (int k = 0; k < 9; k++) cdg.m[k] = fmaf(scale, B.m[k], (k % 4 == 0) ? 1.f : 0.f);
      Mat3 F;
      F.m[0] = q1.x; F.m[1] = q1.y; F.m[2] = q1.z; F.m[3] = q1.w; F.m[4] = q2.x; F.m[5] = q2.y; F.m[6] = q2.z; F.m[7] = q2.w; F.m[8] = q3.x;
      float ps = q0.w;
      plasticity(mat, cdg, F, ps);
      float3 x = make_float3(fmaf(v.x, P.dt, q0.x), fmaf(v.y, P.dt, q0.y), fmaf(v.z, P.dt, q0.z));  // 951
      bool in_dom;
      uint32_t key = make_key(P, x.x, x.y, x.z, in_dom);
      if (P.clean_boundary && reference_deletes(P, x, v)) key = KEY_DEAD;
      if (!(isfinite(x.x) && isfinite(x.y) && isfinite(x.z))) key = KEY_DEAD;
      // write at the sorted position j: storage order follows the sort
      V.qn[0][j] = make_float4(x.x, x.y, x.z, ps);
      V.qn[1][j] = make_float4(F.m[0], F.m[1], F.m[2], F.m[3]);
      V.qn[2][j] = make_float4(F.m[4], F.m[5], F.m[6], F.m[7]);
      V.qn[3][j] = make_float4(F.m[8], B.m[0], B.m[1], B.m[2]);
      V.qn[4][j] = make_float4(B.m[3], B.m[4], B.m[5], B.m[6]);
      V.qn[5][j] = make_float4(B.m[7], B.m[8], v.x, v.y);
      V.qn[6][j] = make_float4(v.z, q6.y, q6.z, q6.w);
      V.keys_next[j] = key;
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
    float4 g = gather_node(V.arena, gx & 3, gy & 3, gz & 3, [&](int ox, int oy, int oz) {
      int x = tx - ox, y = ty - oy, z = tz - oz;
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

  int64_t cap = 0;       // particle slots allocated
  int n_bound = 0;       // slots that may hold live particles (host upper bound)
  int cur = 0;           // which q buffer is current
  float4 *q[2][N_Q] = {};
  uint32_t *keys[2] = {};       // keys in storage order (cur / next)
  uint32_t *keys_sorted = nullptr, *perm = nullptr, *iota = nullptr;
  void *cub_temp = nullptr;
  size_t cub_bytes = 0;
  int key_bits = 32;

  int cap_tiles = 0;
  int *tile_id = nullptr, *tile_begin = nullptr, *tile_end = nullptr, *slot_map = nullptr;
  float4 *arena = nullptr;
  float4 *sdf4 = nullptr;
  Counters *cnt = nullptr;

  int stage = 0;  // 0 idle/after resample, 1 after sort, 2 after rasterize
  int num_sms = 148;
  int64_t launches = 0;

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

static View make_view(MpmbEngine *h) {
  View V{};
  for (int k = 0; k < N_Q; k++) {
    V.q[k] = h->q[h->cur][k];
    V.qn[k] = h->q[h->cur ^ 1][k];
  }
  V.keys_sorted = h->keys_sorted;
  V.perm = h->perm;
  V.keys_next = h->keys[h->cur ^ 1];
  V.tile_id = h->tile_id;
  V.tile_begin = h->tile_begin;
  V.tile_end = h->tile_end;
  V.slot_map = h->slot_map;
  V.arena = h->arena;
  V.sdf4 = h->sdf4;
  V.cnt = h->cnt;
  V.cap_tiles = h->cap_tiles;
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
  for (int b = 0; b < 2; b++) {
    for (int k = 0; k < N_Q; k++) { cudaFree(h->q[b][k]); h->q[b][k] = nullptr; }
    cudaFree(h->keys[b]); h->keys[b] = nullptr;
  }
  cudaFree(h->keys_sorted); cudaFree(h->perm); cudaFree(h->iota); cudaFree(h->cub_temp);
  h->keys_sorted = h->perm = h->iota = nullptr;
  h->cub_temp = nullptr;
  h->cap = 0;
  return 0;
}

static int alloc_particles(MpmbEngine *h, int64_t cap) {
  free_particles(h);
  if (cap >= (1ll << 26)) return fail(h, MPMB_ERR_CAPACITY, "capacity %lld exceeds 2^26 particles per GPU", (long long)cap);
  for (int b = 0; b < 2; b++) {
    for (int k = 0; k < N_Q; k++) CUDA_TRY(h, cudaMalloc(&h->q[b][k], sizeof(float4) * cap));
    CUDA_TRY(h, cudaMalloc(&h->keys[b], sizeof(uint32_t) * cap));
    CUDA_TRY(h, cudaMemsetAsync(h->keys[b], 0xFF, sizeof(uint32_t) * cap, h->stream));
  }
  CUDA_TRY(h, cudaMalloc(&h->keys_sorted, sizeof(uint32_t) * cap));
  CUDA_TRY(h, cudaMalloc(&h->perm, sizeof(uint32_t) * cap));
  CUDA_TRY(h, cudaMalloc(&h->iota, sizeof(uint32_t) * cap));
  k_iota<<<(unsigned)((cap + 255) / 256), 256, 0, h->stream>>>(h->iota, (int)cap);
  h->cub_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, h->cub_bytes, h->keys[0], h->keys_sorted, h->iota, h->perm, (int)cap, 0, 32, h->stream);
  CUDA_TRY(h, cudaMalloc(&h->cub_temp, h->cub_bytes));
  h->cap = cap;
  return MPMB_OK;
}

extern "C" {

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
    ntot *= (size_t)P.nt[d];
  }
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
  h->key_bits = 6;
  while ((1ull << (h->key_bits - 6)) < ntot) h->key_bits++;
  if (h->key_bits > 31) { delete h; return fail(nullptr, MPMB_ERR_INVALID, "tile grid too large for 32-bit keys"); }
  h->key_bits = 32;  // special keys (dead / migrating) use the top of the range
  if (cudaSetDevice(cfg->device) != cudaSuccess) { delete h; return fail(nullptr, MPMB_ERR_CUDA, "cudaSetDevice failed"); }
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, cfg->device);
  h->num_sms = prop.multiProcessorCount;
  // tiles: every tile of the (slab of the) domain can be active
  int64_t slab_layers = (h->cfg.world > 1) ? (int64_t)(P.tile_z1 - P.tile_z0) + 2 : P.nt[2];
  int64_t cap_tiles = (int64_t)P.nt[0] * P.nt[1] * slab_layers;
  h->cap_tiles = (int)cap_tiles;
  auto bail = [&](const char *what) {
    std::string msg = std::string(what) + ": " + cudaGetErrorString(cudaGetLastError());
    mpmb_destroy(h);
    return fail(nullptr, MPMB_ERR_CUDA, "%s", msg.c_str());
  };
  if (cudaMalloc(&h->tile_id, sizeof(int) * cap_tiles) != cudaSuccess) return bail("cudaMalloc tile_id");
  if (cudaMalloc(&h->tile_begin, sizeof(int) * cap_tiles) != cudaSuccess) return bail("cudaMalloc tile_begin");
  if (cudaMalloc(&h->tile_end, sizeof(int) * cap_tiles) != cudaSuccess) return bail("cudaMalloc tile_end");
  if (cudaMalloc(&h->slot_map, sizeof(int) * ntot) != cudaSuccess) return bail("cudaMalloc slot_map");
  if (cudaMalloc(&h->arena, sizeof(float4) * ARENA * cap_tiles) != cudaSuccess) return bail("cudaMalloc arena");
  if (cudaMalloc(&h->cnt, sizeof(Counters)) != cudaSuccess) return bail("cudaMalloc counters");
  cudaMemset(h->slot_map, 0xFF, sizeof(int) * ntot);
  cudaMemset(h->cnt, 0, sizeof(Counters));
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
  cudaFree(h->tile_id); cudaFree(h->tile_begin); cudaFree(h->tile_end); cudaFree(h->slot_map);
  cudaFree(h->arena); cudaFree(h->sdf4); cudaFree(h->cnt);
  for (auto &e : h->events) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); }
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
  return MPMB_OK;
}

int mpmb_set_material(MpmbHandle h, int32_t group, int32_t kind, const float *params, int32_t n_params) {
  CHECK_HANDLE(h);
  if (group < 0 || group >= MPMB_MAX_GROUPS) return fail(h, MPMB_ERR_INVALID, "group %d out of range", group);
  if (kind < MPMB_MAT_LINEAR || kind > MPMB_MAT_SAND) return fail(h, MPMB_ERR_INVALID, "unknown material kind %d", kind);
  if (n_params < 0 || n_params > MPMB_MAT_PARAMS || (n_params > 0 && !params)) return fail(h, MPMB_ERR_INVALID, "bad parameter vector");
  h->P.mats[group].kind = kind;
  for (int k = 0; k < 8; k++) h->P.mats[group].p[k] = k < n_params ? params[k] : 0.f;
  return MPMB_OK;
}

int mpmb_set_sdf(MpmbHandle h, const float *sdf4, float friction) {
  CHECK_HANDLE(h);
  size_t n = (size_t)h->P.nnode[0] * h->P.nnode[1] * h->P.nnode[2];
  if (!sdf4) {
    h->P.has_sdf = 0;
    return MPMB_OK;
  }
  if (!h->sdf4) CUDA_TRY(h, cudaMalloc(&h->sdf4, sizeof(float4) * n));
  CUDA_TRY(h, cudaMemcpyAsync(h->sdf4, sdf4, sizeof(float4) * n, cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  h->P.has_sdf = 1;
  h->P.friction = friction;
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
  return MPMB_OK;
}

static int ensure_capacity(MpmbEngine *h, int64_t n) {
  int64_t want = n;
  if (h->cfg.world > 1) want = n + 4 * (h->cfg.migrate_capacity > 0 ? h->cfg.migrate_capacity : 0) + n / 4;
  if (h->cfg.capacity > 0) {
    if (n > h->cap) return fail(h, MPMB_ERR_CAPACITY, "%lld particles exceed the configured capacity %lld", (long long)n, (long long)h->cap);
    return MPMB_OK;
  }
  if (want > h->cap) return alloc_particles(h, want);
  return MPMB_OK;
}

static int finish_upload(MpmbEngine *h, int64_t n) {
  // slots beyond n hold dead keys
  if (h->cap > n) CUDA_TRY(h, cudaMemsetAsync(h->keys[h->cur] + n, 0xFF, sizeof(uint32_t) * (h->cap - n), h->stream));
  h->n_bound = (int)n;
  h->stage = 0;
  CUDA_TRY(h, cudaMemsetAsync(&h->cnt->n_alive, 0, sizeof(int), h->stream));
  int nn = (int)n;
  CUDA_TRY(h, cudaMemcpyAsync(&h->cnt->n_alive, &nn, sizeof(int), cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  return MPMB_OK;
}

int mpmb_upload_particles(MpmbHandle h, int64_t n, const float *x, const float *v, const float *F, const float *b,
                          const float *mass, const float *vol, const float *scalar, const int32_t *group) {
  CHECK_HANDLE(h);
  if (n < 0 || (n > 0 && (!x || !v || !mass || !vol))) return fail(h, MPMB_ERR_INVALID, "x, v, mass, vol are required");
  int rc = ensure_capacity(h, n);
  if (rc != MPMB_OK) return rc;
  if (n == 0) return finish_upload(h, 0);
  // stage the field arrays on the device (one allocation)
  size_t fl = (size_t)n * (3 + 3 + (F ? 9 : 0) + (b ? 9 : 0) + 1 + 1 + (scalar ? 1 : 0)) + (group ? (size_t)n : 0);
  float *stage = nullptr;
  CUDA_TRY(h, cudaMalloc(&stage, fl * sizeof(float)));
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
  k_pack_particles<<<(unsigned)((n + 127) / 128), 128, 0, h->stream>>>(V, h->P, (int)n, dx_, dv, dF, db, dm, dvol, ds, dg, h->keys[h->cur]);
  h->launches++;
  cudaError_t e = cudaStreamSynchronize(h->stream);
  cudaFree(stage);
  CUDA_TRY(h, e);
  CUDA_TRY(h, cudaGetLastError());
  return finish_upload(h, n);
}

int mpmb_upload_aos(MpmbHandle h, int64_t n, const void *pool, int64_t pool_slots, const uint32_t *indices,
                    const MpmbAosLayout *L, const int32_t *group) {
  CHECK_HANDLE(h);
  if (n < 0 || !pool || !indices || !L || L->stride <= 0) return fail(h, MPMB_ERR_INVALID, "pool, indices and layout are required");
  int rc = ensure_capacity(h, n);
  if (rc != MPMB_OK) return rc;
  if (n == 0) return finish_upload(h, 0);
  unsigned char *d_pool = nullptr;
  uint32_t *d_idx = nullptr;
  int *d_grp = nullptr;
  CUDA_TRY(h, cudaMalloc(&d_pool, (size_t)pool_slots * L->stride));
  CUDA_TRY(h, cudaMalloc(&d_idx, sizeof(uint32_t) * n));
  if (group) CUDA_TRY(h, cudaMalloc(&d_grp, sizeof(int) * n));
  cudaMemcpyAsync(d_pool, pool, (size_t)pool_slots * L->stride, cudaMemcpyHostToDevice, h->stream);
  cudaMemcpyAsync(d_idx, indices, sizeof(uint32_t) * n, cudaMemcpyHostToDevice, h->stream);
  if (group) cudaMemcpyAsync(d_grp, group, sizeof(int) * n, cudaMemcpyHostToDevice, h->stream);
  View V = make_view(h);
  k_pack_aos<<<(unsigned)((n + 127) / 128), 128, 0, h->stream>>>(V, h->P, (int)n, d_pool, d_idx, *L, d_grp, h->keys[h->cur]);
  h->launches++;
  cudaError_t e = cudaStreamSynchronize(h->stream);
  cudaFree(d_pool); cudaFree(d_idx); cudaFree(d_grp);
  CUDA_TRY(h, e);
  CUDA_TRY(h, cudaGetLastError());
  return finish_upload(h, n);
}

int mpmb_num_particles(MpmbHandle h, int64_t *n) {
  CHECK_HANDLE(h);
  if (!n) return fail(h, MPMB_ERR_INVALID, "null argument");
  // live = storage-order keys that are not special
  int rc = mpmb_synchronize(h);
  if (rc != MPMB_OK) return rc;
  if (h->n_bound == 0) { *n = 0; return MPMB_OK; }
  std::vector<uint32_t> keys(h->n_bound);
  CUDA_TRY(h, cudaMemcpy(keys.data(), h->keys[h->cur], sizeof(uint32_t) * h->n_bound, cudaMemcpyDeviceToHost));
  int64_t c = 0;
  for (uint32_t k : keys) c += (k < KEY_SPECIAL_MIN);
  *n = c;
  return MPMB_OK;
}

int mpmb_download_particles(MpmbHandle h, int64_t cap, int64_t *n_out, uint32_t *id, float *x, float *v, float *F, float *b,
                            float *mass, float *vol, float *scalar, int32_t *group) {
  CHECK_HANDLE(h);
  if (!n_out) return fail(h, MPMB_ERR_INVALID, "n_out is required");
  int rc = mpmb_synchronize(h);
  if (rc != MPMB_OK) return rc;
  const int n = h->n_bound;
  *n_out = 0;
  if (n == 0) return MPMB_OK;
  // exclusive prefix of the alive flags (CUB scan)
  int *flags = nullptr, *prefix = nullptr;
  CUDA_TRY(h, cudaMalloc(&flags, sizeof(int) * n));
  CUDA_TRY(h, cudaMalloc(&prefix, sizeof(int) * (n + 1)));
  k_alive_flags<<<(n + 255) / 256, 256, 0, h->stream>>>(h->keys[h->cur], n, flags);
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
  if (fl) CUDA_TRY(h, cudaMalloc(&stage, fl * sizeof(float)));
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
  k_unpack_particles<<<(n + 127) / 128, 128, 0, h->stream>>>(V, h->keys[h->cur], n, nullptr, did, dx_, dv, dF, db, dm, dvol, ds, dg, prefix);
  h->launches += 2;
  auto get = [&](void *dst, const void *src, size_t per) {
    if (dst) cudaMemcpyAsync(dst, src, per * alive * sizeof(float), cudaMemcpyDeviceToHost, h->stream);
  };
  get(x, dx_, 3); get(v, dv, 3); get(F, dF, 9); get(b, db, 9); get(mass, dm, 1); get(vol, dvol, 1); get(scalar, ds, 1);
  get(id, did, 1); get(group, dg, 1);
  cudaError_t e = cudaStreamSynchronize(h->stream);
  cudaFree(flags); cudaFree(prefix); cudaFree(tmp); cudaFree(stage);
  CUDA_TRY(h, e);
  CUDA_TRY(h, cudaGetLastError());
  *n_out = alive;
  return MPMB_OK;
}

int mpmb_download_aos(MpmbHandle h, void *pool, int64_t pool_slots, uint32_t *indices, int64_t n_indices, const MpmbAosLayout *L,
                      int64_t *n_alive) {
  CHECK_HANDLE(h);
  if (!pool || !indices || !L || !n_alive) return fail(h, MPMB_ERR_INVALID, "null argument");
  int64_t cap = h->n_bound, n = 0;
  std::vector<uint32_t> id(cap);
  std::vector<float> x(3 * cap), v(3 * cap), F(9 * cap), b(9 * cap), mass(cap), vol(cap), sc(cap);
  int rc = mpmb_download_particles(h, cap, &n, id.data(), x.data(), v.data(), F.data(), b.data(), mass.data(), vol.data(), sc.data(), nullptr);
  if (rc != MPMB_OK) return rc;
  std::vector<uint32_t> survivors;
  survivors.reserve(n);
  std::vector<std::pair<uint32_t, int64_t>> by_id(n);
  for (int64_t k = 0; k < n; k++) by_id[k] = {id[k], k};
  std::sort(by_id.begin(), by_id.end());
  for (auto &pr : by_id) {
    if ((int64_t)pr.first >= n_indices) return fail(h, MPMB_ERR_INVALID, "particle id %u outside the index vector", pr.first);
    uint32_t slot = indices[pr.first];
    if ((int64_t)slot >= pool_slots) return fail(h, MPMB_ERR_INVALID, "slot %u outside the pool", slot);
    int64_t k = pr.second;
    unsigned char *s = (unsigned char *)pool + (size_t)slot * L->stride;
    float *pos = (float *)(s + L->off_pos), *vm = (float *)(s + L->off_v_and_m);
    for (int d = 0; d < 3; d++) { pos[d] = x[3 * k + d]; vm[d] = v[3 * k + d]; }
    vm[3] = mass[k];
    for (int c = 0; c < 3; c++) {
      float *fc = (float *)(s + L->off_dg_e + c * L->col_pitch), *bc = (float *)(s + L->off_apic_b + c * L->col_pitch);
      for (int r = 0; r < 3; r++) { fc[r] = F[9 * k + c * 3 + r]; bc[r] = b[9 * k + c * 3 + r]; }
    }
    if (L->off_scalar >= 0) *(float *)(s + L->off_scalar) = sc[k];
    survivors.push_back(slot);
  }
  for (size_t k = 0; k < survivors.size(); k++) indices[k] = survivors[k];
  *n_alive = n;
  return MPMB_OK;
}

// ------------------------------------------------------------------------------ stages
int mpmb_sort_particles_and_populate_grid(MpmbHandle h) {
  CHECK_HANDLE(h);
  if (h->stage != 0) return fail(h, MPMB_ERR_STATE, "sort must follow resample/upload");
  if (h->cap == 0) return fail(h, MPMB_ERR_STATE, "no particles uploaded");
  prof_begin(h, 0);
  View V = make_view(h);
  k_clear_tiles<<<64, 256, 0, h->stream>>>(V);
  k_reset_counters<<<1, 1, 0, h->stream>>>(h->cnt);
  int n = h->n_bound;
  if (n > 0) {
    cub::DeviceRadixSort::SortPairs(h->cub_temp, h->cub_bytes, h->keys[h->cur], h->keys_sorted, h->iota, h->perm, n, 0, h->key_bits, h->stream);
    k_build_tiles<<<(n + 255) / 256, 256, 0, h->stream>>>(V, n);
  }
  h->launches += 3;
  prof_end(h, 3);
  CUDA_TRY(h, cudaGetLastError());
  h->stage = 1;
  return MPMB_OK;
}

int mpmb_rasterize(MpmbHandle h) {
  CHECK_HANDLE(h);
  if (h->stage != 1) return fail(h, MPMB_ERR_STATE, "rasterize must follow sort_particles_and_populate_grid");
  prof_begin(h, 1);
  View V = make_view(h);
  k_p2g<128><<<h->num_sms * 8, 128, 0, h->stream>>>(V, h->P);
  h->launches += 1;
  prof_end(h, 1);
  CUDA_TRY(h, cudaGetLastError());
  h->stage = 2;
  return MPMB_OK;
}

int mpmb_resample(MpmbHandle h) {
  CHECK_HANDLE(h);
  if (h->stage != 2) return fail(h, MPMB_ERR_STATE, "resample must follow rasterize");
  prof_begin(h, 2);
  View V = make_view(h);
  // slots the kernel does not write (dead tail) must carry dead keys
  cudaMemsetAsync(h->keys[h->cur ^ 1], 0xFF, sizeof(uint32_t) * h->n_bound, h->stream);
  k_g2p<128><<<h->num_sms * 8, 128, 0, h->stream>>>(V, h->P);
  h->launches += 1;
  prof_end(h, 1);
  CUDA_TRY(h, cudaGetLastError());
  h->cur ^= 1;
  h->stage = 0;
  return MPMB_OK;
}

int mpmb_substep(MpmbHandle h, int32_t nsub) {
  CHECK_HANDLE(h);
  for (int s = 0; s < nsub; s++) {
    int rc;
    if ((rc = mpmb_sort_particles_and_populate_grid(h)) != MPMB_OK) return rc;
    if ((rc = mpmb_rasterize(h)) != MPMB_OK) return rc;
    if ((rc = mpmb_resample(h)) != MPMB_OK) return rc;
  }
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

// ------------------------------------------------------------------------------ multi-GPU (not built yet)
int64_t mpmb_halo_bytes(MpmbHandle h) { (void)h; return 0; }
int mpmb_halo_pack(MpmbHandle h, int32_t, void *) { return fail(h, MPMB_ERR_STATE, "z-slab exchange not implemented yet"); }
int mpmb_halo_unpack(MpmbHandle h, int32_t, const void *) { return fail(h, MPMB_ERR_STATE, "z-slab exchange not implemented yet"); }
int64_t mpmb_migrate_bytes(MpmbHandle h) { (void)h; return 0; }
int mpmb_migrate_pack(MpmbHandle h, int32_t, void *) { return fail(h, MPMB_ERR_STATE, "z-slab exchange not implemented yet"); }
int mpmb_migrate_unpack(MpmbHandle h, int32_t, const void *) { return fail(h, MPMB_ERR_STATE, "z-slab exchange not implemented yet"); }

}  // extern "C"
