// C entry points over the DEVICE math header compiled for the host (see stub/cuda_runtime.h).
// Each function loops over n independent items so the Python tests can feed batches.
#include <cuda_runtime.h>  // the stub
#include "mpmb_math.cuh"
#include <cstring>

using namespace mpmb;

extern "C" {

// material_step: plasticity(cdg) + the force of the NEXT rasterize, as k_g2p evaluates it.
void hm_material_step(int64_t n, int kind, const float *params8, const float *cdg, float *F, float *ps, const float *vol, float *force) {
  Material mat;
  mat.kind = kind;
  std::memcpy(mat.p, params8, sizeof(mat.p));
  for (int64_t i = 0; i < n; i++) {
    Mat3 c, f, out;
    std::memcpy(c.m, cdg + 9 * i, 36);
    std::memcpy(f.m, F + 9 * i, 36);
    float s = ps[i];
    material_step(mat, c, f, s, vol[i], out);
    std::memcpy(F + 9 * i, f.m, 36);
    ps[i] = s;
    std::memcpy(force + 9 * i, out.m, 36);
  }
}

// upload-time force (k_pack_particles) and the two-call form plasticity() -> calculate_force()
void hm_calculate_force(int64_t n, int kind, const float *params8, const float *F, const float *ps, const float *vol, float *force) {
  Material mat;
  mat.kind = kind;
  std::memcpy(mat.p, params8, sizeof(mat.p));
  for (int64_t i = 0; i < n; i++) {
    Mat3 f, out;
    std::memcpy(f.m, F + 9 * i, 36);
    calculate_force(mat, f, ps[i], vol[i], out);
    std::memcpy(force + 9 * i, out.m, 36);
  }
}

void hm_plasticity(int64_t n, int kind, const float *params8, const float *cdg, float *F, float *ps) {
  Material mat;
  mat.kind = kind;
  std::memcpy(mat.p, params8, sizeof(mat.p));
  for (int64_t i = 0; i < n; i++) {
    Mat3 c, f;
    std::memcpy(c.m, cdg + 9 * i, 36);
    std::memcpy(f.m, F + 9 * i, 36);
    float s = ps[i];
    plasticity(mat, c, f, s);
    std::memcpy(F + 9 * i, f.m, 36);
    ps[i] = s;
  }
}

// eigen-system of a symmetric tensor given as (xx,yy,zz,xy,xz,yz)
void hm_eig_sym3(int64_t n, const float *A6, float *U9, float *e3) {
  for (int64_t i = 0; i < n; i++) {
    Sym3 A{A6[6 * i + 0], A6[6 * i + 1], A6[6 * i + 2], A6[6 * i + 3], A6[6 * i + 4], A6[6 * i + 5]};
    Mat3 U;
    float e[3];
    eig_sym3<MPMB_EIG_SWEEPS>(A, U, e);
    std::memcpy(U9 + 9 * i, U.m, 36);
    std::memcpy(e3 + 3 * i, e, 12);
  }
}

void hm_bspline_weights(int64_t n, const float *rel, float *w3) {
  for (int64_t i = 0; i < n; i++) bspline_weights(rel[i], w3 + 3 * i);
}

void hm_friction_project0(int64_t n, const float *v3, const float *n3, float friction, float *out3) {
  for (int64_t i = 0; i < n; i++) {
    float3 o = friction_project0(make_float3(v3[3 * i], v3[3 * i + 1], v3[3 * i + 2]), make_float3(n3[3 * i], n3[3 * i + 1], n3[3 * i + 2]), friction);
    out3[3 * i] = o.x; out3[3 * i + 1] = o.y; out3[3 * i + 2] = o.z;
  }
}

}  // extern "C"
