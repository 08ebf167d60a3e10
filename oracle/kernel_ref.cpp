// TEST INFRASTRUCTURE.  The reference's interpolation kernels (src/kernel.h, included where it lies,
// unmodified) compiled against the stand-in core headers oracle/taichi_stub/taichi/*.h: MPMKernel<3,1..3>
// (weights and gradients, src/kernel.h:75-163) and MPMFastKernel32 (the 27 products the optimized
// transfers use, src/kernel.h:166-209).  The oracle's weight functions are pinned against these.
#include REF_KERNEL_SOURCE
#include <cstdint>

using namespace taichi;

template <int order>
static void weights(const float *pos, float inv_dx, int *start, float *w, float *dw) {
  constexpr int ks = order + 1;
  VectorND<3, real> p(pos[0], pos[1], pos[2]);
  MPMKernel<3, order> k(p, inv_dx);
  for (int d = 0; d < 3; d++) start[d] = MPMKernel<3, order>::get_stencil_start(pos[d]);
  for (int a = 0; a < ks; a++)
    for (int b = 0; b < ks; b++)
      for (int c = 0; c < ks; c++) {
        VectorND<4, real> r = k.get_dw_w(VectorND<3, int>(a, b, c));
        int o = (a * ks + b) * ks + c;
        w[o] = r[3];
        dw[o * 3 + 0] = r[0]; dw[o * 3 + 1] = r[1]; dw[o * 3 + 2] = r[2];
      }
}

extern "C" {
// pos in grid units; w[(order+1)^3], dw[(order+1)^3][3] (already scaled by inv_dx, kernel.h:35)
int ref_kernel(int order, const float *pos, float inv_dx, int *start, float *w, float *dw) {
  switch (order) {
    case 1: weights<1>(pos, inv_dx, start, w, dw); return 0;
    case 2: weights<2>(pos, inv_dx, start, w, dw); return 0;
    case 3: weights<3>(pos, inv_dx, start, w, dw); return 0;
  }
  return -1;
}
void ref_fast_kernel32(const float *pos, float inv_dx, float *w27, float *dw27x3) {
  MPMFastKernel32 k(VectorND<3, real>(pos[0], pos[1], pos[2]), inv_dx);
  for (int a = 0; a < 3; a++)
    for (int b = 0; b < 3; b++)
      for (int c = 0; c < 3; c++) {
        VectorND<4, real> r = k.get_dw_w(VectorND<3, int>(a, b, c));
        int o = (a * 3 + b) * 3 + c;
        w27[o] = r[3];
        dw27x3[o * 3 + 0] = r[0]; dw27x3[o * 3 + 1] = r[1]; dw27x3[o * 3 + 2] = r[2];
      }
}
float ref_inv_D(int order) { return order == 1 ? MPMKernel<3, 1>::inv_D() : (order == 2 ? MPMKernel<3, 2>::inv_D() : MPMKernel<3, 3>::inv_D()); }
}
