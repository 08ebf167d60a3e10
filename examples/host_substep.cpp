// Pure C++ host of the C-ABI (no Python, no torch): builds a small sand block on the reference's
// benchmark lattice (src/mpm.cpp:164-180), runs substeps through include/mpmb.h and prints the
// centre of mass.  Build:  g++ -O2 -std=c++17 examples/host_substep.cpp -Iinclude \
//                              -Ltaichi_mpm_b200/lib -lmpmb -Wl,-rpath,$PWD/taichi_mpm_b200/lib -o host_substep
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "mpmb.h"

#define CHECK(call)                                                                  \
  do {                                                                               \
    int rc_ = (call);                                                                \
    if (rc_ != MPMB_OK) {                                                            \
      std::fprintf(stderr, "%s -> %d: %s\n", #call, rc_, mpmb_last_error(h));        \
      return 1;                                                                      \
    }                                                                                \
  } while (0)

int main(int argc, char **argv) {
  const int res = 64, nsub = argc > 1 ? std::atoi(argv[1]) : 50;
  MpmbConfig cfg{};
  cfg.res[0] = cfg.res[1] = cfg.res[2] = res;
  cfg.dx = 1.0f / res;
  cfg.dt = 2e-5f;
  cfg.gravity[1] = -10.f;
  cfg.particle_gravity = 1;
  cfg.clean_boundary = 1;
  cfg.world = 1;
  MpmbHandle h = nullptr;
  CHECK(mpmb_create(&cfg, &h));
  // SandParticle defaults (src/particles.cpp:570-597)
  const float s = std::sin(30.0f / 180.0f * 3.141592653f);
  const float sand[5] = {136038.0f, 204057.0f, std::sqrt(2.0f / 3.0f) * 2.0f * s / (3.0f - s), 0.0f, 1.0f};
  CHECK(mpmb_set_material(h, 0, MPMB_MAT_SAND, sand, 5));
  const float floor_plane[4] = {0.f, 1.f, 0.f, -10.0f};  // phi = Y - 10 (grid units)
  CHECK(mpmb_set_planes(h, 1, floor_plane, 0.4f));
  std::vector<float> x, v, mass, vol;
  const float dx = cfg.dx;
  for (int i = 24; i < 40; i++)
    for (int j = 10; j < 26; j++)
      for (int k = 24; k < 40; k++)
        for (int c = 0; c < 8; c++) {
          x.push_back((i + 0.5f + ((c & 1) ? 0.25f : -0.25f)) * dx);
          x.push_back((j + 0.5f + ((c & 2) ? 0.25f : -0.25f)) * dx);
          x.push_back((k + 0.5f + ((c & 4) ? 0.25f : -0.25f)) * dx);
          v.insert(v.end(), {0.f, 0.f, 0.f});
          vol.push_back(dx * dx * dx / 8);
          mass.push_back(dx * dx * dx / 8 * 400.f);
        }
  const int64_t n = (int64_t)mass.size();
  CHECK(mpmb_upload_particles(h, n, x.data(), v.data(), nullptr, nullptr, mass.data(), vol.data(), nullptr, nullptr));
  CHECK(mpmb_substep(h, nsub));
  int64_t alive = 0;
  std::vector<uint32_t> id(n);
  CHECK(mpmb_download_particles(h, n, &alive, id.data(), x.data(), v.data(), nullptr, nullptr, nullptr, nullptr, nullptr, nullptr));
  double com[3] = {0, 0, 0}, vy = 0;
  for (int64_t i = 0; i < alive; i++) {
    for (int d = 0; d < 3; d++) com[d] += x[3 * i + d];
    vy += v[3 * i + 1];
  }
  std::printf("alive=%lld com=(%.6f %.6f %.6f) mean_vy=%.6f after %d substeps\n", (long long)alive, com[0] / alive, com[1] / alive,
              com[2] / alive, vy / alive, nsub);
  CHECK(mpmb_destroy(h));
  return 0;
}
