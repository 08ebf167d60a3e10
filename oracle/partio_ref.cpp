// TEST INFRASTRUCTURE (not part of the product): drives the reference's OWN vendored Partio
// (/root/reference/external/partio, compiled from its sources where they lie by oracle/Makefile
// into oracle/_ref/) the way MPM<dim>::write_partio does (src/visualize.cpp:16-100), so that the
// product's frame writer (taichi_mpm_b200/bgeo.py) can be pinned byte for byte against files the
// reference's writer produces.  Attribute order = the order of the addAttribute calls there:
// position, type, index, limit, v [, m, boundary_normal, debug, states, boundary_distance,
// near_boundary, apic_frobenius_norm when verbose_bgeo].
#include <Partio.h>
#include <cstdint>
#include <cstring>

extern "C" int ref_write_partio(const char *file_name, int64_t n, const float *pos3, const float *v3, const int32_t *type1,
                                const int32_t *index1, const int32_t *limit3, int verbose, const float *m1, const float *normal3,
                                const float *debug3, const int32_t *states1, const float *dist1, const int32_t *near1,
                                const float *apic1) {
  Partio::ParticlesDataMutable *parts = Partio::create();
  Partio::ParticleAttribute posH, vH, mH, typeH, normH, statH, boundH, distH, debugH, indexH, limitH, apicH;
  posH = parts->addAttribute("position", Partio::VECTOR, 3);
  typeH = parts->addAttribute("type", Partio::INT, 1);
  indexH = parts->addAttribute("index", Partio::INT, 1);
  limitH = parts->addAttribute("limit", Partio::INT, 3);
  vH = parts->addAttribute("v", Partio::VECTOR, 3);
  if (verbose) {
    mH = parts->addAttribute("m", Partio::VECTOR, 1);
    normH = parts->addAttribute("boundary_normal", Partio::VECTOR, 3);
    debugH = parts->addAttribute("debug", Partio::VECTOR, 3);
    statH = parts->addAttribute("states", Partio::INT, 1);
    distH = parts->addAttribute("boundary_distance", Partio::FLOAT, 1);
    boundH = parts->addAttribute("near_boundary", Partio::INT, 1);
    apicH = parts->addAttribute("apic_frobenius_norm", Partio::FLOAT, 1);
  }
  for (int64_t i = 0; i < n; i++) {
    int idx = parts->addParticle();
    if (verbose) {
      parts->dataWrite<float>(mH, idx)[0] = m1[i];
      std::memcpy(parts->dataWrite<float>(normH, idx), normal3 + 3 * i, 12);
      std::memcpy(parts->dataWrite<float>(debugH, idx), debug3 + 3 * i, 12);
      parts->dataWrite<int>(statH, idx)[0] = states1[i];
      parts->dataWrite<int>(boundH, idx)[0] = near1[i];
      parts->dataWrite<float>(distH, idx)[0] = dist1[i];
      parts->dataWrite<float>(apicH, idx)[0] = apic1[i];
    }
    std::memcpy(parts->dataWrite<float>(vH, idx), v3 + 3 * i, 12);
    parts->dataWrite<int>(typeH, idx)[0] = type1[i];
    parts->dataWrite<int>(indexH, idx)[0] = index1[i];
    std::memcpy(parts->dataWrite<int>(limitH, idx), limit3 + 3 * i, 12);
    std::memcpy(parts->dataWrite<float>(posH, idx), pos3 + 3 * i, 12);
  }
  Partio::write(file_name, *parts);
  parts->release();
  return 0;
}
