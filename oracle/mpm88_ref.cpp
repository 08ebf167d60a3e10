// TEST INFRASTRUCTURE.  Compiles the reference's own 88-line 2-D program (mls-mpm88.cpp, included
// where it lies, unmodified) against the stand-in header oracle/taichi_stub/taichi.h and exposes its
// advance() so that the oracle's restatement (mpm88_advance in mpm_oracle.cpp) can be pinned against
// the reference's lines 16-69 executed here.  Only the header vocabulary is ours (see taichi_stub).
#define main mls_mpm88_reference_main  // keep the program's endless GUI loop out of the way
#include REF88_SOURCE
#undef main
#include <cstdint>

extern "C" {
// Replaces the particle set (x, v: [n][2]; F, C: [n][4] row-major 2x2; Jp: [n]) and the plastic switch.
void ref88_set(int64_t n, const float *x, const float *v, const float *F, const float *C, const float *Jp, int plastic_flag) {
  plastic = plastic_flag != 0;
  particles.clear();
  for (int64_t i = 0; i < n; i++) {
    Particle p(Vec(x[2 * i], x[2 * i + 1]), 0, Vec(v[2 * i], v[2 * i + 1]));
    p.F[0] = Vec(F[4 * i + 0], F[4 * i + 2]);  // column 0 = (F00, F10)
    p.F[1] = Vec(F[4 * i + 1], F[4 * i + 3]);
    p.C[0] = Vec(C[4 * i + 0], C[4 * i + 2]);
    p.C[1] = Vec(C[4 * i + 1], C[4 * i + 3]);
    p.Jp = Jp[i];
    particles.push_back(p);
  }
}
void ref88_advance(int steps) {
  for (int s = 0; s < steps; s++) advance(dt);
}
void ref88_get(float *x, float *v, float *F, float *C, float *Jp) {
  for (size_t i = 0; i < particles.size(); i++) {
    const Particle &p = particles[i];
    x[2 * i] = p.x.x; x[2 * i + 1] = p.x.y;
    v[2 * i] = p.v.x; v[2 * i + 1] = p.v.y;
    F[4 * i + 0] = p.F[0].x; F[4 * i + 2] = p.F[0].y; F[4 * i + 1] = p.F[1].x; F[4 * i + 3] = p.F[1].y;
    C[4 * i + 0] = p.C[0].x; C[4 * i + 2] = p.C[0].y; C[4 * i + 1] = p.C[1].x; C[4 * i + 3] = p.C[1].y;
    Jp[i] = p.Jp;
  }
}
// grid[(n+1)^2][3] after the last advance (velocity, mass), row-major [i][j]
void ref88_grid(float *out) {
  for (int i = 0; i <= n; i++)
    for (int j = 0; j <= n; j++)
      for (int c = 0; c < 3; c++) out[(size_t(i) * (n + 1) + j) * 3 + c] = grid[i][j][c];
}
int ref88_n() { return n; }
}
