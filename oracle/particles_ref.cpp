// TEST INFRASTRUCTURE.  Compiles the reference's own constitutive models (src/particles.cpp with
// src/particles.h and src/mpm_fwd.h, included where they lie, unmodified) against the stand-in core
// oracle/taichi_stub/taichi/*.h and exposes, per registered particle type, exactly what the hot path
// calls: plasticity(cdg) (src/transfer.cpp:950) and calculate_force() (src/transfer.cpp:509), plus
// friction_project (src/mpm_fwd.h:25-57).  The oracle's restatements are pinned against these.
#include REF_PARTICLES_SOURCE
#include <cstdint>
#include <cstring>

namespace {
using namespace taichi;
using M3 = MatrixND<3, real>;

M3 load(const float *m) {  // column-major 9 floats, m[c*3+r]
  M3 r;
  for (int c = 0; c < 3; c++) for (int q = 0; q < 3; q++) r[c][q] = m[c * 3 + q];
  return r;
}
void store(const M3 &a, float *m) {
  for (int c = 0; c < 3; c++) for (int q = 0; q < 3; q++) m[c * 3 + q] = a[c][q];
}

// kind / params: the oracle's numbering and parameter vectors (oracle/mpm_oracle.cpp, MaterialKind)
template <class P>
void run(P &p, const float *cdg, float *F, float vol, float *force, int do_plasticity) {
  p.dg_e = load(F);
  p.vol = vol;
  if (do_plasticity) p.plasticity(load(cdg));
  store(p.dg_e, F);
  if (force) store(p.calculate_force(), force);
}
}  // namespace

extern "C" {
// One particle: optional plasticity(cdg), then calculate_force() of the resulting state.
// ps is the plastic scalar in/out (snow Jp, water j, sand logJp).  Returns 0, or -1 for an unknown kind.
int ref_particle(int kind, const float *params, const float *cdg, float *F, float *ps, float vol, float *force, int do_plasticity) {
  Config cfg;
  switch (kind) {
    case 0: {  // linear: E, nu enter through mu, lambda (src/particles.cpp:317-323)
      LinearParticle<3> p;
      p.initialize(cfg);
      p.mu = params[0]; p.lambda = params[1];
      run(p, cdg, F, vol, force, do_plasticity);
      return 0;
    }
    case 1: {
      JellyParticle<3> p;
      p.initialize(cfg);
      p.mu = params[0]; p.lambda = params[1];
      run(p, cdg, F, vol, force, do_plasticity);
      return 0;
    }
    case 2: {
      SnowParticle<3> p;
      cfg.set("mu_0", params[0]).set("lambda_0", params[1]).set("hardening", params[2]).set("theta_c", params[3]).set("theta_s", params[4])
          .set("min_Jp", params[5]).set("max_Jp", params[6]).set("Jp", *ps);
      p.initialize(cfg);
      run(p, cdg, F, vol, force, do_plasticity);
      *ps = p.Jp;
      return 0;
    }
    case 3: {
      WaterParticle<3> p;
      cfg.set("k", params[0]).set("gamma", params[1]);
      p.initialize(cfg);
      p.j = *ps;
      run(p, cdg, F, vol, force, do_plasticity);
      *ps = p.j;
      return 0;
    }
    case 4: {
      SandParticle<3> p;
      cfg.set("mu_0", params[0]).set("lambda_0", params[1]).set("cohesion", params[3]).set("beta", params[4]);
      p.initialize(cfg);
      p.alpha = params[2];  // the oracle passes alpha itself; initialize() derives it from friction_angle (592-593)
      p.logJp = *ps;
      run(p, cdg, F, vol, force, do_plasticity);
      *ps = p.logJp;
      return 0;
    }
    case 5: {  // elastic (src/particles.cpp:764-841): E, nu enter through mu_0, lambda_0 (775-781)
      ElasticParticle<3> p;
      p.initialize(cfg);
      p.mu_0 = params[0]; p.lambda_0 = params[1];
      run(p, cdg, F, vol, force, do_plasticity);
      return 0;
    }
    case 6: {  // von_mises (src/particles.cpp:679-761)
      VonMisesParticle<3> p;
      cfg.set("yield_stress", params[2]);
      p.initialize(cfg);
      p.mu_0 = params[0]; p.lambda_0 = params[1];
      run(p, cdg, F, vol, force, do_plasticity);
      return 0;
    }
    case 7: {  // visco (src/particles.cpp:40-163); scalar = visco_tau
      ViscoParticle<3> p;
      cfg.set("nu", params[2]).set("kappa", params[3]).set("base_delta_t", params[4]);
      p.initialize(cfg);
      p.mu_0 = params[0]; p.lambda_0 = params[1];
      p.visco_tau = *ps;
      run(p, cdg, F, vol, force, do_plasticity);
      *ps = p.visco_tau;
      return 0;
    }
  }
  return -1;
}

// MPMParticle::get_allowed_dt(dx) of the registered type (the strength limit AsyncMPM's scheduler reads,
// src/async/async_mpm.cpp:105-110).  Returns the reference's value; -1 for an unknown kind.
float ref_allowed_dt(int kind, const float *params, const float *F, float ps, float mass, float vol, const float *v, float dx) {
  Config cfg;
  auto fill = [&](auto &p) {
    p.dg_e = load(F);
    p.vol = vol;
    p.set_mass(mass);
    p.set_velocity(VectorND<3, real>(v[0], v[1], v[2]));
    return p.get_allowed_dt(dx);
  };
  switch (kind) {
    case 0: { LinearParticle<3> p; p.initialize(cfg); return fill(p); }
    case 1: { JellyParticle<3> p; p.initialize(cfg); return fill(p); }
    case 2: {
      SnowParticle<3> p;
      cfg.set("mu_0", params[0]).set("lambda_0", params[1]).set("hardening", params[2]).set("Jp", ps);
      p.initialize(cfg);
      return fill(p);
    }
    case 3: { WaterParticle<3> p; cfg.set("k", params[0]).set("gamma", params[1]); p.initialize(cfg); p.j = ps; return fill(p); }
    case 4: { SandParticle<3> p; cfg.set("mu_0", params[0]).set("lambda_0", params[1]); p.initialize(cfg); return fill(p); }
    case 5: { ElasticParticle<3> p; p.initialize(cfg); p.mu_0 = params[0]; p.lambda_0 = params[1]; return fill(p); }
    case 6: { VonMisesParticle<3> p; p.initialize(cfg); p.mu_0 = params[0]; p.lambda_0 = params[1]; return fill(p); }
    case 7: { ViscoParticle<3> p; p.initialize(cfg); p.mu_0 = params[0]; p.lambda_0 = params[1]; return fill(p); }
  }
  return -1.0f;
}

// alpha as SandParticle::initialize computes it from the friction angle in degrees (src/particles.cpp:591-593)
float ref_sand_alpha(float friction_angle) {
  Config cfg;
  cfg.set("friction_angle", friction_angle);
  SandParticle<3> p;
  p.initialize(cfg);
  return p.alpha;
}

// default parameters after initialize(empty config): out = {mu, lambda, ...} in the oracle's layout
int ref_default_params(int kind, float *out) {
  Config cfg;
  for (int i = 0; i < 8; i++) out[i] = 0;
  switch (kind) {
    case 0: { LinearParticle<3> p; p.initialize(cfg); out[0] = p.mu; out[1] = p.lambda; return 0; }
    case 1: { JellyParticle<3> p; p.initialize(cfg); out[0] = p.mu; out[1] = p.lambda; return 0; }
    case 2: { SnowParticle<3> p; p.initialize(cfg); out[0] = p.mu_0; out[1] = p.lambda_0; out[2] = p.hardening; out[3] = p.theta_c; out[4] = p.theta_s; out[5] = p.min_Jp; out[6] = p.max_Jp; return 0; }
    case 3: { WaterParticle<3> p; p.initialize(cfg); out[0] = p.k; out[1] = p.gamma; return 0; }
    case 4: { SandParticle<3> p; p.initialize(cfg); out[0] = p.mu_0; out[1] = p.lambda_0; out[2] = p.alpha; out[3] = p.cohesion; out[4] = p.beta; return 0; }
    case 5: { ElasticParticle<3> p; p.initialize(cfg); out[0] = p.mu_0; out[1] = p.lambda_0; return 0; }
    case 6: { VonMisesParticle<3> p; p.initialize(cfg); out[0] = p.mu_0; out[1] = p.lambda_0; out[2] = p.yield_stress; return 0; }
    case 7: { ViscoParticle<3> p; p.initialize(cfg); out[0] = p.mu_0; out[1] = p.lambda_0; out[2] = p.visco_nu; out[3] = p.visco_kappa; out[4] = p.dt; return 0; }
  }
  return -1;
}

void ref_friction_project(const float *v, const float *base, const float *n, float friction, float *out) {
  VectorND<3, real> r = friction_project<3>(VectorND<3, real>(v[0], v[1], v[2]), VectorND<3, real>(base[0], base[1], base[2]),
                                            VectorND<3, real>(n[0], n[1], n[2]), friction);
  out[0] = r[0]; out[1] = r[1]; out[2] = r[2];
}
}
