// Host stand-in for <cuda_runtime.h>, used ONLY by tests/host_math: it lets plain g++ compile
// taichi_mpm_b200/csrc/mpmb_math.cuh so that the arithmetic of the CUDA kernels' constitutive
// step can be checked against the oracle on a machine without a GPU.  Test infrastructure; the
// product never includes this file (nvcc finds the real header first).
#pragma once
#include <math.h>
#include <stdint.h>
#define MPMB_HOST_MATH 1
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
struct float3 { float x, y, z; };
struct float4 { float x, y, z, w; };
static inline float3 make_float3(float x, float y, float z) { return float3{x, y, z}; }
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
// device intrinsics -> their IEEE host meaning (a warp of one lane)
// (glibc already declares __expf/__logf as internal names, hence macros)
#define __expf(x) expf(x)
#define __logf(x) logf(x)
#define __fdividef(a, b) ((a) / (b))
static inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
static inline unsigned __activemask() { return 1u; }
static inline int __any_sync(unsigned, int p) { return p; }
static inline int __reduce_max_sync(unsigned, int v) { return v; }
