// TEST INFRASTRUCTURE: the few CUB entry points the engine uses, on top of the SIMT emulator.
#pragma once
#include "../../simt.h"
namespace cub {
// block-wide collectives: deposit in shared storage, barrier, combine in thread order, barrier
template <class T, int BLOCK>
struct BlockReduce {
  struct TempStorage { T v[BLOCK]; };
  TempStorage &s;
  explicit BlockReduce(TempStorage &t) : s(t) {}
  T Sum(T x) {                       // result valid in thread 0 (as in CUB); here every thread gets it
    s.v[threadIdx.x] = x;
    __syncthreads();
    T r = 0;
    for (unsigned i = 0; i < blockDim.x; i++) r += s.v[i];
    __syncthreads();
    return r;
  }
};
template <class T, int BLOCK>
struct BlockScan {
  struct TempStorage { T v[BLOCK]; };
  TempStorage &s;
  explicit BlockScan(TempStorage &t) : s(t) {}
  void ExclusiveSum(T x, T &out) { T agg; ExclusiveSum(x, out, agg); }
  void ExclusiveSum(T x, T &out, T &agg) {
    s.v[threadIdx.x] = x;
    __syncthreads();
    T r = 0, a = 0;
    for (unsigned i = 0; i < blockDim.x; i++) { if (i < threadIdx.x) r += s.v[i]; a += s.v[i]; }
    __syncthreads();
    out = r; agg = a;
  }
};
struct DeviceRadixSort {
  template <class K, class V>
  static cudaError_t SortPairs(void *tmp, size_t &bytes, const K *kin, K *kout, const V *vin, V *vout, int n, int begin_bit = 0, int end_bit = sizeof(K) * 8,
                               cudaStream_t = nullptr) {
    if (!tmp) { bytes = 16; return cudaSuccess; }
    std::vector<int> idx(n);
    for (int i = 0; i < n; i++) idx[i] = i;
    const unsigned long long mask = end_bit - begin_bit >= 64 ? ~0ull : ((1ull << (end_bit - begin_bit)) - 1ull);
    std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) { return (((unsigned long long)kin[a] >> begin_bit) & mask) < (((unsigned long long)kin[b] >> begin_bit) & mask); });
    for (int i = 0; i < n; i++) { kout[i] = kin[idx[i]]; vout[i] = vin[idx[i]]; }
    return cudaSuccess;
  }
};
struct DeviceScan {
  template <class In, class Out>
  static cudaError_t ExclusiveSum(void *tmp, size_t &bytes, In in, Out out, int n, cudaStream_t = nullptr) {
    if (!tmp) { bytes = 16; return cudaSuccess; }
    long long run = 0;
    for (int i = 0; i < n; i++) { auto v = in[i]; out[i] = (decltype(+out[i]))run; run += v; }
    return cudaSuccess;
  }
};
}  // namespace cub
