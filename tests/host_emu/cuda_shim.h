// cuda_shim.h -- TEST INFRASTRUCTURE ONLY: lets g++ compile the element-wise kernels of maskflownet_b200/csrc/augment.cu for
// the host, one "thread" at a time (no shuffles, no real barriers: kernels that need them are not emulated).  The development
// container has no GPU; this is how the kernel arithmetic is checked against the oracle before a GPU run (tests/test_host_logic.py).
#pragma once
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <algorithm>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static

struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
static dim3 threadIdx, blockIdx, blockDim, gridDim;

template <typename T>
static inline T __ldg(const T* p) { return *p; }
// round-to-nearest single operations that the compiler may not contract (g++ -ffp-contract=off is also passed)
static inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
static inline float __fsub_rn(float a, float b) { volatile float r = a - b; return r; }
static inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
static inline float __fdiv_rn(float a, float b) { volatile float r = a / b; return r; }
static inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((unsigned long long)a * b) >> 32); }
static inline float __shfl_xor_sync(unsigned, float v, int) { return v; }   // placeholder: kernels using it are not emulated
static inline void __syncthreads() {}
using std::max;
using std::min;
