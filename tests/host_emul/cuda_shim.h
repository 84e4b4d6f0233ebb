// cuda_shim.h - plain C++ stand-ins for the CUDA builtins used by vela_device.cuh / vela_chain.cuh, so that the very
// same per-(TOA, walker) arithmetic can be compiled with g++ (-ffp-contract=off, like nvcc -fmad=false) and compared
// with the oracle on the CPU.  TEST INFRASTRUCTURE ONLY: nothing in the product includes this file.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>

#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
#define __constant__
#define __grid_constant__

#define CUDART_PI 3.1415926535897931e+0
#define CUDART_NAN std::numeric_limits<double>::quiet_NaN()
#define CUDART_INF std::numeric_limits<double>::infinity()

static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __ddiv_rn(double a, double b) { return a / b; }
static inline double __fma_rn(double a, double b, double c) { return std::fma(a, b, c); }
template <class T>
static inline T __ldg(const T* p) { return *p; }
static inline long long __double_as_longlong(double v) { long long r; std::memcpy(&r, &v, 8); return r; }
static inline double __longlong_as_double(long long v) { double r; std::memcpy(&r, &v, 8); return r; }
static inline int __double2hiint(double v) { return (int)(__double_as_longlong(v) >> 32); }
static inline int __double2loint(double v) { return (int)(__double_as_longlong(v) & 0xffffffffLL); }
static inline double __hiloint2double(int hi, int lo) {
    return __longlong_as_double(((long long)hi << 32) | (unsigned int)lo);
}
static inline double rsqrt(double x) { return 1.0 / std::sqrt(x); }
// sincos and exp10 come from glibc (GNU extensions, declared by <cmath> under _GNU_SOURCE)
using std::isfinite;
using std::isnan;

// thread coordinates: the emulator runs one "thread" at a time
struct EmulDim { unsigned x, y, z; };
static thread_local EmulDim blockIdx = {0, 0, 0}, threadIdx = {0, 0, 0}, blockDim = {1, 1, 1};
