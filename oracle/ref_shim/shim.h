// Shim that lets g++ compile the reference's BSDF device headers in place
// (/root/reference/Lumen_Engine/LumenPT/src/CUDAKernels/{disney,ggxmdf,frosted,bsdf_math}.cuh).
// TEST INFRASTRUCTURE ONLY. Nothing from the reference is copied; the headers are
// included from where they lie and only the symbols nvcc would have provided are supplied here.
#pragma once
#include <cmath>
#include <cstdlib>
#include <algorithm>
#include <cuda_runtime.h>   // host_defines.h makes __device__/__forceinline__ harmless under g++
// sutil/vec_math.h defines a scalar lerp that collides with bsdf_math.cuh's own: hide sutil's.
#define lerp sutil_scalar_lerp_hidden
#include <sutil/vec_math.h>
#undef lerp
static inline float3 lerp(const float3& a, const float3& b, const float t) { return a + t * (b - a); }
// CUDA math builtins that the headers use unqualified
static inline float saturate(float x) { return x < 0.f ? 0.f : (x > 1.f ? 1.f : x); }
static inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
using std::min; using std::max; using std::abs;
// On the device CUDA supplies float overloads of the unqualified maths calls the headers make (fabs(float), sqrt(float), pow(float,float));
// plain <cmath> leaves only C's double versions in the global namespace, which would silently evaluate those expressions in double.
using std::fabs; using std::sqrt; using std::pow; using std::exp; using std::log; using std::sin; using std::cos; using std::floor;
static inline void sincosf_(float x, float* s, float* c) { *s = sinf(x); *c = cosf(x); }
// disney.cuh / frosted.cuh declare the `adjoint` default parameter only under __CUDACC__ but use it
// in the body unconditionally; a file-scope constant with the default value restores that behaviour.
static const bool adjoint = false;
