// Shared helpers for libgspb200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <algorithm>
#include <stdio.h>
#include <string.h>

#define GSP_OK 0
#define GSP_ERR_ARG (-1)
#define GSP_ERR_CUDA (-2)
#define GSP_ERR_UNSUPPORTED (-3)

namespace gsp {

// thread-local message returned by gsp_last_error()
char* error_buffer();

inline int fail(int code, const char* fmt, const char* a = "", const char* b = "") {
  snprintf(error_buffer(), 512, fmt, a, b);
  return code;
}

// kernels launched by this library since load (bench.py's gpu_launches)
void note_launch(int n);

inline int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return GSP_OK;
  return fail(GSP_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
}

#define GSP_CUDA(call)                                        \
  do {                                                        \
    int _rc = gsp::check_cuda((call), #call);                 \
    if (_rc != GSP_OK) return _rc;                            \
  } while (0)

#define GSP_LAUNCH_CHECK(name)                                \
  do {                                                        \
    gsp::note_launch(1);                                      \
    int _rc = gsp::check_cuda(cudaGetLastError(), name);      \
    if (_rc != GSP_OK) return _rc;                            \
  } while (0)

#define GSP_REQUIRE(cond, msg)                                \
  do {                                                        \
    if (!(cond)) return gsp::fail(GSP_ERR_ARG, "%s (%s)", msg, #cond); \
  } while (0)

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// number of SMs of the current device (cached per device)
int sm_count();

// ----- the fused recurrence step (csrc/cheby.cu), shared by Lanczos and the C ABI ---------
// x_new = alpha (L x_cur) + beta x_cur + gamma x_old over rows [rb, re);
//   add_source == false: r_i (+)= ck[i] x_new          (reference order, approximations.py:107-109)
//   add_source == true : x_new += ck[0] * r[row, :]    (Clenshaw form, r is a read-only source)
template <typename T>
int cheby_step(bool first, int64_t rb, int64_t re, const int32_t* indptr, const int32_t* indices,
               const T* vals, const T* x_cur, const T* x_old, T* x_new, T* r, int64_t r_rows,
               int nsig, int nscales, const double* ck, const double* c0, double alpha,
               double beta, double gamma, cudaStream_t st, bool add_source = false,
               const int64_t* out_perm = nullptr);   // x_new row of local row i is out_perm[i]

// dst[i,:] = src[idx[i],:] (scatter: dst[idx[i],:] = src[i,:]) -- csrc/graph.cu
template <typename T>
int move_rows(bool scatter, int64_t rows, const int64_t* idx, const T* src, int64_t width, T* dst,
              cudaStream_t st);

// ----- vector types: 16-byte packets of T --------------------------------
template <typename T, int VEC> struct Pack;
template <> struct Pack<float, 4> { typedef float4 type; };
template <> struct Pack<float, 2> { typedef float2 type; };
template <> struct Pack<float, 1> { typedef float type; };
template <> struct Pack<double, 2> { typedef double2 type; };
template <> struct Pack<double, 1> { typedef double type; };

template <typename T, int VEC>
struct Vec {
  T v[VEC];
};

template <typename T, int VEC>
__device__ __forceinline__ Vec<T, VEC> load_vec(const T* p) {
  typedef typename Pack<T, VEC>::type P;
  union { P p; Vec<T, VEC> v; } u;
  u.p = *reinterpret_cast<const P*>(p);
  return u.v;
}

// read-only (non-coherent) path: data that no thread of this launch writes
template <typename T, int VEC>
__device__ __forceinline__ Vec<T, VEC> load_vec_ro(const T* p) {
  typedef typename Pack<T, VEC>::type P;
  union { P p; Vec<T, VEC> v; } u;
  u.p = __ldg(reinterpret_cast<const P*>(p));
  return u.v;
}

// streaming load: touched once per launch, do not keep in L1
template <typename T, int VEC>
__device__ __forceinline__ Vec<T, VEC> load_vec_stream(const T* p) {
  typedef typename Pack<T, VEC>::type P;
  union { P p; Vec<T, VEC> v; } u;
  u.p = __ldcs(reinterpret_cast<const P*>(p));
  return u.v;
}

template <typename T, int VEC>
__device__ __forceinline__ void store_vec_stream(T* p, const Vec<T, VEC>& v) {
  typedef typename Pack<T, VEC>::type P;
  union { P p; Vec<T, VEC> v; } u;
  u.v = v;
  __stcs(reinterpret_cast<P*>(p), u.p);
}

}  // namespace gsp
