// Device Lanczos for Graph.estimate_lmax (pygsp/graphs/graph.py:858-931).
//
// Replaces ARPACK's dsaupd/dseupd reached through scipy.sparse.linalg.eigsh
// (graph.py:911-917): a three-term Lanczos recurrence whose operator is the
// same CSR SpMV kernel family as the filter, with the scalar recurrence
// coefficients kept ON DEVICE so that a whole batch of iterations is enqueued
// without a host round trip.  The host only reads the (alpha, beta) arrays
// back to diagonalise the small tridiagonal matrix and test convergence
// (|beta_m s_m| <= tol |theta|, ARPACK's criterion with tol = 5e-3).
#include "common.cuh"
#include "gspb200.h"

namespace gsp {

constexpr int kVecThreads = 256;
constexpr int kMaxVecBlocks = 2048;   // partial sums per reduction (scal_dev layout)

// sum over the block, returned to every thread; safe to call repeatedly
__device__ __forceinline__ double block_allreduce(double v) {
  __shared__ double part[kVecThreads / 32];
  __shared__ double total;
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x < 32) {
    double s = threadIdx.x < kVecThreads / 32 ? part[threadIdx.x] : 0.0;
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (threadIdx.x == 0) total = s;
  }
  __syncthreads();
  const double out = total;
  __syncthreads();
  return out;
}

// Reductions are two-level and ORDER-FIXED (no floating-point atomics): every
// block writes one partial, every consumer block re-adds the partials in the
// same order, so a run is bit-reproducible.
__device__ __forceinline__ double sum_partials(const double* part, int count) {
  double acc = 0;
  for (int i = threadIdx.x; i < count; i += kVecThreads) acc += part[i];
  return block_allreduce(acc);
}

// counter-based generator: the start vector depends only on (seed, index)
__device__ __forceinline__ double hash_uniform(uint64_t seed, uint64_t i) {
  uint64_t z = seed + 0x9E3779B97F4A7C15ull * (i + 1);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  return double(z >> 11) * (1.0 / 9007199254740992.0) * 2.0 - 1.0;
}

template <typename T>
__global__ void lanczos_seed_kernel(int64_t n, T* v, uint64_t seed, double* part_out) {
  double acc = 0;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += int64_t(gridDim.x) * blockDim.x) {
    const T x = T(hash_uniform(seed, (uint64_t)i));
    v[i] = x;
    acc += double(x) * double(x);
  }
  acc = block_allreduce(acc);
  if (threadIdx.x == 0) part_out[blockIdx.x] = acc;
}

// partial sums of w . v
template <typename T>
__global__ void lanczos_dot_kernel(int64_t n, const T* __restrict__ a, const T* __restrict__ b,
                                   double* part_out) {
  double acc = 0;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += int64_t(gridDim.x) * blockDim.x)
    acc += double(a[i]) * double(b[i]);
  acc = block_allreduce(acc);
  if (threadIdx.x == 0) part_out[blockIdx.x] = acc;
}

// alpha = sum(part_in); w -= alpha v + beta_prev v_prev ; partial |w|^2 -> part_out
template <typename T>
__global__ void lanczos_update_kernel(int64_t n, T* __restrict__ w, const T* __restrict__ v,
                                      const T* __restrict__ v_prev, const double* part_in,
                                      int parts, const double* beta_prev, double* alpha_out,
                                      double* part_out) {
  const double a = sum_partials(part_in, parts);
  if (blockIdx.x == 0 && threadIdx.x == 0) *alpha_out = a;
  const double b = beta_prev ? *beta_prev : 0.0;
  double acc = 0;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += int64_t(gridDim.x) * blockDim.x) {
    double t = double(w[i]) - a * double(v[i]);
    if (v_prev) t -= b * double(v_prev[i]);
    w[i] = T(t);
    acc += t * t;
  }
  acc = block_allreduce(acc);
  if (threadIdx.x == 0) part_out[blockIdx.x] = acc;
}

// beta = sqrt(sum(part_in)); w /= beta
template <typename T>
__global__ void lanczos_scale_kernel(int64_t n, T* w, const double* part_in, int parts,
                                     double* beta_out) {
  const double nb = sqrt(sum_partials(part_in, parts));
  const double inv = nb > 0 ? 1.0 / nb : 0.0;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += int64_t(gridDim.x) * blockDim.x)
    w[i] = T(double(w[i]) * inv);
  if (beta_out && blockIdx.x == 0 && threadIdx.x == 0) *beta_out = nb;
}

static inline int vec_blocks(int64_t n) {
  return (int)std::min<int64_t>(ceil_div(n > 0 ? n : 1, kVecThreads),
                                std::min<int64_t>(int64_t(sm_count()) * 8, kMaxVecBlocks));
}

// Iterations [j0, j1) of the recurrence.  V holds three n-vectors; the Lanczos
// vector v_j lives in slot j % 3.
// scal = alpha[0..cap) | beta[0..cap) | partials A[2048] | partials B[2048].
template <typename T>
int lanczos_run(int64_t n, const int32_t* indptr, const int32_t* indices, const T* data, T* V,
                int j0, int j1, int cap, uint64_t seed, double* scal, cudaStream_t st) {
  GSP_REQUIRE(n >= 1 && j0 >= 0 && j1 <= cap && j0 <= j1, "bad Lanczos range");
  double* alpha = scal;
  double* beta = scal + cap;
  double* part_a = scal + 2 * cap;
  double* part_b = part_a + kMaxVecBlocks;
  const int gb = vec_blocks(n);
  double zero = 0;
  if (j0 == 0) {
    GSP_CUDA(cudaMemsetAsync(scal, 0, sizeof(double) * (2 * cap + 2 * kMaxVecBlocks), st));
    lanczos_seed_kernel<T><<<gb, kVecThreads, 0, st>>>(n, V, seed, part_b);
    lanczos_scale_kernel<T><<<gb, kVecThreads, 0, st>>>(n, V, part_b, gb, nullptr);
    note_launch(1);
    GSP_LAUNCH_CHECK("lanczos_seed");
  }
  for (int j = j0; j < j1; ++j) {
    T* v = V + int64_t(j % 3) * n;
    T* w = V + int64_t((j + 1) % 3) * n;
    const T* vp = j > 0 ? V + int64_t((j + 2) % 3) * n : nullptr;
    // w = L v
    int rc = cheby_step<T>(true, 0, n, indptr, indices, data, v, nullptr, w, w, n, 1, 0, &zero,
                           &zero, 1.0, 0.0, 0.0, st);
    if (rc != GSP_OK) return rc;
    lanczos_dot_kernel<T><<<gb, kVecThreads, 0, st>>>(n, w, v, part_a);
    lanczos_update_kernel<T><<<gb, kVecThreads, 0, st>>>(
        n, w, v, vp, part_a, gb, j > 0 ? beta + j - 1 : nullptr, alpha + j, part_b);
    lanczos_scale_kernel<T><<<gb, kVecThreads, 0, st>>>(n, w, part_b, gb, beta + j);
    note_launch(2);
    GSP_LAUNCH_CHECK("lanczos_step");
  }
  return GSP_OK;
}

}  // namespace gsp

extern "C" {
int gsp_lanczos_f32(int64_t n, const int32_t* indptr, const int32_t* indices, const float* data,
                    float* V3, int j0, int j1, int cap, uint64_t seed, double* scal_dev,
                    void* stream) {
  return gsp::lanczos_run<float>(n, indptr, indices, data, V3, j0, j1, cap, seed, scal_dev,
                                 gsp::as_stream(stream));
}
int gsp_lanczos_f64(int64_t n, const int32_t* indptr, const int32_t* indices, const double* data,
                    double* V3, int j0, int j1, int cap, uint64_t seed, double* scal_dev,
                    void* stream) {
  return gsp::lanczos_run<double>(n, indptr, indices, data, V3, j0, j1, cap, seed, scal_dev,
                                  gsp::as_stream(stream));
}
}
