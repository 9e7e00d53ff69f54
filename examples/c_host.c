/* A plain-C host of libgspb200: the call a non-Python embedder of the Chebyshev path makes.
 *
 *   gcc -std=c99 -Iinclude examples/c_host.c -o c_host -Lpygsp_b200/_lib -lgspb200 \
 *       -L/usr/local/cuda/lib64 -lcudart -Wl,-rpath,$PWD/pygsp_b200/_lib
 *
 * Filters 8 constant signals on a ring of n vertices (L 1 = 0, so p(L) 1 = p(0) 1 =
 * c_0/2 + sum_k (-1)^k c_k on every vertex) with pygsp/filters/approximations.py:58-114's
 * recurrence running in gsp_cheby_op_f32, and checks that value.  Everything the library needs
 * is device pointers, sizes and a stream; errors come back as codes + gsp_last_error().
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "gspb200.h"

/* the four CUDA runtime calls used, declared here so that the file needs no CUDA headers */
extern int cudaMalloc(void** p, size_t bytes);
extern int cudaFree(void* p);
extern int cudaMemcpy(void* dst, const void* src, size_t bytes, int kind); /* 1 = H2D, 2 = D2H */
extern int cudaDeviceSynchronize(void);

#define CHECK(call)                                                         \
  do {                                                                      \
    int rc_ = (call);                                                       \
    if (rc_ != 0) {                                                         \
      fprintf(stderr, "%s failed (%d): %s\n", #call, rc_, gsp_last_error()); \
      return 1;                                                             \
    }                                                                       \
  } while (0)

int main(void) {
  const int64_t n = 1 << 16, nsig = 8;
  const int m = 21; /* order 20 */
  /* combinatorial Laplacian of the unit-weight ring: 2 on the diagonal, -1 to both neighbours,
   * rows sorted by column (the canonical CSR graph.py:618-620 produces) */
  const int64_t nnz = 3 * n;
  int32_t* indptr = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n + 1));
  int32_t* indices = (int32_t*)malloc(sizeof(int32_t) * (size_t)nnz);
  float* data = (float*)malloc(sizeof(float) * (size_t)nnz);
  for (int64_t i = 0; i < n; ++i) {
    int32_t c[3] = {(int32_t)((i + n - 1) % n), (int32_t)i, (int32_t)((i + 1) % n)};
    float v[3] = {-1.f, 2.f, -1.f};
    for (int a = 0; a < 3; ++a)          /* sort the three columns (wrap-around rows) */
      for (int b = a + 1; b < 3; ++b)
        if (c[b] < c[a]) { int32_t tc = c[a]; c[a] = c[b]; c[b] = tc; float tv = v[a]; v[a] = v[b]; v[b] = tv; }
    indptr[i] = (int32_t)(3 * i);
    for (int a = 0; a < 3; ++a) { indices[3 * i + a] = c[a]; data[3 * i + a] = v[a]; }
  }
  indptr[n] = (int32_t)nnz;
  const double lmax = 4.0;
  double coeffs[21], p0 = 0.0;
  for (int k = 0; k < m; ++k) {          /* any coefficients: a decaying series */
    coeffs[k] = 1.0 / ((k + 1.0) * (k + 1.0));
    p0 += (k == 0 ? 0.5 : (k % 2 ? -1.0 : 1.0)) * coeffs[k];
  }
  float* x = (float*)malloc(sizeof(float) * (size_t)(n * nsig));
  for (int64_t i = 0; i < n * nsig; ++i) x[i] = 1.f;

  void *d_ptr, *d_idx, *d_val, *d_x, *d_r, *d_work;
  if (cudaMalloc(&d_ptr, sizeof(int32_t) * (size_t)(n + 1)) || cudaMalloc(&d_idx, sizeof(int32_t) * (size_t)nnz) ||
      cudaMalloc(&d_val, sizeof(float) * (size_t)nnz) || cudaMalloc(&d_x, sizeof(float) * (size_t)(n * nsig)) ||
      cudaMalloc(&d_r, sizeof(float) * (size_t)(n * nsig)) || cudaMalloc(&d_work, 2 * sizeof(float) * (size_t)(n * nsig))) {
    fprintf(stderr, "cudaMalloc failed (no CUDA device?)\n");
    return 1;
  }
  cudaMemcpy(d_ptr, indptr, sizeof(int32_t) * (size_t)(n + 1), 1);
  cudaMemcpy(d_idx, indices, sizeof(int32_t) * (size_t)nnz, 1);
  cudaMemcpy(d_val, data, sizeof(float) * (size_t)nnz, 1);
  cudaMemcpy(d_x, x, sizeof(float) * (size_t)(n * nsig), 1);

  gsp_tile_plan plan;                    /* TMA-tiled kernel for this (matrix, nsig, nscales) */
  CHECK(gsp_cheby_tile_plan(n, (const int32_t*)d_ptr, nsig, 1, &plan, NULL));
  CHECK(gsp_cheby_op_f32(n, nnz, (const int32_t*)d_ptr, (const int32_t*)d_idx, (const float*)d_val, lmax,
                         coeffs, 1, m, (const float*)d_x, nsig, (float*)d_r, (float*)d_work,
                         plan.rows_per_tile > 0 ? &plan : NULL, NULL));
  cudaDeviceSynchronize();
  cudaMemcpy(x, d_r, sizeof(float) * (size_t)(n * nsig), 2);
  double worst = 0.0;
  for (int64_t i = 0; i < n * nsig; ++i) worst = fmax(worst, fabs(x[i] - p0));
  printf("p(0) = %.7f, max |r - p(0)| = %.2e over %lld values (tiled kernel: %s)\n", p0, worst,
         (long long)(n * nsig), plan.rows_per_tile > 0 ? "yes" : "no");
  cudaFree(d_ptr); cudaFree(d_idx); cudaFree(d_val); cudaFree(d_x); cudaFree(d_r); cudaFree(d_work);
  free(indptr); free(indices); free(data); free(x);
  return worst <= 1e-5 * fabs(p0) + 1e-6 ? 0 : 2;
}
