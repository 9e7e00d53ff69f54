// Host <-> device staging of COLUMN CHUNKS of a row-major signal block.
//
// Filter.filter() with host signals (pygsp/filters/filter.py:146-328 takes and returns NumPy
// arrays) is dominated by the two PCIe transfers unless they overlap the recurrence.  The
// recurrence needs every ROW of its operand before its first step, so the block is split by
// COLUMNS: chunk j+1 is uploaded and chunk j-1 downloaded while chunk j is filtered.  A column
// chunk of a row-major (n, nsig) host block is a strided 2-D region: either the copy engines
// move it (gsp_copy2d_async = cudaMemcpy2DAsync, no SM is used) or a small kernel reads /
// writes the pinned host memory directly through its unified address (gsp_stage_cols).
#include "common.cuh"
#include "gspb200.h"

namespace gsp {

// dst[r, 0:width) = src[r, 0:width), 16 bytes per thread and trip; either side may be
// pinned host memory (zero-copy over PCIe).  width, pitches and bases are multiples of 16.
__global__ void stage_cols_kernel(unsigned char* __restrict__ dst, size_t dpitch,
                                  const unsigned char* __restrict__ src, size_t spitch,
                                  int vec_per_row, int64_t total) {
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += int64_t(gridDim.x) * blockDim.x) {
    const int64_t r = i / vec_per_row;
    const int v = int(i - r * vec_per_row);
    const int4 t = *reinterpret_cast<const int4*>(src + r * spitch + size_t(v) * 16);
    *reinterpret_cast<int4*>(dst + r * dpitch + size_t(v) * 16) = t;
  }
}

}  // namespace gsp

extern "C" {

int gsp_copy2d_async(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width_bytes,
                     size_t height, int kind, void* stream) {
  GSP_REQUIRE(kind >= 1 && kind <= 3, "kind: 1 = host to device, 2 = device to host, 3 = device to device");
  if (width_bytes == 0 || height == 0) return GSP_OK;
  const cudaMemcpyKind k = kind == 1 ? cudaMemcpyHostToDevice
                         : kind == 2 ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
  GSP_CUDA(cudaMemcpy2DAsync(dst, dpitch, src, spitch, width_bytes, height, k,
                             gsp::as_stream(stream)));
  return GSP_OK;
}

int gsp_stage_cols(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width_bytes,
                   size_t height, int max_blocks, void* stream) {
  if (width_bytes == 0 || height == 0) return GSP_OK;
  GSP_REQUIRE(width_bytes % 16 == 0 && dpitch % 16 == 0 && spitch % 16 == 0 &&
                  (reinterpret_cast<uintptr_t>(dst) & 15u) == 0 &&
                  (reinterpret_cast<uintptr_t>(src) & 15u) == 0,
              "staging needs 16-byte aligned rows");
  const int vec = int(width_bytes / 16);
  const int64_t total = int64_t(height) * vec;
  int blocks = (int)std::min<int64_t>(gsp::ceil_div(total, 256),
                                      max_blocks > 0 ? max_blocks : gsp::sm_count());
  gsp::stage_cols_kernel<<<blocks, 256, 0, gsp::as_stream(stream)>>>(
      static_cast<unsigned char*>(dst), dpitch, static_cast<const unsigned char*>(src), spitch, vec,
      total);
  GSP_LAUNCH_CHECK("stage_cols");
  return GSP_OK;
}

}  // extern "C"
