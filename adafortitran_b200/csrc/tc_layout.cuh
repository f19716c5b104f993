// Memory layouts of the bf16 tcgen05 operand images.
//
// Every MMA operand is kept -- in HBM and in shared memory alike -- as a byte-exact image of the
// canonical UMMA K-major SWIZZLE_128B shared-memory layout, so that moving an operand is a plain
// 1-D bulk copy (cp.async.bulk) and needs no tensor map:
//
//   K is cut into chunks of 64 bf16 (= 128 bytes).  A chunk block holds `rows` rows of 128 bytes
//   each (row pitch 128 B, so an 8-row group is the 1024-byte swizzle atom; SBO = 1024).  Inside a
//   row the eight 16-byte units are permuted by XOR with (row & 7)  (Swizzle<3,4,3>).  Chunk blocks
//   follow each other; each starts 1024-byte aligned.
//
// The residual stream image ("X image") of one 280-token sequence: 2 chunks x 288 rows (rows 280..287
// are zero) = 73,728 bytes.
#pragma once

#include <cuda_bf16.h>
#include <stdint.h>

#include "aft_internal.cuh"

namespace aft {

constexpr int kSPad = 288;                            // token rows padded to a multiple of 8 (and 16)
constexpr int kChunkRowBytes = 128;                   // 64 bf16
constexpr int kXChunkBytes = kSPad * kChunkRowBytes;  // 36,864
constexpr int kXImageBytes = 2 * kXChunkBytes;        // 73,728  (K = 128)

// byte offset of element (row, col) -- col a multiple of 8 -- inside an image whose chunk blocks
// have `rows` rows
__host__ __device__ __forceinline__ int image_offset(int row, int col, int rows) {
  const int chunk = col >> 6;
  const int unit = (col & 63) >> 3;
  return chunk * rows * kChunkRowBytes + row * kChunkRowBytes + (((unit ^ (row & 7)) << 4));
}

__host__ __device__ __forceinline__ int ximage_offset(int row, int col) { return image_offset(row, col, kSPad); }

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  // cvt.rn.bf16x2.f32 d, a, b : a -> upper half, b -> lower half
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

}  // namespace aft
