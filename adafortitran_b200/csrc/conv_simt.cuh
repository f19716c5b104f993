// ConvEnhancer (reference src/models/blocks/enhancers.py:12-20) on CUDA cores, one image per CTA,
// everything resident in shared memory.  Used by the frontend and the head kernels of both
// precisions (the conv stacks stay in fp32; they are 4.5 % of the FLOPs).
//
// Layout: every activation plane is stored zero-padded as (120+2) x (14+2) floats so the 3x3 taps
// need no bounds checks.  The 32-channel intermediate never exists in full: conv2 (8->32) and
// conv3 (32->8) are fused over strips of kStripRows output rows, with the 32-channel strip (plus a
// one-row halo each side) in shared memory.
#pragma once

#include "aft_internal.cuh"

namespace aft {

constexpr int kConvThreads = 416;               // 13 warps: balances the two strip phases (see below)
constexpr int kStripRows = 12;                  // 120 / 12 = 10 strips
constexpr int kMidRows = kStripRows + 2;        // with halo
constexpr int kMidPlane = kMidRows * kPW;       // 224

// shared-memory footprint of one conv stack, in floats
constexpr int kConvWeightsFloats = 72 + 2304 + 2304 + 72 + 8 + 32 + 8 + 8;  // 4808 (b3 padded to 8)
constexpr int kConvSmemFloats = kPlane /*in*/ + 8 * kPlane /*a1*/ + 32 * kMidPlane /*mid*/ + 8 * kPlane /*a3*/ +
                                kConvWeightsFloats;

struct ConvSmem {
  float* in;    // [kPlane]        1-channel padded input (caller fills the interior; border must be 0)
  float* a1;    // [8][kPlane]
  float* mid;   // [32][kMidPlane]
  float* a3;    // [8][kPlane]
  float* w;     // packed weights: w0[72] w1[2304] w2[2304] w3[72] b0[8] b1[32] b2[8] b3[8]
};

__device__ __forceinline__ ConvSmem carve_conv_smem(float* base) {
  ConvSmem s;
  s.in = base;
  s.a1 = s.in + kPlane;
  s.mid = s.a1 + 8 * kPlane;
  s.a3 = s.mid + 32 * kMidPlane;
  s.w = s.a3 + 8 * kPlane;
  return s;
}

// Zero the padded planes (borders stay zero afterwards: kernels only ever write interiors) and
// stage the weights.  Must be followed by __syncthreads() before conv_stack().
__device__ __forceinline__ void conv_prepare(const ConvSmem& s, const ConvPack& p) {
  const int tid = threadIdx.x;
  for (int i = tid; i < kPlane + 8 * kPlane + 32 * kMidPlane + 8 * kPlane; i += blockDim.x) s.in[i] = 0.f;
  float* w = s.w;
  for (int i = tid; i < 72; i += blockDim.x) w[i] = p.w0[i];
  for (int i = tid; i < 2304; i += blockDim.x) w[72 + i] = p.w1[i];
  for (int i = tid; i < 2304; i += blockDim.x) w[72 + 2304 + i] = p.w2[i];
  for (int i = tid; i < 72; i += blockDim.x) w[72 + 4608 + i] = p.w3[i];
  if (tid < 8) w[4752 + tid] = p.b0[tid];
  if (tid < 32) w[4760 + tid] = p.b1[tid];
  if (tid < 8) w[4792 + tid] = p.b2[tid];
  if (tid == 0) w[4800] = p.b3[0];
}

// Runs the 4-layer stack on s.in and writes the 1-channel result, unpadded, to out[1680] (shared or
// global).  All threads of the CTA must call it; ends with a __syncthreads().
__device__ __forceinline__ void conv_stack(const ConvSmem& s, float* out) {
  const int tid = threadIdx.x, nt = blockDim.x;
  const float* w0 = s.w;
  const float* w1 = s.w + 72;
  const float* w2 = s.w + 72 + 2304;
  const float* w3 = s.w + 72 + 4608;
  const float* b0 = s.w + 4752;
  const float* b1 = s.w + 4760;
  const float* b2 = s.w + 4792;
  const float b3 = s.w[4800];

  // ---- conv1: 1 -> 8, ReLU.  thread <-> pixel, all 8 output channels in registers ----
  for (int p = tid; p < kPix; p += nt) {
    const int r = p / kGridW, c = p - r * kGridW;
    float acc[8];
#pragma unroll
    for (int o = 0; o < 8; ++o) acc[o] = b0[o];
#pragma unroll
    for (int dy = 0; dy < 3; ++dy)
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        const float a = s.in[(r + dy) * kPW + c + dx];
        const float4 wa = *reinterpret_cast<const float4*>(w0 + (dy * 3 + dx) * 8);
        const float4 wb = *reinterpret_cast<const float4*>(w0 + (dy * 3 + dx) * 8 + 4);
        acc[0] = fmaf(a, wa.x, acc[0]); acc[1] = fmaf(a, wa.y, acc[1]);
        acc[2] = fmaf(a, wa.z, acc[2]); acc[3] = fmaf(a, wa.w, acc[3]);
        acc[4] = fmaf(a, wb.x, acc[4]); acc[5] = fmaf(a, wb.y, acc[5]);
        acc[6] = fmaf(a, wb.z, acc[6]); acc[7] = fmaf(a, wb.w, acc[7]);
      }
#pragma unroll
    for (int o = 0; o < 8; ++o) s.a1[o * kPlane + (r + 1) * kPW + c + 1] = fmaxf(acc[o], 0.f);
  }
  __syncthreads();

  // ---- conv2 (8 -> 32, ReLU) + conv3 (32 -> 8, ReLU), fused over strips ----
  for (int r0 = 0; r0 < kGridH; r0 += kStripRows) {
    // conv2 on rows r0-1 .. r0+kStripRows (halo rows outside the image are zero padding for conv3)
    for (int it = tid; it < 4 * kMidRows * kGridW; it += nt) {
      const int cg = it / (kMidRows * kGridW);
      const int pix = it - cg * (kMidRows * kGridW);
      const int lr = pix / kGridW, c = pix - lr * kGridW;
      const int rr = r0 - 1 + lr;
      float acc[8];
      if (rr >= 0 && rr < kGridH) {
#pragma unroll
        for (int o = 0; o < 8; ++o) acc[o] = b1[cg * 8 + o];
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          const int dy = t / 3, dx = t - dy * 3;
          const float* ap = s.a1 + (rr + dy) * kPW + c + dx;
          const float* wp = w1 + t * 256 + cg * 8;
#pragma unroll
          for (int ci = 0; ci < 8; ++ci) {
            const float a = ap[ci * kPlane];
            const float4 wa = *reinterpret_cast<const float4*>(wp + ci * 32);
            const float4 wb = *reinterpret_cast<const float4*>(wp + ci * 32 + 4);
            acc[0] = fmaf(a, wa.x, acc[0]); acc[1] = fmaf(a, wa.y, acc[1]);
            acc[2] = fmaf(a, wa.z, acc[2]); acc[3] = fmaf(a, wa.w, acc[3]);
            acc[4] = fmaf(a, wb.x, acc[4]); acc[5] = fmaf(a, wb.y, acc[5]);
            acc[6] = fmaf(a, wb.z, acc[6]); acc[7] = fmaf(a, wb.w, acc[7]);
          }
        }
#pragma unroll
        for (int o = 0; o < 8; ++o) acc[o] = fmaxf(acc[o], 0.f);
      } else {
#pragma unroll
        for (int o = 0; o < 8; ++o) acc[o] = 0.f;
      }
#pragma unroll
      for (int o = 0; o < 8; ++o) s.mid[(cg * 8 + o) * kMidPlane + lr * kPW + c + 1] = acc[o];
    }
    __syncthreads();
    // conv3 on rows r0 .. r0+kStripRows-1; thread <-> (4 output channels, pixel)
    for (int it = tid; it < 2 * kStripRows * kGridW; it += nt) {
      const int ch = it / (kStripRows * kGridW);
      const int pix = it - ch * (kStripRows * kGridW);
      const int lr = pix / kGridW, c = pix - lr * kGridW;
      float acc[4];
#pragma unroll
      for (int o = 0; o < 4; ++o) acc[o] = b2[ch * 4 + o];
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const int dy = t / 3, dx = t - dy * 3;
        const float* ap = s.mid + (lr + dy) * kPW + c + dx;
        const float* wp = w2 + t * 256 + ch * 4;
#pragma unroll 8
        for (int ci = 0; ci < 32; ++ci) {
          const float a = ap[ci * kMidPlane];
          const float4 wa = *reinterpret_cast<const float4*>(wp + ci * 8);
          acc[0] = fmaf(a, wa.x, acc[0]); acc[1] = fmaf(a, wa.y, acc[1]);
          acc[2] = fmaf(a, wa.z, acc[2]); acc[3] = fmaf(a, wa.w, acc[3]);
        }
      }
      const int r = r0 + lr;
#pragma unroll
      for (int o = 0; o < 4; ++o) s.a3[(ch * 4 + o) * kPlane + (r + 1) * kPW + c + 1] = fmaxf(acc[o], 0.f);
    }
    __syncthreads();
  }

  // ---- conv4: 8 -> 1 (no activation) ----
  for (int p = tid; p < kPix; p += nt) {
    const int r = p / kGridW, c = p - r * kGridW;
    float acc = b3;
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const int dy = t / 3, dx = t - dy * 3;
      const float* ap = s.a3 + (r + dy) * kPW + c + dx;
#pragma unroll
      for (int ci = 0; ci < 8; ++ci) acc = fmaf(ap[ci * kPlane], w3[t * 8 + ci], acc);
    }
    out[p] = acc;
  }
  __syncthreads();
}

}  // namespace aft
