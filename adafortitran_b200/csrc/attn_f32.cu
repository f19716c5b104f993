// fp32 multi-head self-attention of the AFT_FP32 path (reference: torch _native_multi_head_attention as
// reached from src/models/blocks/encoders.py:69): per (sequence, head)  softmax(q k^T / sqrt(dh)) v.
// One CTA per (sequence, head); K and V of that head live in shared memory; one thread per query row,
// streaming over the 280 keys with a running max / sum (mathematically identical to the two-pass
// softmax; fp32 throughout, expf not __expf).
#include "aft_internal.cuh"

namespace aft {

namespace {

constexpr int kAttnThreads = 288;   // 9 warps >= 280 query rows

__global__ void __launch_bounds__(kAttnThreads)
attn_f32_kernel(const float* __restrict__ qkv, float* __restrict__ out) {
  extern __shared__ __align__(16) float sm[];
  float* Ks = sm;                 // [280][32]
  float* Vs = sm + kS * kDh;      // [280][32]
  const int tid = threadIdx.x;
  const int64_t seq = blockIdx.x >> 2;
  const int head = blockIdx.x & 3;
  const float* base = qkv + seq * (int64_t)kS * 3 * kD;

  for (int i = tid; i < kS * (kDh / 4); i += kAttnThreads) {
    const int j = i >> 3, c4 = (i & 7) * 4;
    *reinterpret_cast<float4*>(Ks + j * kDh + c4) =
        *reinterpret_cast<const float4*>(base + (int64_t)j * 3 * kD + kD + head * kDh + c4);
    *reinterpret_cast<float4*>(Vs + j * kDh + c4) =
        *reinterpret_cast<const float4*>(base + (int64_t)j * 3 * kD + 2 * kD + head * kDh + c4);
  }
  __syncthreads();
  if (tid >= kS) return;

  float q[kDh], o[kDh];
  const float scale = 0.17677669529663688110f;   // 1/sqrt(32), applied to q as torch's fast path does
#pragma unroll
  for (int c4 = 0; c4 < kDh; c4 += 4) {
    const float4 v = *reinterpret_cast<const float4*>(base + (int64_t)tid * 3 * kD + head * kDh + c4);
    q[c4] = v.x * scale; q[c4 + 1] = v.y * scale; q[c4 + 2] = v.z * scale; q[c4 + 3] = v.w * scale;
  }
#pragma unroll
  for (int c = 0; c < kDh; ++c) o[c] = 0.f;
  float m = -INFINITY, l = 0.f;
  for (int j = 0; j < kS; ++j) {
    float s = 0.f;
#pragma unroll
    for (int c4 = 0; c4 < kDh; c4 += 4) {
      const float4 k = *reinterpret_cast<const float4*>(Ks + j * kDh + c4);
      s = fmaf(q[c4], k.x, s); s = fmaf(q[c4 + 1], k.y, s); s = fmaf(q[c4 + 2], k.z, s); s = fmaf(q[c4 + 3], k.w, s);
    }
    const float mn = fmaxf(m, s);
    const float corr = expf(m - mn);     // exp(-inf) = 0 on the first key
    const float p = expf(s - mn);
    l = fmaf(l, corr, p);
#pragma unroll
    for (int c4 = 0; c4 < kDh; c4 += 4) {
      const float4 v = *reinterpret_cast<const float4*>(Vs + j * kDh + c4);
      o[c4] = fmaf(o[c4], corr, p * v.x); o[c4 + 1] = fmaf(o[c4 + 1], corr, p * v.y);
      o[c4 + 2] = fmaf(o[c4 + 2], corr, p * v.z); o[c4 + 3] = fmaf(o[c4 + 3], corr, p * v.w);
    }
    m = mn;
  }
  const float inv = 1.0f / l;
  float* op = out + (seq * kS + tid) * kD + head * kDh;
#pragma unroll
  for (int c4 = 0; c4 < kDh; c4 += 4)
    *reinterpret_cast<float4*>(op + c4) = make_float4(o[c4] * inv, o[c4 + 1] * inv, o[c4 + 2] * inv, o[c4 + 3] * inv);
}

}  // namespace

bool launch_attn_f32(const float* qkv, float* out, int64_t nseq, cudaStream_t st) {
  const size_t smem = 2 * kS * kDh * sizeof(float);   // 71,680
  if (cudaFuncSetAttribute(attn_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
    set_error("attn_f32: cannot opt in to %zu bytes of shared memory", smem);
    return false;
  }
  if (nseq <= 0) return true;
  attn_f32_kernel<<<(unsigned)(nseq * kH), kAttnThreads, smem, st>>>(qkv, out);
  count_launch();
  return check_launch("attn_f32_kernel");
}

}  // namespace aft
