// Head of the real-valued pass (reference src/models/blocks/encoders.py:70 + fortitran.py:225-231,180):
//   linear_2 (128 -> 6) -> Fold(3x2) -> + conv_enhanced -> ConvEnhancer(final_refiner) -> torch.complex
// One CTA per SAMPLE: it runs the real and the imaginary sequence back to back and then emits the
// estimate as interleaved complex64 with fully coalesced float2 stores.
#include "conv_simt.cuh"

namespace aft {

namespace {

constexpr int kHeadSmemFloats = kConvSmemFloats + 2 * kPix;

__global__ void __launch_bounds__(kConvThreads, 1)
head_kernel(HeadPack p, const float* __restrict__ h, const float* __restrict__ enh, float2* __restrict__ out) {
  extern __shared__ __align__(16) float smem[];
  const ConvSmem cs = carve_conv_smem(smem);
  float* res = smem + kConvSmemFloats;   // [2][1680]
  const int tid = threadIdx.x, nt = blockDim.x;
  const int warp = tid >> 5, lane = tid & 31, nwarp = nt >> 5;
  const int64_t sample = blockIdx.x;

  conv_prepare(cs, p.refine);
  // linear_2 rows, this lane's 4 columns of each of the 6 output features
  float4 w2[kPatchLen];
#pragma unroll
  for (int f = 0; f < kPatchLen; ++f) w2[f] = *reinterpret_cast<const float4*>(p.l2_w + f * kD + lane * 4);
  const float b2 = lane < kPatchLen ? p.l2_b[lane] : 0.f;
  __syncthreads();

  for (int part = 0; part < 2; ++part) {
    const int64_t seq = 2 * sample + part;
    const float* hs = h + seq * (int64_t)kS * kD;
    const float* es = enh + seq * (int64_t)kPix;
    // token -> 6 pixel residuals, folded back to the grid and added to conv_enhanced
    for (int t = warp; t < kS; t += nwarp) {
      const float4 hv = *reinterpret_cast<const float4*>(hs + t * kD + lane * 4);
      float acc[kPatchLen];
#pragma unroll
      for (int f = 0; f < kPatchLen; ++f)
        acc[f] = hv.x * w2[f].x + hv.y * w2[f].y + hv.z * w2[f].z + hv.w * w2[f].w;
#pragma unroll
      for (int off = 16; off > 0; off >>= 1)
#pragma unroll
        for (int f = 0; f < kPatchLen; ++f) acc[f] += __shfl_xor_sync(0xffffffffu, acc[f], off);
      if (lane < kPatchLen) {
        float v = acc[0];
#pragma unroll
        for (int f = 1; f < kPatchLen; ++f) v = (lane == f) ? acc[f] : v;
        const int pi = t / kTokW, pj = t - pi * kTokW;
        const int a = lane / kPatchW, b = lane - a * kPatchW;
        const int r = kPatchH * pi + a, c = kPatchW * pj + b;
        cs.in[(r + 1) * kPW + c + 1] = v + b2 + es[r * kGridW + c];
      }
    }
    __syncthreads();
    conv_stack(cs, res + part * kPix);
  }
  for (int i = tid; i < kPix; i += nt) out[sample * kPix + i] = make_float2(res[i], res[kPix + i]);
}

}  // namespace

bool launch_head(const HeadPack& p, const float* h, const float* enh, float2* out, int64_t nsamples, cudaStream_t st) {
  const size_t smem = kHeadSmemFloats * sizeof(float);
  if (cudaFuncSetAttribute(head_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
    set_error("head: cannot opt in to %zu bytes of shared memory: %s", smem, cudaGetErrorString(cudaGetLastError()));
    return false;
  }
  if (nsamples <= 0) return true;
  head_kernel<<<(unsigned)nsamples, kConvThreads, smem, st>>>(p, h, enh, out);
  count_launch();
  return check_launch("head_kernel");
}

}  // namespace aft
