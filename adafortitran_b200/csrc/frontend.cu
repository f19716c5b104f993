// Frontend of one real-valued pass (reference src/models/fortitran.py:199-217 + encoders.py:67-68):
//   pilots (re or im) -> Linear(24,1680) -> ConvEnhancer -> Unfold(3x2) [+ ChannelAdapter] -> linear_1 + pos
// One CTA per real-valued sequence (sequence n = 2*sample + {0: real, 1: imag}); everything between the
// 24 input floats and the 280x128 token matrix stays in shared memory.
#include "conv_simt.cuh"
#include "tc_layout.cuh"

namespace aft {

namespace {

constexpr int kFrontSmemFloats = kConvSmemFloats + kPix /*enh plane*/ + 32 /*x*/;
// after the conv stack the a1 region (8*kPlane = 15616 floats) is dead and is reused for:
//   tok_in [280][12] (3360) | z [3][560] (1680) | l1_wt [12][128] (1536) | hidden [2][64] (128)
static_assert(3360 + 1680 + 1536 + 128 <= 8 * kPlane, "frontend scratch must fit in the a1 region");

__global__ void __launch_bounds__(kConvThreads, 1)
frontend_kernel(FrontPack p, const float2* __restrict__ pilots, const float* __restrict__ snr,
                const float* __restrict__ ds, const float* __restrict__ dop, float* __restrict__ enh_out,
                float* __restrict__ h_out, __nv_bfloat16* __restrict__ hb_out) {
  extern __shared__ __align__(16) float smem[];
  const ConvSmem cs = carve_conv_smem(smem);
  float* enh = smem + kConvSmemFloats;   // [1680]
  float* x = enh + kPix;                 // [24]
  const int tid = threadIdx.x, nt = blockDim.x;
  const int64_t seq = blockIdx.x;
  const int64_t sample = seq >> 1;
  const int part = seq & 1;

  conv_prepare(cs, p.enh);
  if (tid < kPilots) {
    const float2 v = pilots[sample * kPilots + tid];
    x[tid] = part ? v.y : v.x;
  }
  __syncthreads();

  // Stage 1: upsample, y[pix] = sum_k W[pix][k] x[k] + b[pix]  (fortitran.py:203), into the padded plane
  for (int pix = tid; pix < kPix; pix += nt) {
    float acc = p.up_b[pix];
#pragma unroll
    for (int k = 0; k < kPilots; ++k) acc = fmaf(p.up_wt[k * kPix + pix], x[k], acc);
    const int r = pix / kGridW, c = pix - r * kGridW;
    cs.in[(r + 1) * kPW + c + 1] = acc;
  }
  __syncthreads();

  // Stage 2: ConvEnhancer (fortitran.py:209)
  conv_stack(cs, enh);
  for (int i = tid; i < kPix; i += nt) enh_out[seq * kPix + i] = enh[i];

  // Stage 3/4: tokens = [patch(6) | adaptive(6)]  (fortitran.py:212-217)
  float* tok = cs.a1;             // [280][in_dim]
  float* z = tok + 3360;          // [3][560]
  float* w1t = z + 1680;          // [in_dim][128]
  float* hid = w1t + 1536;        // [2][64]
  const int in_dim = p.in_dim;
  for (int i = tid; i < kS * kPatchLen; i += nt) {
    const int t = i / kPatchLen, f = i - t * kPatchLen;
    const int pi = t / kTokW, pj = t - pi * kTokW;
    const int a = f / kPatchW, b = f - a * kPatchW;
    tok[t * in_dim + f] = enh[(kPatchH * pi + a) * kGridW + kPatchW * pj + b];
  }
  for (int i = tid; i < in_dim * kD; i += nt) w1t[i] = p.l1_wt[i];
  if (p.adaptive) {
    // ChannelAdapter (channel_adaptivity.py:59-63): three scalar -> h1 -> h2 -> 560 MLPs, raw inputs.
    const float cond[3] = {snr[sample], ds[sample], dop[sample]};
    for (int m = 0; m < 3; ++m) {
      const MlpPack& mp = p.mlp[m];
      if (tid < p.h1) hid[tid] = fmaxf(fmaf(mp.w0[tid], cond[m], mp.b0[tid]), 0.f);
      __syncthreads();
      if (tid < p.h2) {
        float acc = mp.b1[tid];
        for (int k = 0; k < p.h1; ++k) acc = fmaf(mp.w1[tid * p.h1 + k], hid[k], acc);
        hid[64 + tid] = fmaxf(acc, 0.f);
      }
      __syncthreads();
      for (int j = tid; j < 2 * kS; j += nt) {
        float acc = mp.b2[j];
        for (int k = 0; k < p.h2; ++k) acc = fmaf(mp.w2t[k * 2 * kS + j], hid[64 + k], acc);
        z[m * 2 * kS + j] = acc;
      }
      __syncthreads();
    }
    // token t, feature 6 + 2m + e  <-  z_m[2t + e]
    for (int i = tid; i < kS * kAda; i += nt) {
      const int t = i / kAda, f = i - t * kAda;
      const int m = f >> 1, e = f & 1;
      tok[t * in_dim + kPatchLen + f] = z[m * 2 * kS + 2 * t + e];
    }
  }
  __syncthreads();

  // Stage 5a: h = tok . W1^T + b1 + pos  (encoders.py:67-68).  item = (token, 8 consecutive columns)
  for (int it = tid; it < kS * (kD / 8); it += nt) {
    const int t = it >> 4, c8 = (it & 15) * 8;
    const float4 pa = *reinterpret_cast<const float4*>(p.posb + t * kD + c8);
    const float4 pb = *reinterpret_cast<const float4*>(p.posb + t * kD + c8 + 4);
    float acc[8] = {pa.x, pa.y, pa.z, pa.w, pb.x, pb.y, pb.z, pb.w};
    for (int k = 0; k < in_dim; ++k) {
      const float a = tok[t * in_dim + k];
      const float4 wa = *reinterpret_cast<const float4*>(w1t + k * kD + c8);
      const float4 wb = *reinterpret_cast<const float4*>(w1t + k * kD + c8 + 4);
      acc[0] = fmaf(a, wa.x, acc[0]); acc[1] = fmaf(a, wa.y, acc[1]);
      acc[2] = fmaf(a, wa.z, acc[2]); acc[3] = fmaf(a, wa.w, acc[3]);
      acc[4] = fmaf(a, wb.x, acc[4]); acc[5] = fmaf(a, wb.y, acc[5]);
      acc[6] = fmaf(a, wb.z, acc[6]); acc[7] = fmaf(a, wb.w, acc[7]);
    }
    if (h_out != nullptr) {
      float* hp = h_out + (seq * kS + t) * kD + c8;
      *reinterpret_cast<float4*>(hp) = make_float4(acc[0], acc[1], acc[2], acc[3]);
      *reinterpret_cast<float4*>(hp + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
    }
    if (hb_out != nullptr) {
      // bf16 operand image of the residual stream (tc_layout.cuh): one 16-byte chunk per item
      uint4 pk;
      pk.x = pack_bf16x2(acc[0], acc[1]); pk.y = pack_bf16x2(acc[2], acc[3]);
      pk.z = pack_bf16x2(acc[4], acc[5]); pk.w = pack_bf16x2(acc[6], acc[7]);
      char* base = reinterpret_cast<char*>(hb_out) + seq * (int64_t)kXImageBytes;
      *reinterpret_cast<uint4*>(base + ximage_offset(t, c8)) = pk;
    }
  }
  // rows 280..287 of the image are zero padding
  if (hb_out != nullptr) {
    char* base = reinterpret_cast<char*>(hb_out) + seq * (int64_t)kXImageBytes;
    for (int it = tid; it < (kSPad - kS) * (kD / 8); it += nt) {
      const int t = kS + (it >> 4), c8 = (it & 15) * 8;
      *reinterpret_cast<uint4*>(base + ximage_offset(t, c8)) = make_uint4(0, 0, 0, 0);
    }
  }
}

}  // namespace

bool launch_frontend(const FrontPack& p, const float2* pilots, const float* snr, const float* ds, const float* dop,
                     float* enh, float* h, __nv_bfloat16* hb, int64_t nsamples, cudaStream_t st) {
  const size_t smem = kFrontSmemFloats * sizeof(float);
  if (cudaFuncSetAttribute(frontend_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
    set_error("frontend: cannot opt in to %zu bytes of shared memory: %s", smem, cudaGetErrorString(cudaGetLastError()));
    return false;
  }
  if (nsamples <= 0) return true;
  frontend_kernel<<<(unsigned)(2 * nsamples), kConvThreads, smem, st>>>(p, pilots, snr, ds, dop, enh, h, hb);
  count_launch();
  return check_launch("frontend_kernel");
}

}  // namespace aft
