// AFT_FP32 path for grids other than the reference default (e.g. BASELINE config 5: 3276 x 14 subcarriers x symbols,
// 7644 tokens).  Same arithmetic as the specialised fp32 kernels (frontend.cu / attn_f32.cu / head.cu), with every
// grid-dependent extent a runtime value and the activations in global memory instead of one CTA's shared memory:
//   pilots -> Linear(P, H*W) -> ConvEnhancer -> Unfold(ph x pw) [+ ChannelAdapter] -> linear_1 + pos      (generic_frontend)
//   softmax(q k^T / sqrt(dh)) v with a streaming softmax over key tiles                                   (generic_attention)
//   linear_2 -> Fold -> + enhanced -> ConvEnhancer -> complex64                                           (generic_head)
// The encoder GEMMs are the shape-generic ones of gemm_f32.cu (model_dim 128, ff 256).  Plain FMA, erff, expf: a parity
// path, not a tuned one.  Reference: src/models/fortitran.py:184-233, blocks/*.py.
#include "aft_internal.cuh"

namespace aft {

namespace {

// x[seq][k] = re / im of pilots[sample][k], seq = 2 * sample + part
__global__ void g_split_pilots(const float2* __restrict__ pilots, float* __restrict__ x, int64_t nsamples, int P) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nsamples * P) return;
  const int64_t s = i / P;
  const int k = (int)(i - s * P);
  const float2 v = pilots[i];
  x[(2 * s) * P + k] = v.x;
  x[(2 * s + 1) * P + k] = v.y;
}

// 3x3 cross-correlation, zero padding 1, NCHW planes in global memory: in [nseq][CIN][H*W] -> out [nseq][COUT][H*W].
// Weights tap-major [9][CIN][COUT] (ConvPack layout), staged in shared memory.  One thread per output pixel.
template <int CIN, int COUT, bool RELU>
__global__ void __launch_bounds__(256)
g_conv3x3(const float* __restrict__ in, float* __restrict__ out, const float* __restrict__ w, const float* __restrict__ bias, int H, int W) {
  __shared__ float ws[9 * CIN * COUT];
  __shared__ float bs[COUT];
  for (int i = threadIdx.x; i < 9 * CIN * COUT; i += blockDim.x) ws[i] = w[i];
  if (threadIdx.x < COUT) bs[threadIdx.x] = bias[threadIdx.x];
  __syncthreads();
  const int pix = H * W;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= pix) return;
  const int64_t seq = blockIdx.y;
  const int r = p / W, c = p - r * W;
  const float* ip = in + seq * (int64_t)CIN * pix;
  float acc[COUT];
#pragma unroll
  for (int o = 0; o < COUT; ++o) acc[o] = bs[o];
  for (int ci = 0; ci < CIN; ++ci) {
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const int rr = r + t / 3 - 1, cc = c + t % 3 - 1;
      if (rr < 0 || rr >= H || cc < 0 || cc >= W) continue;
      const float v = ip[(int64_t)ci * pix + rr * W + cc];
      const float* wp = ws + (t * CIN + ci) * COUT;
#pragma unroll
      for (int o = 0; o < COUT; ++o) acc[o] = fmaf(v, wp[o], acc[o]);
    }
  }
  float* op = out + seq * (int64_t)COUT * pix;
#pragma unroll
  for (int o = 0; o < COUT; ++o) op[(int64_t)o * pix + p] = RELU ? fmaxf(acc[o], 0.f) : acc[o];
}

// ChannelAdapter (channel_adaptivity.py:59-63): z[sample][m][j], j < 2 S, for the three conditions m
__global__ void __launch_bounds__(256)
g_adapter(FrontPack p, const float* __restrict__ snr, const float* __restrict__ ds, const float* __restrict__ dop, float* __restrict__ z, int S2) {
  __shared__ float hid[2 * kMaxAdaHidden];
  const int64_t sample = blockIdx.x;
  const int m = blockIdx.y;
  const MlpPack& mp = p.mlp[m];
  const float cond = m == 0 ? snr[sample] : (m == 1 ? ds[sample] : dop[sample]);
  if ((int)threadIdx.x < p.h1) hid[threadIdx.x] = fmaxf(fmaf(mp.w0[threadIdx.x], cond, mp.b0[threadIdx.x]), 0.f);
  __syncthreads();
  if ((int)threadIdx.x < p.h2) {
    float acc = mp.b1[threadIdx.x];
    for (int k = 0; k < p.h1; ++k) acc = fmaf(mp.w1[threadIdx.x * p.h1 + k], hid[k], acc);
    hid[kMaxAdaHidden + threadIdx.x] = fmaxf(acc, 0.f);
  }
  __syncthreads();
  float* zo = z + (sample * 3 + m) * (int64_t)S2;
  for (int j = threadIdx.x; j < S2; j += blockDim.x) {
    float acc = mp.b2[j];
    for (int k = 0; k < p.h2; ++k) acc = fmaf(mp.w2t[(int64_t)k * S2 + j], hid[kMaxAdaHidden + k], acc);
    zo[j] = acc;
  }
}

// tokens (fortitran.py:212-217) and linear_1 + positional table (encoders.py:67-68): one thread per (row, 4 columns)
__global__ void __launch_bounds__(256)
g_tokens_linear1(FrontPack p, const float* __restrict__ enh, const float* __restrict__ z, float* __restrict__ h, int64_t nseq, int S, int W,
                 int ph, int pw) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nseq * S * (kD / 4)) return;
  const int c4 = (int)(i % (kD / 4)) * 4;
  const int64_t row = i / (kD / 4);
  const int64_t seq = row / S;
  const int t = (int)(row - seq * S);
  const int tokW = W / pw, pl = ph * pw;
  const int pi = t / tokW, pj = t - pi * tokW;
  const int pix = (int)((int64_t)(S / tokW) * ph) * W;
  const float4 pb = *reinterpret_cast<const float4*>(p.posb + (int64_t)t * kD + c4);
  float acc[4] = {pb.x, pb.y, pb.z, pb.w};
  for (int f = 0; f < p.in_dim; ++f) {
    float a;
    if (f < pl) {
      const int ra = f / pw, cb = f - ra * pw;
      a = enh[seq * pix + (ph * pi + ra) * W + pw * pj + cb];
    } else {
      const int g = f - pl, m = g >> 1, e = g & 1;          // token t takes outputs 2t, 2t+1 of condition m
      a = z[((seq >> 1) * 3 + m) * (int64_t)(2 * S) + 2 * t + e];
    }
    const float4 wv = *reinterpret_cast<const float4*>(p.l1_wt + f * kD + c4);
    acc[0] = fmaf(a, wv.x, acc[0]); acc[1] = fmaf(a, wv.y, acc[1]); acc[2] = fmaf(a, wv.z, acc[2]); acc[3] = fmaf(a, wv.w, acc[3]);
  }
  *reinterpret_cast<float4*>(h + row * kD + c4) = make_float4(acc[0], acc[1], acc[2], acc[3]);
}

// fp32 attention for any sequence length: CTA = (query block of 128 rows, head, sequence), one thread per query row,
// K / V of the head streamed through shared memory in tiles of 64 keys, running max / sum (as attn_f32.cu).
constexpr int kQB = 128, kKT = 64;
__global__ void __launch_bounds__(kQB)
g_attention(const float* __restrict__ qkv, float* __restrict__ out, int S) {
  __shared__ __align__(16) float Ks[kKT * kDh];
  __shared__ __align__(16) float Vs[kKT * kDh];
  const int tid = threadIdx.x;
  const int head = blockIdx.y;
  const int64_t seq = blockIdx.z;
  const int row = blockIdx.x * kQB + tid;
  const float* base = qkv + seq * (int64_t)S * 3 * kD;
  const bool live = row < S;
  float q[kDh], o[kDh];
  const float scale = 0.17677669529663688110f;   // 1/sqrt(32), applied to q as torch's fast path does
#pragma unroll
  for (int c4 = 0; c4 < kDh; c4 += 4) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (live) v = *reinterpret_cast<const float4*>(base + (int64_t)row * 3 * kD + head * kDh + c4);
    q[c4] = v.x * scale; q[c4 + 1] = v.y * scale; q[c4 + 2] = v.z * scale; q[c4 + 3] = v.w * scale;
  }
#pragma unroll
  for (int c = 0; c < kDh; ++c) o[c] = 0.f;
  float m = -INFINITY, l = 0.f;
  for (int k0 = 0; k0 < S; k0 += kKT) {
    const int nk = S - k0 < kKT ? S - k0 : kKT;
    __syncthreads();
    for (int i = tid; i < nk * (kDh / 4); i += kQB) {
      const int j = i >> 3, c4 = (i & 7) * 4;
      *reinterpret_cast<float4*>(Ks + j * kDh + c4) = *reinterpret_cast<const float4*>(base + (int64_t)(k0 + j) * 3 * kD + kD + head * kDh + c4);
      *reinterpret_cast<float4*>(Vs + j * kDh + c4) = *reinterpret_cast<const float4*>(base + (int64_t)(k0 + j) * 3 * kD + 2 * kD + head * kDh + c4);
    }
    __syncthreads();
    for (int j = 0; j < nk; ++j) {
      float s = 0.f;
#pragma unroll
      for (int c4 = 0; c4 < kDh; c4 += 4) {
        const float4 k = *reinterpret_cast<const float4*>(Ks + j * kDh + c4);
        s = fmaf(q[c4], k.x, s); s = fmaf(q[c4 + 1], k.y, s); s = fmaf(q[c4 + 2], k.z, s); s = fmaf(q[c4 + 3], k.w, s);
      }
      const float mn = fmaxf(m, s);
      const float corr = expf(m - mn);
      const float pexp = expf(s - mn);
      l = fmaf(l, corr, pexp);
#pragma unroll
      for (int c4 = 0; c4 < kDh; c4 += 4) {
        const float4 v = *reinterpret_cast<const float4*>(Vs + j * kDh + c4);
        o[c4] = fmaf(o[c4], corr, pexp * v.x); o[c4 + 1] = fmaf(o[c4 + 1], corr, pexp * v.y);
        o[c4 + 2] = fmaf(o[c4 + 2], corr, pexp * v.z); o[c4 + 3] = fmaf(o[c4 + 3], corr, pexp * v.w);
      }
      m = mn;
    }
  }
  if (!live) return;
  const float inv = 1.0f / l;
  float* op = out + (seq * S + row) * kD + head * kDh;
#pragma unroll
  for (int c4 = 0; c4 < kDh; c4 += 4)
    *reinterpret_cast<float4*>(op + c4) = make_float4(o[c4] * inv, o[c4 + 1] * inv, o[c4 + 2] * inv, o[c4 + 3] * inv);
}

// linear_2 (encoders.py:70), Fold (patch_processors.py:69-71) and the residual (fortitran.py:228): one warp per token
__global__ void __launch_bounds__(256)
g_linear2_fold(HeadPack p, const float* __restrict__ h, const float* __restrict__ enh, float* __restrict__ y, int64_t nseq, int S, int W, int ph,
               int pw) {
  const int lane = threadIdx.x & 31;
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= nseq * S) return;
  const int64_t seq = row / S;
  const int t = (int)(row - seq * S);
  const int tokW = W / pw, pl = ph * pw;
  const int pi = t / tokW, pj = t - pi * tokW;
  const int pix = (S / tokW) * ph * W;
  const float4 hv = *reinterpret_cast<const float4*>(h + row * kD + lane * 4);
  for (int f = 0; f < pl; ++f) {
    const float4 wv = *reinterpret_cast<const float4*>(p.l2_w + f * kD + lane * 4);
    float acc = hv.x * wv.x + hv.y * wv.y + hv.z * wv.z + hv.w * wv.w;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if (lane == 0) {
      const int ra = f / pw, cb = f - ra * pw;
      const int64_t idx = seq * pix + (ph * pi + ra) * W + pw * pj + cb;
      y[idx] = acc + p.l2_b[f] + enh[idx];
    }
  }
}

// torch.complex (fortitran.py:180)
__global__ void g_interleave(const float* __restrict__ res, float2* __restrict__ out, int64_t nsamples, int pix) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nsamples * pix) return;
  const int64_t s = i / pix;
  const int p = (int)(i - s * pix);
  out[i] = make_float2(res[(2 * s) * pix + p], res[(2 * s + 1) * pix + p]);
}

inline unsigned nblk(int64_t n, int per = 256) { return (unsigned)((n + per - 1) / per); }

bool conv_stack(const ConvPack& cp, const float* in, float* a, float* b, float* out, int64_t nseq, int H, int W, cudaStream_t st) {
  const dim3 grid(nblk((int64_t)H * W), (unsigned)nseq);
  g_conv3x3<1, 8, true><<<grid, 256, 0, st>>>(in, a, cp.w0, cp.b0, H, W);
  g_conv3x3<8, 32, true><<<grid, 256, 0, st>>>(a, b, cp.w1, cp.b1, H, W);
  g_conv3x3<32, 8, true><<<grid, 256, 0, st>>>(b, a, cp.w2, cp.b2, H, W);
  g_conv3x3<8, 1, false><<<grid, 256, 0, st>>>(a, out, cp.w3, cp.b3, H, W);
  count_launch(4);
  return check_launch("generic conv stack");
}

}  // namespace

// scratch (floats): x [nseq*P] | u [nseq*pix] | a [nseq*8*pix] | b [nseq*32*pix] | z [B*3*2S]
size_t generic_front_scratch_floats(int64_t nsamples, int P, int pix, int S) {
  const size_t nseq = 2 * (size_t)nsamples;
  return nseq * P + nseq * pix + nseq * 8 * (size_t)pix + nseq * 32 * (size_t)pix + (size_t)nsamples * 3 * 2 * S;
}

bool launch_generic_frontend(const FrontPack& p, const float2* pilots, const float* snr, const float* ds, const float* dop, float* enh,
                             float* h, float* scratch, int64_t nsamples, int H, int W, int P, int ph, int pw, cudaStream_t st) {
  if (nsamples <= 0) return true;
  const int64_t nseq = 2 * nsamples;
  const int pix = H * W, S = (H / ph) * (W / pw);
  if (nseq > 65535) { set_error("generic frontend: chunk too large"); return false; }
  float* x = scratch;
  float* u = x + nseq * P;
  float* a = u + nseq * pix;
  float* b = a + nseq * 8 * (int64_t)pix;
  float* z = b + nseq * 32 * (int64_t)pix;
  g_split_pilots<<<nblk(nsamples * P), 256, 0, st>>>(pilots, x, nsamples, P);
  count_launch();
  // generic fp32 linear: up_wt holds the torch layout [pix][P] here (see aft_api.cu)
  // (large extents: the tiled GEMM; the LinearEstimator kernel of aux.cu is built for a few dozen inputs)
  if ((int64_t)P * pix > (1 << 20) ? !launch_gemm_f32_any(x, p.up_wt, p.up_b, u, nseq, pix, P, st) : !launch_linear(p.up_wt, p.up_b, x, u, nseq, P, pix, st))
    return false;
  if (!conv_stack(p.enh, u, a, b, enh, nseq, H, W, st)) return false;
  if (p.adaptive) {
    g_adapter<<<dim3((unsigned)nsamples, 3), 256, 0, st>>>(p, snr, ds, dop, z, 2 * S);
    count_launch();
  }
  g_tokens_linear1<<<nblk(nseq * S * (kD / 4)), 256, 0, st>>>(p, enh, z, h, nseq, S, W, ph, pw);
  count_launch();
  return check_launch("generic frontend");
}

bool launch_generic_attention(const float* qkv, float* out, int64_t nseq, int S, cudaStream_t st) {
  if (nseq <= 0) return true;
  if (nseq > 65535) { set_error("generic attention: chunk too large"); return false; }
  g_attention<<<dim3((unsigned)((S + kQB - 1) / kQB), kH, (unsigned)nseq), kQB, 0, st>>>(qkv, out, S);
  count_launch();
  return check_launch("generic attention");
}

// scratch (floats): y [nseq*pix] | a [nseq*8*pix] | b [nseq*32*pix] | res [nseq*pix]
size_t generic_head_scratch_floats(int64_t nsamples, int pix) {
  const size_t nseq = 2 * (size_t)nsamples;
  return nseq * pix * 2 + nseq * 8 * (size_t)pix + nseq * 32 * (size_t)pix;
}

bool launch_generic_head(const HeadPack& p, const float* h, const float* enh, float2* out, float* scratch, int64_t nsamples, int H, int W,
                         int ph, int pw, cudaStream_t st) {
  if (nsamples <= 0) return true;
  const int64_t nseq = 2 * nsamples;
  const int pix = H * W, S = (H / ph) * (W / pw);
  float* y = scratch;
  float* a = y + nseq * pix;
  float* b = a + nseq * 8 * (int64_t)pix;
  float* res = b + nseq * 32 * (int64_t)pix;
  g_linear2_fold<<<nblk(nseq * S * 32), 256, 0, st>>>(p, h, enh, y, nseq, S, W, ph, pw);
  count_launch();
  if (!conv_stack(p.refine, y, a, b, res, nseq, H, W, st)) return false;
  g_interleave<<<nblk(nsamples * pix), 256, 0, st>>>(res, out, nsamples, pix);
  count_launch();
  return check_launch("generic head");
}

}  // namespace aft
