// "Next" rows N2 and N3, the callers / models either side of the hot path:
//   * pilot extraction -- what MatDataset._process_channel_data does per file (reference src/data/dataset.py:95-144):
//     the LS estimate arrives as a [subcarriers, symbols] grid that is zero off the pilot positions; the pilots are its
//     non-zero entries in row-major order.  Batched, on the device: one warp per sample, ballot + prefix compaction.
//     Index / byte work, HBM bound: reads 8 * scs * symbols bytes per sample.
//   * LinearEstimator -- y = W x + b with W [out, in] (reference src/models/linear.py:60-95).  HBM bound on the result
//     (4 * out bytes per sample against 4 * in read); W^T is staged in shared memory once per CTA.
#include "aft_internal.cuh"

namespace aft {

namespace {

// grid[b][i] != 0+0j exactly as the reference's boolean mask (NaN counts as non-zero, -0.0 as zero)
__device__ __forceinline__ bool nonzero(float2 v) { return v.x != 0.0f || v.y != 0.0f; }

__global__ void __launch_bounds__(256)
extract_pilots_kernel(const float2* __restrict__ grid, float2* __restrict__ pilots, int32_t* __restrict__ counts, int64_t batch,
                      int cells, int expected) {
  const int lane = threadIdx.x & 31;
  const unsigned lt = (1u << lane) - 1u;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t b = warp; b < batch; b += nwarps) {
    const float2* g = grid + b * cells;
    float2* out = pilots + b * expected;
    int found = 0;
    int base = 0;
    if ((cells & 1) == 0) {
      // two entries (16 bytes) per lane and four such loads in flight: 2 KB of the row per warp iteration
      const float4* g4 = reinterpret_cast<const float4*>(g);
      const int pairs = cells >> 1;
      for (; base + 4 * 32 <= pairs; base += 4 * 32) {
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = g4[base + u * 32 + lane];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float2 a = make_float2(v[u].x, v[u].y), c = make_float2(v[u].z, v[u].w);
          const bool na = nonzero(a), nc = nonzero(c);
          const unsigned ma = __ballot_sync(0xffffffffu, na), mc = __ballot_sync(0xffffffffu, nc);
          const int pa = found + __popc(ma & lt) + __popc(mc & lt);   // row-major: entries 2 lane, 2 lane + 1
          if (na && pa < expected) out[pa] = a;
          if (nc && pa + (na ? 1 : 0) < expected) out[pa + (na ? 1 : 0)] = c;
          found += __popc(ma) + __popc(mc);
        }
      }
      base *= 2;   // continue entry-wise
    }
    for (; base < cells; base += 32) {
      const int i = base + lane;
      float2 v = make_float2(0.f, 0.f);
      if (i < cells) v = g[i];
      const bool nz = i < cells && nonzero(v);
      const unsigned m = __ballot_sync(0xffffffffu, nz);
      const int pos = found + __popc(m & lt);
      if (nz && pos < expected) out[pos] = v;
      found += __popc(m);
    }
    if (lane == 0) counts[b] = found;
    // samples with fewer non-zeros than expected leave the remaining slots untouched; the host checks counts
  }
}

// LinearEstimator, weights in registers: thread <-> kOut output features (rows of W), samples streamed through shared
// memory in groups of kGroup (x rows are broadcast reads: kIn / 4 LDS.128 feed kOut * kIn FMAs), coalesced stores.
template <int kIn, int kOut>
__global__ void __launch_bounds__(864)
linear_regw_kernel(const float* __restrict__ w, const float* __restrict__ bias, const float* __restrict__ x, float* __restrict__ y,
                   int64_t batch, int out_dim) {
  constexpr int kGroup = 64;
  __shared__ __align__(16) float xs[2][kGroup * kIn];
  const int nthr = (out_dim + kOut - 1) / kOut;          // threads that own outputs: o = tid + j * nthr
  const bool owner = (int)threadIdx.x < nthr;
  float wr[kOut][kIn], br[kOut];
#pragma unroll
  for (int j = 0; j < kOut; ++j) {
    const int o = threadIdx.x + j * nthr;
    const bool ok = owner && o < out_dim;
    br[j] = ok ? bias[o] : 0.f;
#pragma unroll
    for (int k = 0; k < kIn; ++k) wr[j][k] = ok ? w[(size_t)o * kIn + k] : 0.f;
  }
  int buf = 0;
  for (int64_t b0 = (int64_t)blockIdx.x * kGroup; b0 < batch; b0 += (int64_t)gridDim.x * kGroup, buf ^= 1) {
    const int nb = (int)(batch - b0 < kGroup ? batch - b0 : kGroup);
    for (int i = threadIdx.x; i < nb * kIn; i += blockDim.x) xs[buf][i] = x[b0 * kIn + i];
    __syncthreads();   // double buffered: the next group's stores cannot overtake this group's reads
    if (owner) {
      for (int s = 0; s < nb; ++s) {
        float acc[kOut];
#pragma unroll
        for (int j = 0; j < kOut; ++j) acc[j] = br[j];
#pragma unroll
        for (int k4 = 0; k4 < kIn / 4; ++k4) {
          const float4 xv = *reinterpret_cast<const float4*>(&xs[buf][s * kIn + 4 * k4]);
#pragma unroll
          for (int j = 0; j < kOut; ++j) {
            acc[j] = fmaf(wr[j][4 * k4], xv.x, acc[j]); acc[j] = fmaf(wr[j][4 * k4 + 1], xv.y, acc[j]);
            acc[j] = fmaf(wr[j][4 * k4 + 2], xv.z, acc[j]); acc[j] = fmaf(wr[j][4 * k4 + 3], xv.w, acc[j]);
          }
        }
#pragma unroll
        for (int j = 0; j < kOut; ++j) {
          const int o = threadIdx.x + j * nthr;
          if (o < out_dim) y[(b0 + s) * out_dim + o] = acc[j];
        }
      }
    }
  }
}

// One CTA: W^T [in][out] in shared memory (when it fits), thread <-> output index, loop over the CTA's samples.
template <bool kSmemW>
__global__ void __launch_bounds__(256)
linear_kernel(const float* __restrict__ w, const float* __restrict__ bias, const float* __restrict__ x, float* __restrict__ y,
              int64_t batch, int in_dim, int out_dim) {
  extern __shared__ float sm[];
  float* wt = sm;                                   // [in][out] (kSmemW)
  float* xs = sm + (kSmemW ? (size_t)in_dim * out_dim : 0);   // [8][in]: a group of samples
  if (kSmemW) {
    for (int i = threadIdx.x; i < in_dim * out_dim; i += blockDim.x) {
      const int o = i / in_dim, k = i - o * in_dim;   // coalesced read of W [out][in]
      wt[k * out_dim + o] = w[i];
    }
  }
  __syncthreads();
  constexpr int kGroup = 8;
  for (int64_t b0 = (int64_t)blockIdx.x * kGroup; b0 < batch; b0 += (int64_t)gridDim.x * kGroup) {
    const int nb = (int)(batch - b0 < kGroup ? batch - b0 : kGroup);
    for (int i = threadIdx.x; i < nb * in_dim; i += blockDim.x) xs[i] = x[b0 * in_dim + i];
    __syncthreads();
    for (int o = threadIdx.x; o < out_dim; o += blockDim.x) {
      float acc[kGroup];
      const float bo = bias[o];
#pragma unroll
      for (int j = 0; j < kGroup; ++j) acc[j] = bo;
      for (int k = 0; k < in_dim; ++k) {
        const float wk = kSmemW ? wt[k * out_dim + o] : w[(size_t)o * in_dim + k];
#pragma unroll
        for (int j = 0; j < kGroup; ++j) acc[j] = fmaf(wk, xs[j * in_dim + k], acc[j]);   // xs: broadcast reads
      }
#pragma unroll
      for (int j = 0; j < kGroup; ++j)
        if (j < nb) y[(b0 + j) * out_dim + o] = acc[j];
    }
    __syncthreads();
  }
}

}  // namespace

bool launch_extract_pilots(const float2* grid, float2* pilots, int32_t* counts, int64_t batch, int cells, int expected, cudaStream_t st) {
  if (batch <= 0) return true;
  int64_t blocks = (batch + 7) / 8;           // 8 warps per CTA, one sample per warp per pass
  if (blocks > 148 * 8) blocks = 148 * 8;
  extract_pilots_kernel<<<(unsigned)blocks, 256, 0, st>>>(grid, pilots, counts, batch, cells, expected);
  count_launch();
  return check_launch("extract_pilots_kernel");
}

bool launch_linear(const float* w, const float* bias, const float* x, float* y, int64_t batch, int in_dim, int out_dim, cudaStream_t st) {
  if (batch <= 0) return true;
  if (in_dim == 24 && out_dim <= 2 * 864) {   // the reference configuration (24 pilots -> 1680 grid points)
    int64_t nb = (batch + 63) / 64;
    if (nb > 148) nb = 148;                   // 864 threads x 72 registers: one CTA per SM, grid-stride over sample groups
    linear_regw_kernel<24, 2><<<(unsigned)nb, 864, 0, st>>>(w, bias, x, y, batch, out_dim);
    count_launch();
    return check_launch("linear_regw_kernel");
  }
  const size_t wbytes = (size_t)in_dim * out_dim * sizeof(float), xbytes = (size_t)8 * in_dim * sizeof(float);
  int64_t blocks = (batch + 7) / 8;
  if (blocks > 148) blocks = 148;              // persistent: W^T is staged once per CTA
  if (wbytes + xbytes <= 200 * 1024) {
    if (cudaFuncSetAttribute(linear_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(wbytes + xbytes)) != cudaSuccess) {
      set_error("linear_kernel: cannot opt in to %zu bytes of shared memory", wbytes + xbytes);
      return false;
    }
    linear_kernel<true><<<(unsigned)blocks, 256, wbytes + xbytes, st>>>(w, bias, x, y, batch, in_dim, out_dim);
  } else {
    if (xbytes > 48 * 1024 &&
        cudaFuncSetAttribute(linear_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)xbytes) != cudaSuccess) {
      set_error("linear_kernel: cannot opt in to %zu bytes of shared memory", xbytes);
      return false;
    }
    if (blocks < 148 && out_dim > 4096) blocks = 148;   // few samples, many outputs: idle CTAs cost nothing
    linear_kernel<false><<<(unsigned)blocks, 256, xbytes, st>>>(w, bias, x, y, batch, in_dim, out_dim);
  }
  count_launch();
  return check_launch("linear_kernel");
}

}  // namespace aft
