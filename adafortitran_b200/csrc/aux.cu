// "Next" rows N2 and N3, the callers / models either side of the hot path:
//   * pilot extraction -- what MatDataset._process_channel_data does per file (reference src/data/dataset.py:95-144):
//     the LS estimate arrives as a [subcarriers, symbols] grid that is zero off the pilot positions; the pilots are its
//     non-zero entries in row-major order.  Batched, on the device: one warp per sample, ballot + prefix compaction.
//     Index / byte work, HBM bound: reads 8 * scs * symbols bytes per sample.
//   * LinearEstimator -- y = W x + b with W [out, in] (reference src/models/linear.py:60-95).  HBM bound on the result
//     (4 * out bytes per sample against 4 * in read); W^T is staged in shared memory once per CTA.
#include "aft_internal.cuh"
#include "tc_ptx.cuh"

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

// LinearEstimator, weights in registers: thread <-> FOUR CONSECUTIVE output features (rows of W), each row held as
// kIn / 2 packed fp32 pairs (w_k, w_k+1).  Samples are streamed through shared memory in groups of kGroup; a 16-byte
// broadcast load delivers (x_k .. x_k+3), i.e. the packed operands of two FFMA2 per output: the accumulator pair of an
// output holds its even-k and odd-k partial sums, added at the end.  The result leaves as ONE 16-byte store per thread and
// sample: every variant with 4-byte stores (scalar FMA, FFMA2, 2 / 4 / 5 outputs per thread, 11 .. 27 warps) ran at the
// same ~800 clk per sample and SM, i.e. one warp-wide store instruction per ~15 clk -- the store path is paid per
// instruction, not per byte (a 16-byte-per-thread fill reaches 7.5 TB/s on this GPU, tools/write_bw.py).
template <int kIn, int kOut, int kThreadsL>
__global__ void __launch_bounds__(kThreadsL)
linear_regw_kernel(const float* __restrict__ w, const float* __restrict__ bias, const float* __restrict__ x, float* __restrict__ y,
                   int64_t batch, int out_dim) {
  static_assert(kIn % 4 == 0, "four k steps per shared load");
  static_assert(kOut % 4 == 0, "16-byte stores");
  constexpr int kGroup = 64;
  __shared__ __align__(16) float xs[2][kGroup * kIn + 4];   // + 4: the pipelined load runs one step past the last sample
  const int nthr = out_dim / kOut;                          // out_dim % kOut == 0 (checked by the launcher)
  const bool owner = (int)threadIdx.x < nthr;
  unsigned long long wr[kOut][kIn / 2];
  float br[kOut];
#pragma unroll
  for (int j = 0; j < kOut; ++j) {
    const size_t o = (size_t)threadIdx.x * kOut + j;
    br[j] = owner ? bias[o] : 0.f;
#pragma unroll
    for (int k2 = 0; k2 < kIn / 2; ++k2) {
      const float a = owner ? w[o * kIn + 2 * k2] : 0.f, b = owner ? w[o * kIn + 2 * k2 + 1] : 0.f;
      asm("mov.b64 %0, {%1, %2};" : "=l"(wr[j][k2]) : "f"(a), "f"(b));
    }
  }
  // x groups arrive by 1-D bulk copies (one elected thread, completion on an mbarrier per buffer): the copy of group
  // g + 1 is issued before group g is computed, so no warp ever waits for a global load
  __shared__ __align__(8) unsigned long long bars[2];
  const uint32_t bar0 = ptx::smem_u32(&bars[0]);
  const int64_t stride = (int64_t)gridDim.x * kGroup;
  auto fetch = [&](int64_t g0, int b) {
    const uint32_t bytes = (uint32_t)((batch - g0 < kGroup ? batch - g0 : kGroup) * kIn * sizeof(float));
    ptx::mbar_arrive_expect_tx(bar0 + 8 * b, bytes);
    ptx::bulk_g2s(ptx::smem_u32(&xs[b][0]), x + g0 * kIn, bytes, bar0 + 8 * b);
  };
  if (threadIdx.x == 0) {
    ptx::mbar_init(bar0, 1);
    ptx::mbar_init(bar0 + 8, 1);
    ptx::fence_mbar_init();
    if ((int64_t)blockIdx.x * kGroup < batch) fetch((int64_t)blockIdx.x * kGroup, 0);
  }
  __syncthreads();
  int buf = 0;
  uint32_t it = 0;
  for (int64_t b0 = (int64_t)blockIdx.x * kGroup; b0 < batch; b0 += stride, buf ^= 1, ++it) {
    const int nb = (int)(batch - b0 < kGroup ? batch - b0 : kGroup);
    __syncthreads();   // every warp has finished the previous group: its buffer may be refilled
    if (threadIdx.x == 0 && b0 + stride < batch) {
      ptx::fence_proxy_async_smem();
      fetch(b0 + stride, buf ^ 1);
    }
    ptx::mbar_wait(bar0 + 8 * buf, (it >> 1) & 1);
    if (owner) {
      float4* yp = reinterpret_cast<float4*>(y + b0 * out_dim) + threadIdx.x * (kOut / 4);
      const ulonglong2* xp = reinterpret_cast<const ulonglong2*>(&xs[buf][0]);
      ulonglong2 xa = xp[0];                               // software pipeline: the next 16 bytes of x are in flight
      for (int s = 0; s < nb; ++s, yp += out_dim / 4, xp += kIn / 4) {
        unsigned long long acc[kOut];
#pragma unroll
        for (int j = 0; j < kOut; ++j) asm("mov.b64 %0, {%1, %2};" : "=l"(acc[j]) : "f"(br[j]), "f"(0.f));
#pragma unroll
        for (int k4 = 0; k4 < kIn / 4; ++k4) {
          const ulonglong2 xb = xp[k4 + 1];                // (x_k+4 .. x_k+7), or the next sample's first four (pad at the end)
#pragma unroll
          for (int j = 0; j < kOut; ++j) {
            asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc[j]) : "l"(wr[j][2 * k4]), "l"(xa.x));
            asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc[j]) : "l"(wr[j][2 * k4 + 1]), "l"(xa.y));
          }
          xa = xb;
        }
        float r[kOut];
#pragma unroll
        for (int j = 0; j < kOut; ++j) {
          float a0, a1;
          asm("mov.b64 {%0, %1}, %2;" : "=f"(a0), "=f"(a1) : "l"(acc[j]));
          r[j] = a0 + a1;
        }
#pragma unroll
        for (int j = 0; j < kOut; j += 4) yp[j / 4] = make_float4(r[j], r[j + 1], r[j + 2], r[j + 3]);
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
  if (in_dim == 24 && out_dim % 4 == 0 && out_dim <= 4 * 448 && ((reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(x)) & 15) == 0) {
    // the reference configuration (24 pilots -> 1680 grid points): 420 owner threads x 4 consecutive outputs
    int64_t nb = (batch + 63) / 64;
    if (nb > 148) nb = 148;                   // 14 warps x 128 registers: one CTA per SM, grid-stride over sample groups
    linear_regw_kernel<24, 4, 448><<<(unsigned)nb, 448, 0, st>>>(w, bias, x, y, batch, out_dim);
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
