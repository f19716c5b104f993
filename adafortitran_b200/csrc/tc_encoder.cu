// AFT_BF16 path: the 6-layer post-norm transformer encoder (reference src/models/blocks/encoders.py:44-55,69 ->
// torch _transformer_encoder_layer_fwd) as ONE persistent kernel on the 5th-gen tensor cores.
//
// One CTA per SM; a CTA owns one 280-token sequence at a time and carries it through all layers.  The
// residual stream never leaves the SM: it lives in shared memory as the bf16 A-operand image of the next GEMM.
// Per sequence the only HBM traffic is the 72 KB input image and the 140 KB fp32 result; weights (256 KB bf16
// per layer) are streamed from L2 by 1-D bulk copies (cp.async.bulk) into a small ring.
//
//   warp 0      : producer  -- bulk copies (sequence image, weight slices), mbarrier expect_tx
//   warp 1      : MMA issuer -- one thread issues every tcgen05.mma; owns the TMEM allocation
//   warps 2..9  : compute    -- TMEM -> registers epilogues: bias, softmax, 1/l, residual + LayerNorm, GELU;
//                               write the next operand image (bf16) to shared memory / P to TMEM
//
// All GEMMs use M = 128 row tiles (3 per sequence; the rows past 280 of the third tile read whatever follows in
// shared memory -- rows of A are independent, the corresponding accumulator lanes are never read).
//
// Shared memory map (bytes, every region 1024-aligned; operand layouts in tc_layout.cuh):
//   O    [      0,  73728)  attention output image (A of out_proj)   | FFN: hidden chunk images, 2 x 36864
//   X    [  73728, 147456)  residual stream image (A of QKV / FFN1, residual of both LayerNorms)
//   QKV  [ 147456, 202752)  Q_g | K_g | V_g of the current head, 288 x 64 B each (SWIZZLE_64B)
//                           | out_proj / FFN weight ring: 3 slots x 16384
//   W    [ 202752, 227328)  in_proj slice of one head: rows q_g | k_g | v_g (96 x K=128)
//   MISC [ 227328, 230400)  mbarriers, TMEM base, softmax max / sum exchange
// Tensor memory map (columns): S [0,288)  P [288,432) (bf16 pairs)  O_acc [432,464);
//   QKV accumulators alias S; out_proj / FFN2 accumulators [0,384); FFN1 accumulators [384,448) [448,512).
#include <cmath>
#include <cstdio>
#include <vector>

#include "tc_encoder.cuh"
#include "tc_layout.cuh"
#include "tc_ptx.cuh"

namespace aft {

namespace {

using namespace ptx;

constexpr int kTcThreads = 320;
constexpr int kTcMaxLayers = 8;
constexpr int kVecPerLayer = 1408;   // b_in 384 | b_out 128 | b_l1 256 | b_l2 128 | n1_w n1_b n2_w n2_b 4x128
constexpr int kVecBOut = 384, kVecBL1 = 512, kVecBL2 = 768, kVecN1W = 896, kVecN1B = 1024, kVecN2W = 1152, kVecN2B = 1280;

constexpr uint32_t OFF_O = 0, OFF_X = 73728, OFF_QKV = 147456, OFF_W = 202752, OFF_MISC = 227328;
constexpr uint32_t kQkvPart = 18432;            // 288 rows x 64 B
constexpr uint32_t kRingSlot = 16384;
constexpr uint32_t kHidBytes = 36864;           // 288 rows x 128 B
constexpr uint32_t kWInSlice = 24576;           // 2 chunks x 96 rows x 128 B
constexpr uint32_t kMiscBytes = 3072;
constexpr uint32_t kTcSmemBytes = OFF_MISC + kMiscBytes + 1024;   // + alignment slack

// MISC offsets
constexpr uint32_t MB_MMA_DONE = 0, MB_OPS_READY = 8, MB_X_FULL = 16, MB_X_FREE = 24, MB_ATTN_DONE = 32;
constexpr uint32_t MB_W_FULL = 40;    // 4 barriers: [0] = in_proj slot, [1..3] = ring slots
constexpr uint32_t MB_W_EMPTY = 72;   // 4 barriers
constexpr uint32_t MISC_TMEM_PTR = 128;
constexpr uint32_t MISC_XMAX = 256;   // [2][128] f32
constexpr uint32_t MISC_XSUM = 1280;  // [2][128] f32

constexpr uint32_t TM_S = 0, TM_P = 288, TM_O = 432, TM_QKV = 0, TM_OUT = 0, TM_F1 = 384;

constexpr uint32_t kIdescQkv = make_idesc_bf16(128, 96, false, false);
constexpr uint32_t kIdescS = make_idesc_bf16(128, 144, false, false);
constexpr uint32_t kIdescPV = make_idesc_bf16(128, 32, false, true);     // B = V, MN-major
constexpr uint32_t kIdescN128 = make_idesc_bf16(128, 128, false, false);
constexpr uint32_t kIdescN64 = make_idesc_bf16(128, 64, false, false);

__constant__ float c_vec[kTcMaxLayers * kVecPerLayer];

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void unpack_bf16x2(uint32_t w, float& lo, float& hi) {
  lo = __uint_as_float(w << 16);
  hi = __uint_as_float(w & 0xFFFF0000u);
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint4 ld_shared_v4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void st_shared_f32(uint32_t addr, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ float ld_shared_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
  return v;
}

// =============================================================================================
// MMA issue helpers (called by the single issuing thread).  `sb` = 1024-aligned shared base address.
// =============================================================================================
// D[tile] (N cols at d_col) = A(image with 128-B rows, `a_rows` rows per K-chunk)[tile rows] . B(image, b_rows per chunk)^T
__device__ __forceinline__ void issue_gemm_sw128(uint32_t tmem, uint32_t d_col, uint32_t a_base, uint32_t a_chunk_bytes,
                                                 uint32_t b_base, uint32_t b_chunk_bytes, int ksteps, uint32_t idesc,
                                                 bool accumulate_first) {
  for (int ks = 0; ks < ksteps; ++ks) {
    const uint32_t a = a_base + (ks >> 2) * a_chunk_bytes + (ks & 3) * 32;
    const uint32_t b = b_base + (ks >> 2) * b_chunk_bytes + (ks & 3) * 32;
    mma_ss(tmem + d_col, desc_k_sw128(a), desc_k_sw128(b), idesc, accumulate_first || ks > 0);
  }
}
// S[tile t] = Q_g[tile t] . K_g^T   (two N = 144 halves, K = 32)
__device__ __forceinline__ void issue_scores(uint32_t tmem, uint32_t sb, int t) {
  const uint32_t q = sb + OFF_QKV + t * 128 * 64;
  const uint32_t k = sb + OFF_QKV + kQkvPart;
#pragma unroll
  for (int nh = 0; nh < 2; ++nh)
#pragma unroll
    for (int ks = 0; ks < 2; ++ks)
      mma_ss(tmem + TM_S + nh * 144, desc_k_sw64(q + ks * 32), desc_k_sw64(k + nh * 144 * 64 + ks * 32), kIdescS, ks > 0);
}
// O_acc = P (TMEM, bf16 pairs) . V_g   (K = 288 keys = 18 steps, N = 32)
__device__ __forceinline__ void issue_pv(uint32_t tmem, uint32_t sb) {
  const uint32_t v = sb + OFF_QKV + 2 * kQkvPart;
#pragma unroll 1
  for (int ks = 0; ks < 18; ++ks) mma_ts(tmem + TM_O, tmem + TM_P + ks * 8, desc_mn_sw64(v + ks * 1024), kIdescPV, ks > 0);
}

// =============================================================================================
// compute-warp epilogues.  q = warp % 4 (TMEM lane quadrant), half = warpgroup (0: warps 2-5, 1: warps 6-9)
// =============================================================================================
// QKV accumulators of (head g, row tile t) -> + bias -> bf16 -> Q_g / K_g / V_g images (SWIZZLE_64B rows of 64 B)
__device__ __forceinline__ void epi_qkv(uint32_t tmem, uint32_t sb, int l, int g, int t, int q, int lane) {
  const int r = t * 128 + q * 32 + lane;
  const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + TM_QKV + t * 96;
#pragma unroll
  for (int part = 0; part < 3; ++part) {
    uint32_t a[16], b[16];
    tmem_ld16(taddr + part * 32, a);
    tmem_ld16(taddr + part * 32 + 16, b);
    tmem_wait_ld();
    const float* bias = c_vec + l * kVecPerLayer + part * 128 + g * 32;
    uint32_t pk[16];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      pk[j] = pack_bf16x2(__uint_as_float(a[2 * j]) + bias[2 * j], __uint_as_float(a[2 * j + 1]) + bias[2 * j + 1]);
      pk[8 + j] = pack_bf16x2(__uint_as_float(b[2 * j]) + bias[16 + 2 * j], __uint_as_float(b[2 * j + 1]) + bias[16 + 2 * j + 1]);
    }
    const uint32_t row = sb + OFF_QKV + part * kQkvPart + r * 64;
    const int sw = (r >> 1) & 3;
#pragma unroll
    for (int u = 0; u < 4; ++u) st_shared_v4(row + ((u ^ sw) << 4), pk[4 * u], pk[4 * u + 1], pk[4 * u + 2], pk[4 * u + 3]);
  }
}

// softmax over the 280 keys of one query row; this thread covers 144 score columns (half 1: 136 valid).
// Scores are already in log2 units (q rows of in_proj pre-scaled by log2(e)/sqrt(dh)).
__device__ __forceinline__ void epi_softmax(uint32_t tmem, uint32_t sb, bool active, int q, int half, int lane) {
  const int rt = q * 32 + lane;
  const uint32_t lane_addr = tmem + ((uint32_t)(q * 32) << 16);
  float v[144];
  float m = -INFINITY;
  if (active) {
#pragma unroll
    for (int i = 0; i < 9; ++i) {
      uint32_t x[16];
      tmem_ld16(lane_addr + TM_S + half * 144 + i * 16, x);
      tmem_wait_ld();
#pragma unroll
      for (int j = 0; j < 16; ++j) v[i * 16 + j] = __uint_as_float(x[j]);
    }
    if (half == 1) {
#pragma unroll
      for (int j = kS - 144; j < 144; ++j) v[j] = -INFINITY;   // keys 280..287 are padding
    }
#pragma unroll
    for (int j = 0; j < 144; ++j) m = fmaxf(m, v[j]);
    st_shared_f32(sb + OFF_MISC + MISC_XMAX + (half * 128 + rt) * 4, m);
  }
  named_bar_sync(1 + q, 64);
  if (active) {
    m = fmaxf(m, ld_shared_f32(sb + OFF_MISC + MISC_XMAX + ((half ^ 1) * 128 + rt) * 4));
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < 9; ++i) {
      uint32_t pk[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float p0 = ex2(v[i * 16 + 2 * j] - m), p1 = ex2(v[i * 16 + 2 * j + 1] - m);
        sum += p0 + p1;
        pk[j] = pack_bf16x2(p0, p1);
      }
      tmem_st8(lane_addr + TM_P + half * 72 + i * 8, pk);
    }
    tmem_wait_st();
    st_shared_f32(sb + OFF_MISC + MISC_XSUM + (half * 128 + rt) * 4, sum);
  }
}

// O_acc (128 x 32) of (head g, tile t) -> / l -> bf16 -> O image columns g*32 .. g*32+31 (this thread: 16 of them)
__device__ __forceinline__ void epi_o(uint32_t tmem, uint32_t sb, int g, int t, int q, int half, int lane) {
  const int rt = q * 32 + lane, r = t * 128 + rt;
  const float l = ld_shared_f32(sb + OFF_MISC + MISC_XSUM + rt * 4) + ld_shared_f32(sb + OFF_MISC + MISC_XSUM + (128 + rt) * 4);
  const float inv = 1.0f / l;
  uint32_t a[16];
  tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + TM_O + half * 16, a);
  tmem_wait_ld();
  uint32_t pk[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) pk[j] = pack_bf16x2(__uint_as_float(a[2 * j]) * inv, __uint_as_float(a[2 * j + 1]) * inv);
  const uint32_t row = sb + OFF_O + (g >> 1) * kXChunkBytes + r * 128;
  const int u0 = (g & 1) * 4 + half * 2;
  st_shared_v4(row + (((u0) ^ (r & 7)) << 4), pk[0], pk[1], pk[2], pk[3]);
  st_shared_v4(row + (((u0 + 1) ^ (r & 7)) << 4), pk[4], pk[5], pk[6], pk[7]);
}

// out_proj / linear2 accumulators (tile t, 128 columns) + bias + residual (X image) -> LayerNorm -> X image in place
// (+ fp32 rows to h_out after the last layer).  One thread owns one full row: statistics need no exchange.
__device__ __forceinline__ void epi_ln(uint32_t tmem, uint32_t sb, int l, int which, int t, int q, int lane, float* h_out_seq) {
  const int r = t * 128 + q * 32 + lane;
  const float* vec = c_vec + l * kVecPerLayer;
  const float* bias = vec + (which == 1 ? kVecBOut : kVecBL2);
  const float* gam = vec + (which == 1 ? kVecN1W : kVecN2W);
  const float* bet = vec + (which == 1 ? kVecN1B : kVecN2B);
  const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + TM_OUT + t * 128;
  const uint32_t xrow = sb + OFF_X + r * 128;
  float v[128];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    uint32_t a[16];
    tmem_ld16(taddr + i * 16, a);
    tmem_wait_ld();
#pragma unroll
    for (int hu = 0; hu < 2; ++hu) {
      const int u = i * 2 + hu;   // 16-byte unit = 8 columns
      const uint4 xr = ld_shared_v4(xrow + (u >> 3) * kXChunkBytes + (((u & 7) ^ (r & 7)) << 4));
      float x[8];
      unpack_bf16x2(xr.x, x[0], x[1]); unpack_bf16x2(xr.y, x[2], x[3]);
      unpack_bf16x2(xr.z, x[4], x[5]); unpack_bf16x2(xr.w, x[6], x[7]);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[u * 8 + j] = __uint_as_float(a[hu * 8 + j]) + bias[u * 8 + j] + x[j];
    }
  }
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < 128; ++j) s += v[j];
  const float mean = s * (1.0f / 128.0f);
  float qv = 0.f;
#pragma unroll
  for (int j = 0; j < 128; ++j) { const float d = v[j] - mean; qv = fmaf(d, d, qv); }
  const float rstd = rsqrtf(qv * (1.0f / 128.0f) + 1e-5f);
#pragma unroll
  for (int u = 0; u < 16; ++u) {
    float o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = (v[u * 8 + j] - mean) * rstd * gam[u * 8 + j] + bet[u * 8 + j];
    st_shared_v4(xrow + (u >> 3) * kXChunkBytes + (((u & 7) ^ (r & 7)) << 4), pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]),
                 pack_bf16x2(o[4], o[5]), pack_bf16x2(o[6], o[7]));
    if (h_out_seq != nullptr && r < kS) {
      float4* dst = reinterpret_cast<float4*>(h_out_seq + r * kD + u * 8);
      dst[0] = make_float4(o[0], o[1], o[2], o[3]);
      dst[1] = make_float4(o[4], o[5], o[6], o[7]);
    }
  }
}

// FFN1 accumulators (chunk c = 64 hidden units, tile t; this thread: 32 of them) + bias -> GELU / ReLU -> hidden image
__device__ __forceinline__ void epi_act(uint32_t tmem, uint32_t sb, int l, int c, int t, int buf, int act, int q, int half, int lane) {
  const int r = t * 128 + q * 32 + lane;
  const float* bias = c_vec + l * kVecPerLayer + kVecBL1 + c * 64 + half * 32;
  const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + TM_F1 + buf * 64 + half * 32;
  uint32_t a[16], b[16];
  tmem_ld16(taddr, a);
  tmem_ld16(taddr + 16, b);
  tmem_wait_ld();
  float f[32];
#pragma unroll
  for (int j = 0; j < 16; ++j) { f[j] = __uint_as_float(a[j]) + bias[j]; f[16 + j] = __uint_as_float(b[j]) + bias[16 + j]; }
  if (act == AFT_ACT_GELU) {
#pragma unroll
    for (int j = 0; j < 32; ++j) f[j] = gelu_erf(f[j]);
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
  }
  const uint32_t row = sb + OFF_O + (c & 1) * kHidBytes + r * 128;
#pragma unroll
  for (int u = 0; u < 4; ++u)
    st_shared_v4(row + (((half * 4 + u) ^ (r & 7)) << 4), pack_bf16x2(f[8 * u], f[8 * u + 1]), pack_bf16x2(f[8 * u + 2], f[8 * u + 3]),
                 pack_bf16x2(f[8 * u + 4], f[8 * u + 5]), pack_bf16x2(f[8 * u + 6], f[8 * u + 7]));
}

// =============================================================================================
// the persistent encoder kernel
// =============================================================================================
struct EncParams {
  const char* x_images;       // [nseq][kXImageBytes]
  float* h_out;               // [nseq][280][128]
  const TcLayer* layers;      // device table
  int num_layers;
  int activation;
  int64_t nseq;
};

__global__ void __launch_bounds__(kTcThreads, 1) encoder_kernel(EncParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sb = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t misc = sb + OFF_MISC;

  if (threadIdx.x == 0) {
    mbar_init(misc + MB_MMA_DONE, 1);
    mbar_init(misc + MB_OPS_READY, 256);
    mbar_init(misc + MB_X_FULL, 1);
    mbar_init(misc + MB_X_FREE, 256);
    mbar_init(misc + MB_ATTN_DONE, 1);
    for (int i = 0; i < 4; ++i) { mbar_init(misc + MB_W_FULL + 8 * i, 1); mbar_init(misc + MB_W_EMPTY + 8 * i, 1); }
    fence_mbar_init();
  }
  if (warp == 1) { tmem_alloc(misc + MISC_TMEM_PTR, 512); tmem_relinquish(); }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  uint32_t tmem;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(misc + MISC_TMEM_PTR));

  const int L = p.num_layers;

  if (warp == 0) {
    // ----------------------------------------------------------------------------- producer
    if (lane == 0) {
      uint32_t n_in = 0, n_ring = 0, n_attn = 0, n_seq = 0;   // use counters -> slot / parity
      for (int64_t seq = blockIdx.x; seq < p.nseq; seq += gridDim.x, ++n_seq) {
        if (n_seq > 0) mbar_wait(misc + MB_X_FREE, (n_seq - 1) & 1);
        mbar_arrive_expect_tx(misc + MB_X_FULL, kXImageBytes);
        bulk_g2s(sb + OFF_X, p.x_images + seq * (int64_t)kXImageBytes, kXImageBytes, misc + MB_X_FULL);
        for (int l = 0; l < L; ++l) {
          const TcLayer& W = p.layers[l];
          for (int g = 0; g < 4; ++g, ++n_in) {
            if (n_in > 0) mbar_wait(misc + MB_W_EMPTY, (n_in - 1) & 1);
            mbar_arrive_expect_tx(misc + MB_W_FULL, kWInSlice);
            bulk_g2s(sb + OFF_W, reinterpret_cast<const char*>(W.w_in) + g * kWInSlice, kWInSlice, misc + MB_W_FULL);
          }
          mbar_wait(misc + MB_ATTN_DONE, n_attn & 1);   // Q/K/V images dead: the ring may overwrite them
          ++n_attn;
          for (int i = 0; i < 10; ++i, ++n_ring) {
            const int slot = n_ring % 3;
            const uint32_t use = n_ring / 3;              // how many times this slot was filled before
            if (use > 0) mbar_wait(misc + MB_W_EMPTY + 8 * (1 + slot), (use - 1) & 1);
            const char* src = i < 2 ? reinterpret_cast<const char*>(W.w_out) + i * kRingSlot
                                    : ((i & 1) == 0 ? reinterpret_cast<const char*>(W.w_l1) + ((i - 2) >> 1) * kRingSlot
                                                    : reinterpret_cast<const char*>(W.w_l2) + ((i - 3) >> 1) * kRingSlot);
            mbar_arrive_expect_tx(misc + MB_W_FULL + 8 * (1 + slot), kRingSlot);
            bulk_g2s(sb + OFF_QKV + slot * kRingSlot, src, kRingSlot, misc + MB_W_FULL + 8 * (1 + slot));
          }
        }
      }
    }
  } else if (warp == 1) {
    // ----------------------------------------------------------------------------- MMA issuer
    if (lane == 0) {
      uint32_t k = 0;        // step counter (shared schedule with the compute warps)
      uint32_t n_in = 0, n_ring = 0, n_seq = 0;
      auto wait_ops = [&]() {
        if (k > 0) mbar_wait(misc + MB_OPS_READY, (k - 1) & 1);
        tc_fence_after_sync();
      };
      auto end_step = [&]() { mma_commit(misc + MB_MMA_DONE); ++k; };
      auto ring_wait = [&]() -> uint32_t {   // wait for the next ring slot to be full; returns its smem address
        const int slot = n_ring % 3;
        mbar_wait(misc + MB_W_FULL + 8 * (1 + slot), (n_ring / 3) & 1);
        tc_fence_after_sync();
        return sb + OFF_QKV + slot * kRingSlot;
      };
      auto ring_release = [&]() { mma_commit(misc + MB_W_EMPTY + 8 * (1 + n_ring % 3)); ++n_ring; };

      for (int64_t seq = blockIdx.x; seq < p.nseq; seq += gridDim.x, ++n_seq) {
        mbar_wait(misc + MB_X_FULL, n_seq & 1);
        for (int l = 0; l < L; ++l) {
          for (int g = 0; g < 4; ++g, ++n_in) {
            // ---- step: QKV projection of head g, all three row tiles
            wait_ops();
            mbar_wait(misc + MB_W_FULL, n_in & 1);
            tc_fence_after_sync();
            for (int t = 0; t < 3; ++t)
              issue_gemm_sw128(tmem, TM_QKV + t * 96, sb + OFF_X + t * 128 * 128, kXChunkBytes, sb + OFF_W, 96 * 128, 8, kIdescQkv, false);
            mma_commit(misc + MB_W_EMPTY);
            end_step();
            // ---- steps: S0 | S1 + PV0 | S2 + PV1 | PV2
            for (int t = 0; t <= 3; ++t) {
              wait_ops();
              if (t > 0) issue_pv(tmem, sb);
              if (t < 3) issue_scores(tmem, sb, t);
              if (t == 3 && g == 3) mma_commit(misc + MB_ATTN_DONE);
              end_step();
            }
          }
          // ---- step: out_proj (A = O image, B = W_out in two ring slots, one per K-chunk)
          wait_ops();
          {
            const uint32_t w0 = ring_wait();
            for (int t = 0; t < 3; ++t)
              issue_gemm_sw128(tmem, TM_OUT + t * 128, sb + OFF_O + t * 128 * 128, kXChunkBytes, w0, 0, 4, kIdescN128, false);
            ring_release();
            const uint32_t w1 = ring_wait();
            for (int t = 0; t < 3; ++t)
              issue_gemm_sw128(tmem, TM_OUT + t * 128, sb + OFF_O + kXChunkBytes + t * 128 * 128, kXChunkBytes, w1, 0, 4, kIdescN128, true);
            ring_release();
          }
          end_step();
          // ---- FFN: 4 chunks of 64 hidden units; FFN2 partial products accumulate in TMEM over the chunks
          uint32_t w1c = 0;
          for (int c = 0; c < 4; ++c) {
            for (int t = 0; t < 3; ++t) {
              wait_ops();
              if (t == 0) w1c = ring_wait();                     // linear1 rows 64c .. 64c+63 (K = 128)
              issue_gemm_sw128(tmem, TM_F1 + ((3 * c + t) & 1) * 64, sb + OFF_X + t * 128 * 128, kXChunkBytes, w1c, 64 * 128, 8,
                               kIdescN64, false);
              if (t == 2) ring_release();
              end_step();
            }
            // ---- step: FFN2 partial, hidden chunk c (K = 64) x linear2 columns 64c .. 64c+63
            wait_ops();
            {
              const uint32_t w2c = ring_wait();
              for (int t = 0; t < 3; ++t)
                issue_gemm_sw128(tmem, TM_OUT + t * 128, sb + OFF_O + (c & 1) * kHidBytes + t * 128 * 128, 0, w2c, 0, 4, kIdescN128, c > 0);
              ring_release();
            }
            end_step();
          }
        }
      }
    }
  } else {
    // ----------------------------------------------------------------------------- compute warps
    const int q = warp & 3, half = (warp - 2) >> 2;
    const bool tile2_active = (q == 0);   // third row tile: only rows 256..287 exist
    uint32_t k = 0;
    auto wait_mma = [&]() {
      mbar_wait(misc + MB_MMA_DONE, k & 1);
      tc_fence_after_sync();
    };
    auto end_step = [&]() {
      tc_fence_before_sync();
      fence_proxy_async_smem();
      mbar_arrive(misc + MB_OPS_READY);
      ++k;
    };
    for (int64_t seq = blockIdx.x; seq < p.nseq; seq += gridDim.x) {
      float* h_seq = p.h_out + seq * (int64_t)kS * kD;
      for (int l = 0; l < L; ++l) {
        for (int g = 0; g < 4; ++g) {
          wait_mma();
          if (half == 0) {
            epi_qkv(tmem, sb, l, g, 0, q, lane);
            if (tile2_active) epi_qkv(tmem, sb, l, g, 2, q, lane);
          } else {
            epi_qkv(tmem, sb, l, g, 1, q, lane);
          }
          end_step();
          for (int t = 0; t <= 3; ++t) {
            wait_mma();
            if (t > 0 && (t - 1 < 2 || tile2_active)) epi_o(tmem, sb, g, t - 1, q, half, lane);
            if (t < 3) epi_softmax(tmem, sb, t < 2 || tile2_active, q, half, lane);
            end_step();
          }
        }
        wait_mma();
        if (half == 0) {
          epi_ln(tmem, sb, l, 1, 0, q, lane, nullptr);
          if (tile2_active) epi_ln(tmem, sb, l, 1, 2, q, lane, nullptr);
        } else {
          epi_ln(tmem, sb, l, 1, 1, q, lane, nullptr);
        }
        end_step();
        for (int c = 0; c < 4; ++c) {
          for (int t = 0; t < 3; ++t) {
            wait_mma();
            if (t < 2 || tile2_active) epi_act(tmem, sb, l, c, t, (3 * c + t) & 1, p.activation, q, half, lane);
            end_step();
          }
          wait_mma();
          if (c == 3) {
            float* ho = (l == L - 1) ? h_seq : nullptr;
            if (half == 0) {
              epi_ln(tmem, sb, l, 2, 0, q, lane, ho);
              if (tile2_active) epi_ln(tmem, sb, l, 2, 2, q, lane, ho);
            } else {
              epi_ln(tmem, sb, l, 2, 1, q, lane, ho);
            }
            if (l == L - 1) {   // the X image may now be replaced (async proxy) by the next sequence
              fence_proxy_async_smem();
              mbar_arrive(misc + MB_X_FREE);
            }
          }
          end_step();
        }
      }
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

// =============================================================================================
// weight packing: fp32 [N, K] row-major -> bf16 operand image(s)
// =============================================================================================
// One thread per 16-byte unit.  The destination is a sequence of `nblocks` images, block b holding rows
// [row0 + b*row_stride, +rows) and columns [col0 + b*col_stride, +64*chunks) of the source; rows whose index
// (within the block) is < scale_rows are multiplied by `scale` (in_proj q rows).
__global__ void pack_image_kernel(const float* __restrict__ src, int ld, __nv_bfloat16* __restrict__ dst, int nblocks, int rows,
                                  int chunks, int row0, int row_stride, int col0, int col_stride, int scale_rows, float scale) {
  const int units_per_block = rows * chunks * 8;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nblocks * units_per_block) return;
  const int b = i / units_per_block, rem = i - b * units_per_block;
  const int chunk = rem / (rows * 8), rem2 = rem - chunk * rows * 8;
  const int r = rem2 >> 3, u = rem2 & 7;
  const float* s = src + (int64_t)(row0 + b * row_stride + r) * ld + col0 + b * col_stride + chunk * 64 + u * 8;
  const float sc = r < scale_rows ? scale : 1.0f;
  uint4 pk;
  pk.x = pack_bf16x2(s[0] * sc, s[1] * sc); pk.y = pack_bf16x2(s[2] * sc, s[3] * sc);
  pk.z = pack_bf16x2(s[4] * sc, s[5] * sc); pk.w = pack_bf16x2(s[6] * sc, s[7] * sc);
  char* d = reinterpret_cast<char*>(dst) + (int64_t)b * rows * chunks * 128 + image_offset(r, chunk * 64 + u * 8, rows);
  *reinterpret_cast<uint4*>(d) = pk;
}

// in_proj head slice g: rows [q_g | k_g | v_g] gathered from rows g*32, 128+g*32, 256+g*32
__global__ void pack_inproj_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ dst, float qscale) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;   // one 16-byte unit: 4 slices x 2 chunks x 96 rows x 8 units
  if (i >= 4 * 2 * 96 * 8) return;
  const int g = i / 1536, rem = i - g * 1536;
  const int chunk = rem / 768, rem2 = rem - chunk * 768;
  const int r = rem2 >> 3, u = rem2 & 7;
  const int part = r >> 5, rr = r & 31;
  const float* s = w + (int64_t)(part * 128 + g * 32 + rr) * kD + chunk * 64 + u * 8;
  const float sc = part == 0 ? qscale : 1.0f;
  uint4 pk;
  pk.x = pack_bf16x2(s[0] * sc, s[1] * sc); pk.y = pack_bf16x2(s[2] * sc, s[3] * sc);
  pk.z = pack_bf16x2(s[4] * sc, s[5] * sc); pk.w = pack_bf16x2(s[6] * sc, s[7] * sc);
  char* d = reinterpret_cast<char*>(dst) + (int64_t)g * kWInSlice + image_offset(r, chunk * 64 + u * 8, 96);
  *reinterpret_cast<uint4*>(d) = pk;
}

__global__ void pack_vec_kernel(LayerPackF32 L, float* __restrict__ dst, float qscale) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= kVecPerLayer) return;
  float v;
  if (i < 384) v = L.in_b[i] * (i < 128 ? qscale : 1.0f);
  else if (i < kVecBL1) v = L.out_b[i - kVecBOut];
  else if (i < kVecBL2) v = L.l1_b[i - kVecBL1];
  else if (i < kVecN1W) v = L.l2_b[i - kVecBL2];
  else if (i < kVecN1B) v = L.n1_w[i - kVecN1W];
  else if (i < kVecN2W) v = L.n1_b[i - kVecN1B];
  else if (i < kVecN2B) v = L.n2_w[i - kVecN2W];
  else v = L.n2_b[i - kVecN2B];
  dst[i] = v;
}

constexpr size_t kLayerImageBytes = 4 * kWInSlice + 32768 + 65536 + 65536;   // 262,144

}  // namespace

// =============================================================================================
// host side
// =============================================================================================
bool tc_weights_alloc(TcWeights& w, int num_layers) {
  w.num_layers = num_layers;
  w.layers.assign(num_layers, TcLayer{});
  // arena: per layer the operand images + the epilogue vectors; then the device copy of the table
  const size_t per_layer = kLayerImageBytes + kVecPerLayer * sizeof(float);
  w.arena_bytes = num_layers * per_layer + num_layers * sizeof(TcLayer) + 1024;
  if (cudaMalloc(&w.arena, w.arena_bytes) != cudaSuccess) {
    set_error("tc_weights_alloc: cudaMalloc(%zu) failed: %s", w.arena_bytes, cudaGetErrorString(cudaGetLastError()));
    w.arena = nullptr;
    return false;
  }
  char* base = static_cast<char*>(w.arena);
  for (int l = 0; l < num_layers; ++l) {
    char* p = base + l * kLayerImageBytes;
    TcLayer& T = w.layers[l];
    T.w_in = reinterpret_cast<const __nv_bfloat16*>(p);
    T.w_out = reinterpret_cast<const __nv_bfloat16*>(p + 4 * kWInSlice);
    T.w_l1 = reinterpret_cast<const __nv_bfloat16*>(p + 4 * kWInSlice + 32768);
    T.w_l2 = reinterpret_cast<const __nv_bfloat16*>(p + 4 * kWInSlice + 32768 + 65536);
    const float* v = reinterpret_cast<const float*>(base + num_layers * kLayerImageBytes) + l * kVecPerLayer;
    T.b_in = v; T.b_out = v + kVecBOut; T.b_l1 = v + kVecBL1; T.b_l2 = v + kVecBL2;
    T.n1_w = v + kVecN1W; T.n1_b = v + kVecN1B; T.n2_w = v + kVecN2W; T.n2_b = v + kVecN2B;
  }
  w.layers_dev = reinterpret_cast<TcLayer*>(base + num_layers * per_layer);
  return true;
}

void tc_weights_free(TcWeights& w) {
  if (w.arena) cudaFree(w.arena);
  w.arena = nullptr;
}

bool tc_weights_pack(TcWeights& w, const std::vector<LayerPackF32>& src, cudaStream_t st) {
  const float qscale = 1.4426950408889634f / sqrtf((float)kDh);   // log2(e) / sqrt(dh): softmax runs on exp2
  for (int l = 0; l < w.num_layers; ++l) {
    const LayerPackF32& S = src[l];
    const TcLayer& T = w.layers[l];
    auto bf = [](const __nv_bfloat16* p) { return const_cast<__nv_bfloat16*>(p); };
    pack_inproj_kernel<<<(4 * 2 * 96 * 8 + 255) / 256, 256, 0, st>>>(S.in_w, bf(T.w_in), qscale);
    // out_proj: one image, 128 rows, K = 128 (2 chunks)
    pack_image_kernel<<<(128 * 2 * 8 + 255) / 256, 256, 0, st>>>(S.out_w, kD, bf(T.w_out), 1, 128, 2, 0, 0, 0, 0, 0, 1.f);
    // linear1: 4 images of 64 rows, K = 128
    pack_image_kernel<<<(4 * 64 * 2 * 8 + 255) / 256, 256, 0, st>>>(S.l1_w, kD, bf(T.w_l1), 4, 64, 2, 0, 64, 0, 0, 0, 1.f);
    // linear2: 4 images of 128 rows, one K-chunk each (columns 64c .. 64c+63 of the [128, 256] matrix)
    pack_image_kernel<<<(4 * 128 * 1 * 8 + 255) / 256, 256, 0, st>>>(S.l2_w, kFF, bf(T.w_l2), 4, 128, 1, 0, 0, 0, 64, 0, 1.f);
    pack_vec_kernel<<<(kVecPerLayer + 255) / 256, 256, 0, st>>>(S, const_cast<float*>(T.b_in), qscale);
    count_launch(5);
  }
  if (cudaMemcpyAsync(w.layers_dev, w.layers.data(), w.num_layers * sizeof(TcLayer), cudaMemcpyHostToDevice, st) != cudaSuccess) {
    set_error("tc_weights_pack: table upload failed: %s", cudaGetErrorString(cudaGetLastError()));
    return false;
  }
  return check_launch("tc_weights_pack");
}

namespace {
size_t align_up_sz(size_t v, size_t a) { return (v + a - 1) / a * a; }
}

size_t tc_workspace_bytes(int64_t bc) {
  const size_t nseq = 2 * (size_t)bc;
  return align_up_sz(nseq * kPix * sizeof(float), 1024) + align_up_sz(nseq * (size_t)kXImageBytes, 1024) +
         align_up_sz(nseq * (size_t)kS * kD * sizeof(float), 1024);
}

bool tc_forward_chunk(const TcWeights& w, const FrontPack& front, const HeadPack& head, int activation, int sm_count,
                      const float2* pilots, const float* snr, const float* ds, const float* dop, float2* out,
                      int64_t nsamples, void* workspace, cudaStream_t st) {
  if (w.num_layers > kTcMaxLayers) {
    set_error("AFT_BF16 path supports at most %d encoder layers (got %d)", kTcMaxLayers, w.num_layers);
    return false;
  }
  const int64_t nseq = 2 * nsamples;
  char* ws = static_cast<char*>(workspace);
  float* enh = reinterpret_cast<float*>(ws);
  char* ximg = ws + align_up_sz(nseq * kPix * sizeof(float), 1024);
  float* hout = reinterpret_cast<float*>(ximg + align_up_sz(nseq * (size_t)kXImageBytes, 1024));
  if (!launch_frontend(front, pilots, snr, ds, dop, enh, nullptr, reinterpret_cast<__nv_bfloat16*>(ximg), nsamples, st)) return false;
  // per-layer epilogue vectors -> constant bank (device-to-device, stream ordered; 5.5 KB per layer)
  if (cudaMemcpyToSymbolAsync(c_vec, w.layers[0].b_in, (size_t)w.num_layers * kVecPerLayer * sizeof(float), 0,
                              cudaMemcpyDeviceToDevice, st) != cudaSuccess) {
    set_error("tc_forward_chunk: constant upload failed: %s", cudaGetErrorString(cudaGetLastError()));
    return false;
  }
  if (cudaFuncSetAttribute(encoder_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcSmemBytes) != cudaSuccess) {
    set_error("encoder_kernel: cannot opt in to %u bytes of shared memory: %s", kTcSmemBytes, cudaGetErrorString(cudaGetLastError()));
    return false;
  }
  EncParams ep;
  ep.x_images = ximg;
  ep.h_out = hout;
  ep.layers = w.layers_dev;
  ep.num_layers = w.num_layers;
  ep.activation = activation;
  ep.nseq = nseq;
  const unsigned grid = (unsigned)(nseq < sm_count ? nseq : sm_count);
  encoder_kernel<<<grid, kTcThreads, kTcSmemBytes, st>>>(ep);
  count_launch();
  if (!check_launch("encoder_kernel")) return false;
  return launch_head(head, hout, enh, out, nsamples, st);
}

// =============================================================================================
// self tests of the tcgen05 building blocks (aft_selftest)
// =============================================================================================
namespace {

// which = 0: D[128,96] = A[128 rows of a 288-row X image, K=128] . B[96 rows, K=128]^T   (SW128 K-major both)
__global__ void __launch_bounds__(128, 1) selftest_gemm_kernel(const char* a_img, const char* b_img, float* d_out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sb = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar = sb + 110592, tptr = bar + 16, bar2 = bar + 8;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(bar2, 1); fence_mbar_init(); }
  if (warp == 0) { tmem_alloc(tptr, 128); tmem_relinquish(); }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  uint32_t tmem;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(tptr));
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(bar, kXImageBytes + kWInSlice);
    bulk_g2s(sb, a_img, kXImageBytes, bar);
    bulk_g2s(sb + kXImageBytes, b_img, kWInSlice, bar);
    mbar_wait(bar, 0);
    tc_fence_after_sync();
    // second row tile (rows 128..255) to exercise the tile offset
    issue_gemm_sw128(tmem, 0, sb + 128 * 128, kXChunkBytes, sb + kXImageBytes, 96 * 128, 8, kIdescQkv, false);
    mma_commit(bar2);
  }
  mbar_wait(bar2, 0);
  tc_fence_after_sync();
  const int row = warp * 32 + lane;
  for (int i = 0; i < 6; ++i) {
    uint32_t v[16];
    tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + i * 16, v);
    tmem_wait_ld();
    for (int j = 0; j < 16; ++j) d_out[row * 96 + i * 16 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 128);
}

// which = 1: one attention row tile with the production helpers: S = Q K^T (SW64), softmax -> P (TMEM), O = P V (MN-major)
__global__ void __launch_bounds__(kTcThreads, 1) selftest_attn_kernel(const char* qkv_img, float* s_out, float* o_out, int tile) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sb = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t misc = sb + OFF_MISC;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(misc + MB_MMA_DONE, 1); mbar_init(misc + MB_OPS_READY, 256); mbar_init(misc + MB_X_FULL, 1); fence_mbar_init(); }
  if (warp == 1) { tmem_alloc(misc + MISC_TMEM_PTR, 512); tmem_relinquish(); }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  uint32_t tmem;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(misc + MISC_TMEM_PTR));
  if (warp == 1 && lane == 0) {
    mbar_arrive_expect_tx(misc + MB_X_FULL, 3 * kQkvPart);
    bulk_g2s(sb + OFF_QKV, qkv_img, 3 * kQkvPart, misc + MB_X_FULL);
    mbar_wait(misc + MB_X_FULL, 0);
    tc_fence_after_sync();
    issue_scores(tmem, sb, tile);
    mma_commit(misc + MB_MMA_DONE);
    mbar_wait(misc + MB_OPS_READY, 0);
    tc_fence_after_sync();
    issue_pv(tmem, sb);
    mma_commit(misc + MB_MMA_DONE);
  } else if (warp >= 2) {
    const int q = warp & 3, half = (warp - 2) >> 2;
    const bool active = tile < 2 || q == 0;
    mbar_wait(misc + MB_MMA_DONE, 0);
    tc_fence_after_sync();
    if (active) {   // dump raw scores
      for (int i = 0; i < 9; ++i) {
        uint32_t v[16];
        tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + TM_S + half * 144 + i * 16, v);
        tmem_wait_ld();
        for (int j = 0; j < 16; ++j) s_out[(q * 32 + lane) * 288 + half * 144 + i * 16 + j] = __uint_as_float(v[j]);
      }
    }
    epi_softmax(tmem, sb, active, q, half, lane);
    tc_fence_before_sync();
    mbar_arrive(misc + MB_OPS_READY);
    mbar_wait(misc + MB_MMA_DONE, 1);
    tc_fence_after_sync();
    if (active) {
      const int rt = q * 32 + lane;
      const float l = ld_shared_f32(misc + MISC_XSUM + rt * 4) + ld_shared_f32(misc + MISC_XSUM + (128 + rt) * 4);
      uint32_t a[16];
      tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + TM_O + half * 16, a);
      tmem_wait_ld();
      for (int j = 0; j < 16; ++j) o_out[rt * 32 + half * 16 + j] = __uint_as_float(a[j]) / l;
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

float bf16_round_host(float x) {
  uint32_t u;
  memcpy(&u, &x, 4);
  u = (u + 0x7FFFu + ((u >> 16) & 1u)) & 0xFFFF0000u;
  float y;
  memcpy(&y, &u, 4);
  return y;
}
uint16_t bf16_bits_host(float x) {
  uint32_t u;
  memcpy(&u, &x, 4);
  return (uint16_t)((u + 0x7FFFu + ((u >> 16) & 1u)) >> 16);
}
struct Lcg {
  uint64_t s;
  float next() {   // uniform in [-1, 1)
    s = s * 6364136223846793005ull + 1442695040888963407ull;
    return (float)((s >> 40) & 0xFFFFFF) / 8388608.0f - 1.0f;
  }
};

}  // namespace

bool tc_selftest(int which, double* max_err, cudaStream_t st) {
  *max_err = -1.0;
  Lcg rng{12345u + (uint64_t)which};
  if (which == 0) {
    std::vector<float> A(288 * 128), B(96 * 128);
    for (auto& v : A) v = bf16_round_host(rng.next());
    for (auto& v : B) v = bf16_round_host(rng.next());
    std::vector<uint16_t> ai(kXImageBytes / 2, 0), bi(kWInSlice / 2, 0);
    for (int r = 0; r < 288; ++r)
      for (int c = 0; c < 128; ++c) ai[(image_offset(r, c & ~7, kSPad) >> 1) + (c & 7)] = bf16_bits_host(A[r * 128 + c]);
    for (int r = 0; r < 96; ++r)
      for (int c = 0; c < 128; ++c) bi[(image_offset(r, c & ~7, 96) >> 1) + (c & 7)] = bf16_bits_host(B[r * 128 + c]);
    char *da = nullptr, *db = nullptr;
    float* dd = nullptr;
    if (cudaMalloc(&da, kXImageBytes) != cudaSuccess || cudaMalloc(&db, kWInSlice) != cudaSuccess ||
        cudaMalloc(&dd, 128 * 96 * sizeof(float)) != cudaSuccess) {
      set_error("selftest: cudaMalloc failed");
      return false;
    }
    cudaMemcpyAsync(da, ai.data(), kXImageBytes, cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(db, bi.data(), kWInSlice, cudaMemcpyHostToDevice, st);
    const int smem = 110592 + 64 + 1024;
    cudaFuncSetAttribute(selftest_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    selftest_gemm_kernel<<<1, 128, smem, st>>>(da, db, dd);
    count_launch();
    std::vector<float> D(128 * 96);
    cudaMemcpyAsync(D.data(), dd, D.size() * sizeof(float), cudaMemcpyDeviceToHost, st);
    const cudaError_t e = cudaStreamSynchronize(st);
    cudaFree(da); cudaFree(db); cudaFree(dd);
    if (e != cudaSuccess) { set_error("selftest gemm: %s", cudaGetErrorString(e)); return false; }
    double worst = 0.0;
    for (int m = 0; m < 128; ++m)
      for (int n = 0; n < 96; ++n) {
        double ref = 0.0;
        for (int k = 0; k < 128; ++k) ref += (double)A[(128 + m) * 128 + k] * (double)B[n * 128 + k];
        worst = fmax(worst, fabs(ref - (double)D[m * 96 + n]));
      }
    *max_err = worst;
    return true;
  }
  if (which == 1 || which == 2) {
    const int tile = which == 1 ? 1 : 2;
    // Q, K, V [288][32] in SW64 images; values scaled so that scores spread over a few units (log2 domain)
    std::vector<float> Q(288 * 32), K(288 * 32), V(288 * 32);
    for (auto& v : Q) v = bf16_round_host(rng.next() * 1.5f);
    for (auto& v : K) v = bf16_round_host(rng.next() * 1.5f);
    for (auto& v : V) v = bf16_round_host(rng.next());
    std::vector<uint16_t> img(3 * kQkvPart / 2, 0);
    auto put = [&](int part, const std::vector<float>& M) {
      for (int r = 0; r < 288; ++r)
        for (int c = 0; c < 32; ++c) {
          const int u = c >> 3, sw = (r >> 1) & 3;
          img[(part * kQkvPart + r * 64 + ((u ^ sw) << 4)) / 2 + (c & 7)] = bf16_bits_host(M[r * 32 + c]);
        }
    };
    put(0, Q); put(1, K); put(2, V);
    char* di = nullptr;
    float *ds = nullptr, *dO = nullptr;
    if (cudaMalloc(&di, 3 * kQkvPart) != cudaSuccess || cudaMalloc(&ds, 128 * 288 * sizeof(float)) != cudaSuccess ||
        cudaMalloc(&dO, 128 * 32 * sizeof(float)) != cudaSuccess) {
      set_error("selftest: cudaMalloc failed");
      return false;
    }
    cudaMemsetAsync(ds, 0, 128 * 288 * sizeof(float), st);
    cudaMemsetAsync(dO, 0, 128 * 32 * sizeof(float), st);
    cudaMemcpyAsync(di, img.data(), 3 * kQkvPart, cudaMemcpyHostToDevice, st);
    cudaFuncSetAttribute(selftest_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcSmemBytes);
    selftest_attn_kernel<<<1, kTcThreads, kTcSmemBytes, st>>>(di, ds, dO, tile);
    count_launch();
    std::vector<float> S(128 * 288), O(128 * 32);
    cudaMemcpyAsync(S.data(), ds, S.size() * sizeof(float), cudaMemcpyDeviceToHost, st);
    cudaMemcpyAsync(O.data(), dO, O.size() * sizeof(float), cudaMemcpyDeviceToHost, st);
    const cudaError_t e = cudaStreamSynchronize(st);
    cudaFree(di); cudaFree(ds); cudaFree(dO);
    if (e != cudaSuccess) { set_error("selftest attn: %s", cudaGetErrorString(e)); return false; }
    double worst_s = 0.0, worst_o = 0.0;
    const int rows = tile < 2 ? 128 : 24;
    for (int m = 0; m < rows; ++m) {
      const int qi = tile * 128 + m;
      std::vector<double> sc(280);
      double mx = -1e30;
      for (int j = 0; j < 280; ++j) {
        double s = 0.0;
        for (int c = 0; c < 32; ++c) s += (double)Q[qi * 32 + c] * (double)K[j * 32 + c];
        sc[j] = s;
        mx = fmax(mx, s);
        worst_s = fmax(worst_s, fabs(s - (double)S[m * 288 + j]));
      }
      double l = 0.0;
      std::vector<double> o(32, 0.0);
      for (int j = 0; j < 280; ++j) {
        const double pj = exp2(sc[j] - mx);
        l += pj;
        for (int c = 0; c < 32; ++c) o[c] += pj * (double)V[j * 32 + c];
      }
      for (int c = 0; c < 32; ++c) worst_o = fmax(worst_o, fabs(o[c] / l - (double)O[m * 32 + c]));
    }
    // scores must be exact to fp32 accumulation; outputs carry the bf16 rounding of P (~2^-9 relative)
    *max_err = fmax(worst_s, worst_o);
    if (worst_s > 1e-3) { set_error("selftest attn tile %d: scores off by %g (outputs %g)", tile, worst_s, worst_o); }
    return true;
  }
  set_error("aft_selftest: unknown test %d", which);
  return false;
}

}  // namespace aft
