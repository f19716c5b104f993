// AFT_BF16 path, encoder kernel v3: the 6-layer post-norm transformer encoder (reference src/models/blocks/encoders.py:44-55,69
// -> torch _transformer_encoder_layer_fwd) as one persistent kernel with THREE ASYNCHRONOUS ROW-TILE STREAMS per CTA.
//
// v2 (tc_encoder.cu) walks one sequence with 16 compute warps in lock-step: every accumulator row is split over four
// threads (two shared-memory exchanges + named barriers per softmax / LayerNorm), and every hand-off compute warps -> MMA
// issuer -> tensor pipe -> compute warps is exposed (ncu: issue slots 48 % busy, MUFU 24 %, tensor pipe 24 %, nothing
// saturated).  v3 keeps the sequence resident in shared memory in the same operand images but cuts the work by ROW TILE:
//
//   stream 0 : token rows   0..127   compute warps 0..3  + MMA issuer warp 13
//   stream 1 : token rows 128..255   compute warps 4..7  + MMA issuer warp 14
//   stream 2 : token rows 256..279   compute warp 8       + MMA issuer warp 15   (the 24-row tail: M = 128 MMAs whose rows
//              280.. read whatever follows in shared memory -- rows of A are independent, those accumulator lanes are never read)
//   warp 12  : producer (bulk copies of weights / the sequence image, result image back to global memory)
//
// ONE THREAD OWNS ONE TOKEN ROW in every epilogue: softmax statistics, LayerNorm sums, 1/l are thread-private -- no
// exchange, no named barrier.  Each stream has its own MMA issuer, its own 160 TMEM columns and its own mbarriers and
// runs its own static program; the streams only meet where the data flow joins them: K / V of a head are written by all
// streams and read by all (QKV_READY / QKV_FREE), weights are shared (EMPTY barriers count three consumers).  While one
// stream waits for the tensor pipe the other two fill the issue slots / MUFU.
//
// Attention of (stream, head): streaming softmax over 64-key tiles (5 tiles: 4 x 64 + 24 keys), two score buffers, P
// written IN PLACE over the scores (bf16 pairs, A operand of P.V from TMEM), output accumulator resident in TMEM and
// rescaled only when a row maximum outgrows its reference by 2^8 (the scheme of tc_long.cu's attention).
// FFN: 8 chunks of 32 hidden units: FFN1 chunk -> TMEM -> GELU -> hidden image (shared memory) -> FFN2 partial product.
//
// Shared memory map: identical to tc_encoder.cu (O | X | QKV / weight ring | W | MISC); stream t's hidden buffers are the
// rows of ITS tile inside the two chunks of the O image (dead once out_proj of the tile has completed).
// Tensor memory: stream s owns columns [160 s, 160 s + 160):  S0 [0,64) S1 [64,128) O [128,160)  |  QKV accumulators
// [32,128)  |  out_proj / FFN2 accumulator [0,128), FFN1 chunk [128,160).
#include <cstdio>
#include <cstdlib>

#include "tc_encoder.cuh"
#include "tc_layout.cuh"
#include "tc_math.cuh"
#include "tc_ptx.cuh"

namespace aft {

namespace {

using namespace ptx;
using namespace tcm;

#ifndef AFT_V3_POLY_MASK
#define AFT_V3_POLY_MASK 0x88   // bit jj set: pair jj of every 8 pairs of exponentials runs on the FMA pipe (packed Cody-Waite + cubic), the rest on the MUFU
#endif

constexpr int kThreads3 = 512;
// 384 x 144 + 128 x 80 = 65,536.  The MMA issuers run long static programs: with 56 registers they spilled, and every
// reload sits on a hand-off (measured: 229.8 vs 233.8 ms per 65536 estimates)
constexpr int kRegsCompute3 = 144, kRegsCtrl3 = 80;
constexpr int kProducerWarp3 = 12, kMmaWarp0 = 13;
constexpr float kRescale3 = 8.0f;

// per-layer epilogue vectors (layout of tc_encoder.cu's pack_vec_kernel)
constexpr int kVecBlock3 = 384;
constexpr int kVBOut = 0, kVBL1 = 128, kVBL2 = 384, kVN1W = 512, kVN1B = 640, kVN2W = 768, kVN2B = 896;
constexpr uint32_t kBiasBytes3 = 96 * 4, kVecBytes3 = 1024 * 4;

constexpr uint32_t OFF_O = 0, OFF_X = 73728, OFF_QKV = 147456, OFF_W = 202752, OFF_MISC = 227328;
constexpr uint32_t kQkvPart = 18432, kSlot = 16384, kWInSlice = 24576;
constexpr uint32_t OFF_VEC = OFF_QKV + 3 * kSlot;
constexpr uint32_t kSmem3 = OFF_MISC + 5120;   // 232,448
constexpr uint32_t MISC_BIAS = 0, MISC_BARS = 768, MISC_TMEM = 1344;
static_assert(MISC_BARS + 128 + 3 * 144 <= MISC_TMEM, "barrier block");

// mbarriers (byte offsets from the barrier block).  "commit" = completed by tcgen05.commit / expect_tx, "warps" = one
// arrival per compute warp.  Protocol rule (tc_ptx.cuh / DESIGN.md): a waiter tests phase parity, so completion k + 1 of a
// barrier must causally depend on every waiter having passed its wait for completion k - 1.  Barriers that a consumer may
// leave two completions behind are doubled (S_DONE, P_READY, HID_READY, F2_DONE) and used alternately.
enum : uint32_t {
  B_X_FULL = 0,        // commit : sequence image landed
  B_X_DONE = 8,        // warps(9) : last LayerNorm of the sequence written
  B_ATTN_DONE = 16,    // 3 commits: all P.V of the layer complete (producer: ring / vector block may overwrite Q/K/V)
  B_QKV_READY = 24,    // warps(9) : Q/K/V rows of head g written by every stream
  B_QKV_FREE = 32,     // 3 arrivals: Q/K/V region free -- 5 completions per layer: [layer start], head 0..3 done
  B_VEC_FULL = 40,     // commit
  B_BIAS_FULL = 48,    // 2 x commit
  B_W_FULL = 64,       // 4 x commit : [0..2] ring slots, [3] in_proj slot
  B_W_EMPTY = 96,      // 4 x 3 commits
  B_STREAM = 128,      // per-stream blocks of kStreamBars3 bytes
};
enum : uint32_t {
  S_QKV_DONE = 0, S_S_DONE = 8 /* 2 */, S_P_READY = 24 /* 2 */, S_PV_DONE = 128 /* 2 */, S_PROJ_OK = 40, S_O_READY = 48, S_OUT_DONE = 56, S_X1_READY = 64,
  S_F1_DONE = 72, S_F1_FREE = 80, S_HID_READY = 88 /* 2 */, S_F2_DONE = 104 /* 2 */, S_X2_READY = 120,
};
constexpr uint32_t kSlotIn = 3, kStreamBars3 = 144;

constexpr uint32_t kIdQkv = make_idesc_bf16(128, 96, false, false);
constexpr uint32_t kIdS64 = make_idesc_bf16(128, 64, false, false), kIdS32 = make_idesc_bf16(128, 32, false, false);
constexpr uint32_t kIdPV = make_idesc_bf16(128, 32, false, true);
constexpr uint32_t kIdN128 = make_idesc_bf16(128, 128, false, false), kIdN32 = make_idesc_bf16(128, 32, false, false);

constexpr uint32_t kHi128 = (uint32_t)(desc_k_sw128_const() >> 32);
__device__ __forceinline__ uint64_t d128(uint32_t saddr, int ks) {
  return ((uint64_t)kHi128 << 32) | (((uint32_t)desc_k_sw128_const() | ((saddr >> 4) & 0x3FFF)) + (uint32_t)ks * 2);
}
constexpr uint32_t kHi64 = (uint32_t)(((uint64_t)(512 >> 4)) | ((uint64_t)1 << 14) | ((uint64_t)kSwizzle64 << 29));
__device__ __forceinline__ uint32_t lo_k64(uint32_t saddr) { return ((saddr >> 4) & 0x3FFF) | ((16u >> 4) << 16); }
__device__ __forceinline__ uint32_t lo_mn64(uint32_t saddr) { return ((saddr >> 4) & 0x3FFF) | ((512u >> 4) << 16); }
__device__ __forceinline__ uint64_t d64(uint32_t lo) { return ((uint64_t)kHi64 << 32) | lo; }

// D (+)= A . B^T over K = 128: K-chunk 0 at (a0, b0), K-chunk 1 at (a1, b1); SW128 K-major images
__device__ __forceinline__ void gemm_k128(uint32_t d, uint32_t a0, uint32_t a1, uint32_t b0, uint32_t b1, uint32_t idesc, bool el) {
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) mma_ss(d, d128(a0, ks), d128(b0, ks), idesc, ks > 0, el);
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) mma_ss(d, d128(a1, ks), d128(b1, ks), idesc, true, el);
}

__device__ __forceinline__ void tmem_st8p(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]));
}

#ifdef AFT_V3_MMA_POLL
#define MMA_WAIT3 mbar_wait_poll
#else
#define MMA_WAIT3 mbar_wait_spin
#endif

struct Enc3Params {
  unsigned long long* timeline;   // -DAFT_V3_TIMELINE builds: [3][256] (id << 48 | clock) records of block 0 (compute warp 0, compute warp 4, MMA issuer 0)
  char* x_images;
  const TcLayer* layers;
  int num_layers;
  int activation;
  int64_t nseq;
};

#ifdef AFT_V3_TIMELINE
#define TL3(slot, id)                                                                                       \
  do {                                                                                                      \
    if (tl_on && tl_n < 255) p.timeline[(slot) * 256 + (++tl_n)] = ((unsigned long long)(id) << 48) | (clock64() & 0xFFFFFFFFFFFFull); \
  } while (0)
#else
#define TL3(slot, id) do { } while (0)
#endif

// ---------------------------------------------------------------------------------------------
// compute-warp epilogues: the thread owns token row r (TMEM lane 32 q + lane of its stream's columns)
// ---------------------------------------------------------------------------------------------
// QKV accumulators [q_g | k_g | v_g] (96 columns) + bias -> bf16 -> row r of the Q / K / V images (64-byte rows, SWIZZLE_64B)
__device__ __forceinline__ void epi_qkv3(uint32_t taddr, uint32_t sb, uint32_t bias96, int r) {
  uint32_t acc[96];
  tmem_ld_cols(taddr, acc);
  tmem_wait_ld();
  const int sw = (r >> 1) & 3;
#pragma unroll
  for (int mat = 0; mat < 3; ++mat)
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float4 b0 = lds_f4(bias96 + (mat * 32 + u * 8) * 4), b1 = lds_f4(bias96 + (mat * 32 + u * 8) * 4 + 16);
      const uint32_t* a = acc + mat * 32 + u * 8;
      auto sum = [](uint32_t x0, uint32_t x1, float y0, float y1) {
        return pack_bf16_pair(add2(pack2(__uint_as_float(x0), __uint_as_float(x1)), pack2(y0, y1)));
      };
      st_shared_v4(sb + OFF_QKV + mat * kQkvPart + r * 64 + ((u ^ sw) << 4), sum(a[0], a[1], b0.x, b0.y), sum(a[2], a[3], b0.z, b0.w),
                   sum(a[4], a[5], b1.x, b1.y), sum(a[6], a[7], b1.z, b1.w));
    }
}

// One half tile (32 keys) of the streaming softmax: scores x -> running maximum -> exponentials against the reference maximum
// -> P (bf16 pairs) in place at p_addr (16 columns) -> row sum.  `first`: first half tile of the head; kMask: the last 32
// keys of the sequence, of which 24 exist.  The output accumulator is rescaled only when some row's maximum outgrew its
// reference by 2^8 (rare): P.V of the previous key tile must have completed then (`pv_waited` tells the caller that the
// wait has happened).
template <bool kMask, bool kSecond>
__device__ __forceinline__ void softmax_half3(const uint32_t (&x)[32], uint32_t p_addr, uint32_t o_addr, bool first, bool has_prev_pv,
                                              uint32_t pv_bar, uint32_t pv_par, bool& pv_waited, float& m_ref, float& lsum,
                                              unsigned long long* dbg = nullptr) {
#ifdef AFT_V3_TIMELINE
#define DBG3(i) do { if (dbg) dbg[i] = clock64(); } while (0)
#else
#define DBG3(i) do { } while (0)
#endif
  DBG3(0);
  float v[32];
#pragma unroll
  for (int c = 0; c < 32; ++c) v[c] = __uint_as_float(x[c]);
  if (kMask) {
#pragma unroll
    for (int c = kS - 256; c < 32; ++c) v[c] = -INFINITY;   // keys 280..287 are padding
  }
  float m0 = v[0], m1 = v[1], m2 = v[2], m3 = v[3];
#pragma unroll
  for (int c = 4; c < 32; c += 4) { m0 = fmaxf(m0, v[c]); m1 = fmaxf(m1, v[c + 1]); m2 = fmaxf(m2, v[c + 2]); m3 = fmaxf(m3, v[c + 3]); }
  const float mt = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
  DBG3(1);
  if (__any_sync(0xFFFFFFFFu, mt > m_ref + kRescale3)) {
    const float mn = fmaxf(m_ref, mt);
    const float f = ex2(m_ref - mn);   // first half tile: exp2(-inf) = 0
    if (!first) {
      if (has_prev_pv && !pv_waited) {
        mbar_wait_spin(pv_bar, pv_par);
        tc_fence_after_sync();
        pv_waited = true;
      }
      // (the P.V of THIS key tile cannot be in flight: it is issued after both halves of P have been announced)
      uint32_t a[32];
      tmem_ld_cols(o_addr, a);
      tmem_wait_ld();
#pragma unroll
      for (int i = 0; i < 32; ++i) a[i] = __float_as_uint(__uint_as_float(a[i]) * f);
#pragma unroll
      for (int i = 0; i < 4; ++i) tmem_st8p(o_addr + i * 8, a + 8 * i);
      if (kSecond) {
        // the first half of this key tile has already stored its P against the old reference (P.V of the tile is issued after
        // both halves): bring those 16 columns to the new one as well
        uint32_t pp[16];
        tmem_ld16(p_addr - 16, pp);
        tmem_wait_ld();
        const f32x2 f2 = pack2(f, f);
#pragma unroll
        for (int i = 0; i < 16; ++i) pp[i] = pack_bf16_pair(mul2(bf16x2_to_f32x2(pp[i]), f2));
        tmem_st8p(p_addr - 16, pp);
        tmem_st8p(p_addr - 8, pp + 8);
      }
      tmem_wait_st();   // the next half tile may rescale again: its loads must see these stores
    }
    lsum *= f;
    m_ref = mn;
  }
  DBG3(2);
  const f32x2 negm2 = pack2(-m_ref, -m_ref);
  f32x2 s2a = pack2(0.f, 0.f), s2b = pack2(0.f, 0.f);
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    uint32_t pk[8];
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) {
      const int c = i * 16 + 2 * jj;
      const f32x2 x2 = add2(pack2(v[c], v[c + 1]), negm2);
      f32x2 e2;
      if (((AFT_V3_POLY_MASK) >> jj) & 1) {
        e2 = ex2_poly2(x2);
      } else {
        float a, b;
        unpack2(x2, a, b);
        e2 = pack2(ex2(a), ex2(b));
      }
      if (jj & 1) s2b = add2(s2b, e2); else s2a = add2(s2a, e2);
      pk[jj] = pack_bf16_pair(e2);
    }
    DBG3(3 + 2 * i);
    tmem_st8(p_addr + i * 8, pk);
    DBG3(4 + 2 * i);
  }
  float sa, sb2, sc, sd;
  unpack2(s2a, sa, sb2);
  unpack2(s2b, sc, sd);
  lsum += (sa + sb2) + (sc + sd);
}

// output accumulator (32 columns of head g) / l -> bf16 -> O image row r
__device__ __forceinline__ void epi_o3(uint32_t o_addr, uint32_t sb, int g, int r, float lsum, bool valid) {
  uint32_t a[32];
  tmem_ld_cols(o_addr, a);
  tmem_wait_ld();
  const float inv = rcp_approx(lsum);
  const f32x2 il = pack2(inv, inv);
  const uint32_t row = sb + OFF_O + (g >> 1) * kXChunkBytes + r * 128;
  if (valid) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      auto sc = [&](int j) { return pack_bf16_pair(mul2(pack2(__uint_as_float(a[8 * u + j]), __uint_as_float(a[8 * u + j + 1])), il)); };
      st_shared_v4(row + ((((g & 1) * 4 + u) ^ (r & 7)) << 4), sc(0), sc(2), sc(4), sc(6));
    }
  }
}

// accumulator (128 columns) + bias + residual (X row r) -> LayerNorm -> X row r in place.  Two passes over the
// accumulator columns: the pre-norm values go back to TMEM between them instead of pinning 128 registers.
__device__ __forceinline__ void epi_ln3(uint32_t acc_addr, uint32_t sb, uint32_t vec, int which, int r, bool valid) {
  const uint32_t bias = vec + 4 * (which == 1 ? kVBOut : kVBL2);
  const uint32_t gam = vec + 4 * (which == 1 ? kVN1W : kVN2W);
  const uint32_t bet = vec + 4 * (which == 1 ? kVN1B : kVN2B);
  f32x2 s2 = pack2(0.f, 0.f), q2 = pack2(0.f, 0.f);
  // blocks of 32 columns, the next block's accumulators in flight while the current one is processed
  uint32_t buf[2][32];
  tmem_ld_cols(acc_addr, buf[0]);
#pragma unroll
  for (int blk = 0; blk < 4; ++blk) {
    uint32_t (&acc)[32] = buf[blk & 1];
    const uint32_t xrow = sb + OFF_X + (blk >> 1) * kXChunkBytes + r * 128;
    uint4 xr[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) xr[u] = ld_shared_v4(xrow + ((((blk & 1) * 4 + u) ^ (r & 7)) << 4));
    tmem_wait_ld();
    if (blk < 3) tmem_ld_cols(acc_addr + (blk + 1) * 32, buf[(blk + 1) & 1]);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float4 b0 = lds_f4(bias + (blk * 32 + u * 8) * 4), b1 = lds_f4(bias + (blk * 32 + u * 8) * 4 + 16);
      const uint32_t xw[4] = {xr[u].x, xr[u].y, xr[u].z, xr[u].w};
      const f32x2 bb[4] = {pack2(b0.x, b0.y), pack2(b0.z, b0.w), pack2(b1.x, b1.y), pack2(b1.z, b1.w)};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const f32x2 a2 = pack2(__uint_as_float(acc[u * 8 + 2 * j]), __uint_as_float(acc[u * 8 + 2 * j + 1]));
        const f32x2 y = add2(add2(a2, bb[j]), bf16x2_to_f32x2(xw[j]));
        s2 = add2(s2, y);
        q2 = fma2(y, y, q2);
        float ya, yb;
        unpack2(y, ya, yb);
        acc[u * 8 + 2 * j] = __float_as_uint(ya);
        acc[u * 8 + 2 * j + 1] = __float_as_uint(yb);
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) tmem_st8p(acc_addr + blk * 32 + i * 8, acc + 8 * i);
  }
  tmem_wait_st();
  tmem_ld_cols(acc_addr, buf[0]);
  float sa, sb2, qa, qb;
  unpack2(s2, sa, sb2);
  unpack2(q2, qa, qb);
  const float mean = (sa + sb2) * (1.0f / 128.0f);
  const float var = fmaxf(fmaf(-mean, mean, (qa + qb) * (1.0f / 128.0f)), 0.f);
  const float rstd = rsqrtf(var + 1e-5f);
  const f32x2 rstd2 = pack2(rstd, rstd), shift2 = pack2(-mean * rstd, -mean * rstd);
#pragma unroll
  for (int blk = 0; blk < 4; ++blk) {
    uint32_t (&y)[32] = buf[blk & 1];
    const uint32_t xrow = sb + OFF_X + (blk >> 1) * kXChunkBytes + r * 128;
    tmem_wait_ld();
    if (blk < 3) tmem_ld_cols(acc_addr + (blk + 1) * 32, buf[(blk + 1) & 1]);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float4 g0 = lds_f4(gam + (blk * 32 + u * 8) * 4), g1 = lds_f4(gam + (blk * 32 + u * 8) * 4 + 16);
      const float4 e0 = lds_f4(bet + (blk * 32 + u * 8) * 4), e1 = lds_f4(bet + (blk * 32 + u * 8) * 4 + 16);
      const f32x2 gg[4] = {pack2(g0.x, g0.y), pack2(g0.z, g0.w), pack2(g1.x, g1.y), pack2(g1.z, g1.w)};
      const f32x2 ee[4] = {pack2(e0.x, e0.y), pack2(e0.z, e0.w), pack2(e1.x, e1.y), pack2(e1.z, e1.w)};
      uint32_t pk[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const f32x2 yy = pack2(__uint_as_float(y[u * 8 + 2 * j]), __uint_as_float(y[u * 8 + 2 * j + 1]));
        pk[j] = pack_bf16_pair(fma2(fma2(yy, rstd2, shift2), gg[j], ee[j]));
      }
      // padding rows 280..287 stay zero (they feed the padding keys of the next layer)
      if (valid) st_shared_v4(xrow + ((((blk & 1) * 4 + u) ^ (r & 7)) << 4), pk[0], pk[1], pk[2], pk[3]);
    }
  }
}

// FFN1 chunk c (32 hidden units) + bias -> GELU / ReLU -> bf16 pairs
__device__ __forceinline__ void act_chunk3(const uint32_t (&a)[32], uint32_t vec, int c, int act, uint32_t (&pk)[16]) {
  const uint32_t bias = vec + 4 * (kVBL1 + c * 32);
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const float4 b0 = lds_f4(bias + u * 32), b1 = lds_f4(bias + u * 32 + 16);
    const f32x2 bb[4] = {pack2(b0.x, b0.y), pack2(b0.z, b0.w), pack2(b1.x, b1.y), pack2(b1.z, b1.w)};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const f32x2 f = add2(pack2(__uint_as_float(a[u * 8 + 2 * j]), __uint_as_float(a[u * 8 + 2 * j + 1])), bb[j]);
      if (act == AFT_ACT_GELU) {
        pk[u * 4 + j] = pack_bf16_pair(gelu_tanh2(f));
      } else {
        float x, y;
        unpack2(f, x, y);
        pk[u * 4 + j] = pack_bf16x2(fmaxf(x, 0.f), fmaxf(y, 0.f));
      }
    }
  }
}

// =============================================================================================
__global__ void __launch_bounds__(kThreads3, 1) encoder3_kernel(Enc3Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t sb = smem_u32(smem_raw);
  if ((sb & 1023u) != 0) __trap();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t miscb = sb + OFF_MISC, bars = miscb + MISC_BARS;

  if (threadIdx.x == 0) {
    mbar_init(bars + B_X_FULL, 1);
    mbar_init(bars + B_X_DONE, 9);
    mbar_init(bars + B_ATTN_DONE, 3);
    mbar_init(bars + B_QKV_READY, 9);
    mbar_init(bars + B_QKV_FREE, 3);
    mbar_init(bars + B_VEC_FULL, 1);
    mbar_init(bars + B_BIAS_FULL, 1);
    mbar_init(bars + B_BIAS_FULL + 8, 1);
    for (int i = 0; i < 4; ++i) { mbar_init(bars + B_W_FULL + 8 * i, 1); mbar_init(bars + B_W_EMPTY + 8 * i, 3); }
    for (int s = 0; s < 3; ++s) {
      const uint32_t b = bars + B_STREAM + kStreamBars3 * s, nw = s < 2 ? 4 : 1;
      mbar_init(b + S_QKV_DONE, 1);
      mbar_init(b + S_S_DONE, 1); mbar_init(b + S_S_DONE + 8, 1);
      mbar_init(b + S_P_READY, nw); mbar_init(b + S_P_READY + 8, nw);
      mbar_init(b + S_PV_DONE, 1); mbar_init(b + S_PV_DONE + 8, 1);
      mbar_init(b + S_PROJ_OK, 1);
      mbar_init(b + S_O_READY, nw);
      mbar_init(b + S_OUT_DONE, 1);
      mbar_init(b + S_X1_READY, nw);
      mbar_init(b + S_F1_DONE, 1);
      mbar_init(b + S_F1_FREE, nw);
      mbar_init(b + S_HID_READY, nw); mbar_init(b + S_HID_READY + 8, nw);
      mbar_init(b + S_F2_DONE, 1); mbar_init(b + S_F2_DONE + 8, 1);
      mbar_init(b + S_X2_READY, nw);
    }
    fence_mbar_init();
  }
  if (warp == kMmaWarp0) { tmem_alloc(miscb + MISC_TMEM, 512); tmem_relinquish(); }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  uint32_t tmem;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(miscb + MISC_TMEM));

  const int L = p.num_layers;
  const int nseq = (int)p.nseq;

  if (warp >= 12) {
    setmaxnreg_dec<kRegsCtrl3>();
    if (warp == kProducerWarp3) {
      if (lane == 0) {
        // ----------------------------------------------------------------------------- producer
        uint32_t n_in = 0, n_ring = 0, n_seq = 0, n_attn = 0;
        auto fill_in = [&](const char* src, uint32_t bytes) {
          if (n_in > 0) mbar_wait_relaxed(bars + B_W_EMPTY + 8 * kSlotIn, (n_in - 1) & 1);
          mbar_arrive_expect_tx(bars + B_W_FULL + 8 * kSlotIn, bytes);
          bulk_g2s(sb + OFF_W, src, bytes, bars + B_W_FULL + 8 * kSlotIn);
          ++n_in;
        };
        auto fill_ring = [&](const char* src) {
          const uint32_t slot = n_ring % 3, fill = n_ring / 3;
          if (fill > 0) mbar_wait_relaxed(bars + B_W_EMPTY + 8 * slot, (fill - 1) & 1);
          mbar_arrive_expect_tx(bars + B_W_FULL + 8 * slot, kSlot);
          bulk_g2s(sb + OFF_QKV + slot * kSlot, src, kSlot, bars + B_W_FULL + 8 * slot);
          ++n_ring;
        };
        int prev_seq = -1;
        for (int seq = blockIdx.x; seq < nseq; seq += gridDim.x, ++n_seq) {
          if (prev_seq >= 0) {
            // the finished sequence replaces its input image in global memory, then the next image may land
            mbar_wait_relaxed(bars + B_X_DONE, (n_seq - 1) & 1);
            bulk_s2g(p.x_images + prev_seq * (int64_t)kXImageBytes, sb + OFF_X, kXImageBytes);
            bulk_wait_read();
          }
          prev_seq = seq;
          mbar_arrive_expect_tx(bars + B_X_FULL, kXImageBytes);
          bulk_g2s(sb + OFF_X, p.x_images + seq * (int64_t)kXImageBytes, kXImageBytes, bars + B_X_FULL);
          if (seq + (int)gridDim.x < nseq) bulk_prefetch_l2(p.x_images + (seq + gridDim.x) * (int64_t)kXImageBytes, kXImageBytes);
          for (int l = 0; l < L; ++l) {
            const TcLayer& W = p.layers[l];
            for (int g = 0; g < 4; ++g) {
              fill_in(reinterpret_cast<const char*>(W.w_in) + g * kWInSlice, kWInSlice);
              // bias buffer g & 1 was last read by the QKV epilogues of head g - 2; the slot wait above implies that every
              // stream has finished the attention of head g - 2
              mbar_arrive_expect_tx(bars + B_BIAS_FULL + 8 * (g & 1), kBiasBytes3);
              bulk_g2s(miscb + MISC_BIAS + (g & 1) * kBiasBytes3, W.b_in + g * 96, kBiasBytes3, bars + B_BIAS_FULL + 8 * (g & 1));
            }
            mbar_wait_relaxed(bars + B_ATTN_DONE, n_attn & 1);   // Q/K/V images dead
            ++n_attn;
            mbar_arrive_expect_tx(bars + B_VEC_FULL, kVecBytes3);
            bulk_g2s(sb + OFF_VEC, W.b_in + kVecBlock3, kVecBytes3, bars + B_VEC_FULL);
            const char* wout = reinterpret_cast<const char*>(W.w_out);
            const char* wl1 = reinterpret_cast<const char*>(W.w_l1);
            const char* wl2 = reinterpret_cast<const char*>(W.w_l2);
            fill_ring(wout);
            fill_ring(wout + kSlot);
            fill_ring(wl1);
            fill_in(wl1 + kSlot, kSlot);
            fill_ring(wl2);
            fill_ring(wl2 + kSlot);
            fill_ring(wl1 + 2 * kSlot);
            fill_in(wl1 + 3 * kSlot, kSlot);
            fill_ring(wl2 + 2 * kSlot);
            fill_ring(wl2 + 3 * kSlot);
          }
        }
        if (prev_seq >= 0) {
          mbar_wait_relaxed(bars + B_X_DONE, (n_seq - 1) & 1);
          bulk_s2g(p.x_images + prev_seq * (int64_t)kXImageBytes, sb + OFF_X, kXImageBytes);
          bulk_wait_all();
        }
      }
    } else {
      // ----------------------------------------------------------------------------- MMA issuer of stream s
      // The whole warp runs the schedule converged; only tcgen05.mma / commit are predicated on the elected lane.
      const int s = warp - kMmaWarp0;
      const bool el = elect_one();
      const uint32_t tm = tmem + 160 * s;
      uint32_t n_seq = 0, Lg = 0, hg = 0, n_in = 0, ring_base = 0, n_proj = 0;
      const uint32_t sbar = bars + B_STREAM + kStreamBars3 * s;
      const int arow = s < 2 ? 128 * s : 256;   // first token row of the stream's M = 128 tile
      auto ring_wait = [&](uint32_t idx) -> uint32_t {
        MMA_WAIT3(bars + B_W_FULL + 8 * (idx % 3), (idx / 3) & 1);
        tc_fence_after_sync();
        return sb + OFF_QKV + (idx % 3) * kSlot;
      };
      auto ring_release = [&](uint32_t idx) { mma_commit(bars + B_W_EMPTY + 8 * (idx % 3), el); };
      for (int seq = blockIdx.x; seq < nseq; seq += gridDim.x, ++n_seq) {
        MMA_WAIT3(bars + B_X_FULL, n_seq & 1);
        for (int l = 0; l < L; ++l, ++Lg, ring_base += 8) {
#ifdef AFT_V3_TIMELINE
          const bool tl_on = blockIdx.x == 0 && n_seq == 1 && l == 1 && lane == 0 && s == 0;
          uint32_t tl_n = 0;
#endif
          // layer start: the stream's LayerNorm2 of the previous layer is written (X rows, accumulator columns and the
          // vector block are free as far as this stream is concerned)
          if (Lg > 0) MMA_WAIT3(sbar + S_X2_READY, (Lg - 1) & 1);
          tc_fence_after_sync();
          if (lane == 0) mbar_arrive(bars + B_QKV_FREE);
          __syncwarp();
          auto issue_qkv = [&]() {
            MMA_WAIT3(bars + B_W_FULL + 8 * kSlotIn, n_in & 1);
            tc_fence_after_sync();
            const uint32_t a0 = sb + OFF_X + arow * 128;
            gemm_k128(tm + 32, a0, a0 + kXChunkBytes, sb + OFF_W, sb + OFF_W + 96 * 128, kIdQkv, el);
            mma_commit(bars + B_W_EMPTY + 8 * kSlotIn, el);
            mma_commit(sbar + S_QKV_DONE, el);
            ++n_in;
          };
          issue_qkv();
          for (int g = 0; g < 4; ++g, ++hg) {
            const uint32_t hv = hg;
            MMA_WAIT3(bars + B_QKV_READY, hg & 1);
            tc_fence_after_sync();
            TL3(2, g * 100 + 1);
            const uint32_t qd = lo_k64(sb + OFF_QKV + arow * 64);
            const uint32_t kd = lo_k64(sb + OFF_QKV + kQkvPart);
            const uint32_t vd = lo_mn64(sb + OFF_QKV + 2 * kQkvPart);
            auto issue_s = [&](int j) {
              const uint32_t d = tm + 64 * (j & 1);
#pragma unroll
              for (int ks = 0; ks < 2; ++ks) mma_ss(d, d64(qd + ks * 2), d64(kd + j * 256 + ks * 2), j < 4 ? kIdS64 : kIdS32, ks > 0, el);
              mma_commit(sbar + S_S_DONE + 8 * (j & 1), el);
            };
            issue_s(0);
            issue_s(1);
#pragma unroll 1
            for (int j = 0; j < 5; ++j) {
              const uint32_t pidx = (j & 1) ? 2 * hv + (j >> 1) : 3 * hv + (j >> 1);
              MMA_WAIT3(sbar + S_P_READY + 8 * (j & 1), pidx & 1);
              tc_fence_after_sync();
              TL3(2, g * 100 + 10 + j);
              const int nks = j < 4 ? 4 : 2;
              for (int ks = 0; ks < nks; ++ks)
                mma_ts(tm + 128, tm + 64 * (j & 1) + ks * 8, d64(vd + j * 256 + ks * 64), kIdPV, j > 0 || ks > 0, el);
              mma_commit(sbar + S_PV_DONE + 8 * (j & 1), el);
              if (j == 3 && g < 3) mma_commit(sbar + S_PROJ_OK, el);   // this warp's own barrier: it waits for EVERY completion
              // S(j + 2) overwrites the buffer that holds P(j): issued right behind P.V(j) -- tcgen05.mma instructions of one
              // thread execute in issue order, so P.V(j) has read its A operand before the new scores land
              if (j + 2 <= 4) issue_s(j + 2);
              TL3(2, g * 100 + 20 + j);
              if (j == 3 && g < 3) {
                // columns [32, 128) (upper half of buffer 0: the last key tile has 32 keys; buffer 1: P(3)) are free once
                // P.V(3) has completed: the next head's projection runs under the last key tile
                MMA_WAIT3(sbar + S_PROJ_OK, n_proj & 1);
                ++n_proj;
                tc_fence_after_sync();
                TL3(2, g * 100 + 30);
                issue_qkv();
                TL3(2, g * 100 + 31);
              }
            }
            // every P.V of this head by this stream has been issued: arrival when they complete
            mma_commit(bars + B_QKV_FREE, el);
            if (g == 3) mma_commit(bars + B_ATTN_DONE, el);
          }
          // ---- linear part of the layer on this stream's row tile
          const uint32_t lv = Lg;
          MMA_WAIT3(sbar + S_O_READY, Lg & 1);
          tc_fence_after_sync();
          TL3(2, 1000);
          {
            const uint32_t w0 = ring_wait(ring_base + 0), w1 = ring_wait(ring_base + 1);
            const uint32_t a0 = sb + OFF_O + arow * 128;
            TL3(2, 1001);
            gemm_k128(tm, a0, a0 + kXChunkBytes, w0, w1, kIdN128, el);
            mma_commit(sbar + S_OUT_DONE, el);
            TL3(2, 1002);
            ring_release(ring_base + 0);
            ring_release(ring_base + 1);
          }
          MMA_WAIT3(sbar + S_X1_READY, lv & 1);
          tc_fence_after_sync();
          TL3(2, 1003);
          {
            const uint32_t xa0 = sb + OFF_X + arow * 128, xa1 = xa0 + kXChunkBytes;
            uint32_t w1a = 0, w1b = 0, w2 = 0;
            auto issue_f2 = [&](int c) {   // FFN2 partial product over hidden units 32 c .. 32 c + 31
              MMA_WAIT3(sbar + S_HID_READY + 8 * (c & 1), (c >> 1) & 1);
              tc_fence_after_sync();
              TL3(2, 1030 + c);
              const int ch = c >> 1;   // 64-unit chunk = W2 K-chunk = hidden buffer ch & 1
              if ((c & 1) == 0) w2 = ring_wait(ring_base + (ch < 2 ? 3 + ch : 4 + ch));
              const uint32_t ha = sb + OFF_O + (ch & 1) * kXChunkBytes + arow * 128;
#pragma unroll
              for (int i = 0; i < 2; ++i) mma_ss(tm, d128(ha, 2 * (c & 1) + i), d128(w2, 2 * (c & 1) + i), kIdN128, c > 0 || i > 0, el);
              if (c & 1) {
                mma_commit(sbar + S_F2_DONE + 8 * (ch & 1), el);
                ring_release(ring_base + (ch < 2 ? 3 + ch : 4 + ch));
              }
            };
#pragma unroll 1
            for (int c = 0; c < 8; ++c) {
              if ((c & 3) == 0) {
                w1a = ring_wait(ring_base + (c == 0 ? 2 : 5));
                MMA_WAIT3(bars + B_W_FULL + 8 * kSlotIn, n_in & 1);
                tc_fence_after_sync();
                w1b = sb + OFF_W;
              }
              if (c > 0) MMA_WAIT3(sbar + S_F1_FREE, (c - 1) & 1);
              tc_fence_after_sync();
              TL3(2, 1010 + c);
              gemm_k128(tm + 128, xa0, xa1, w1a + 4096 * (c & 3), w1b + 4096 * (c & 3), kIdN32, el);
              mma_commit(sbar + S_F1_DONE, el);
              TL3(2, 1020 + c);
              if ((c & 3) == 3) {
                ring_release(ring_base + (c == 3 ? 2 : 5));
                mma_commit(bars + B_W_EMPTY + 8 * kSlotIn, el);
                ++n_in;
              }
              if (c >= 1) issue_f2(c - 1);
            }
            issue_f2(7);
          }
        }
      }
    }
  } else {
    // ----------------------------------------------------------------------------- compute warps
    // Warps 9..11 only take part in setmaxnreg (whole warpgroups) and the final barrier: the tail is ONE warp on lane quadrant 0.
    // (Rotating the tail over the quadrants to balance the sub-partitions was tried in three forms; the chaos build caught a
    // barrier-phase problem in two of them and a data hazard in all -- the first P.V of the next head overwrites the accumulator
    // lanes the previous owner is still reading -- and the fixed form measured slower than this one: DESIGN.md 4.1c.)
    setmaxnreg_inc<kRegsCompute3>();
    if (warp < 9) {
    const int s = warp >> 2, q = warp & 3;
    const int r = s < 2 ? 128 * s + 32 * q + lane : 256 + lane;     // token row of this thread
    const bool valid = r < kS;
    const uint32_t sbar = bars + B_STREAM + kStreamBars3 * s;
    const uint32_t tl = tmem + ((uint32_t)(q * 32) << 16) + 160 * s;   // this thread's lane, this stream's columns
    const uint32_t vec = sb + OFF_VEC;
    uint32_t n_seq = 0, Lg = 0, hg = 0;
#pragma unroll 1
    for (int seq = blockIdx.x; seq < nseq; seq += gridDim.x, ++n_seq) {
      mbar_wait_spin(bars + B_X_FULL, n_seq & 1);
#pragma unroll 1
      for (int l = 0; l < L; ++l, ++Lg) {
#ifdef AFT_V3_TIMELINE
        const bool tl_on = blockIdx.x == 0 && n_seq == 1 && l == 1 && lane == 0 && (warp == 0 || warp == 4);
        uint32_t tl_n = 0;
        const int tl_slot = warp >> 2;
#endif
#pragma unroll 1
        for (int g = 0; g < 4; ++g, ++hg) {
          // Q/K/V region free: [layer start: every stream's LayerNorm2 of the previous layer], head g - 1 done by every stream
          mbar_wait_spin(bars + B_QKV_FREE, (5 * Lg + g) & 1);
          const uint32_t hv = hg;
          TL3(tl_slot, g * 100 + 1);
          mbar_wait_spin(bars + B_BIAS_FULL + 8 * (g & 1), (hg >> 1) & 1);
          mbar_wait_spin(sbar + S_QKV_DONE, hv & 1);
          tc_fence_after_sync();
          TL3(tl_slot, g * 100 + 2);
          epi_qkv3(tl + 32, sb, miscb + MISC_BIAS + (g & 1) * kBiasBytes3, r);
          tc_fence_before_sync();
          fence_proxy_async_smem();
          warp_arrive(bars + B_QKV_READY, lane);
          TL3(tl_slot, g * 100 + 3);
          float m_ref = -INFINITY, lsum = 0.f;
          // Key loop, software-pipelined by half tiles of 32 keys: the scores of the next half tile travel TMEM -> registers
          // while the exponentials of the current one run (tcgen05.wait::ld waits for every load in flight, so a load is
          // issued right after the wait that completes its predecessor).  xa: first half of a key tile, xb: second half.
          uint32_t xa[32], xb[32];
          mbar_wait_spin(sbar + S_S_DONE, (3 * hv) & 1);
          tc_fence_after_sync();
          TL3(tl_slot, g * 100 + 4);
          tmem_ld_cols(tl, xa);
#pragma unroll 1
          for (int j = 0; j < 4; ++j) {
            const uint32_t sbuf = tl + 64 * (j & 1);
            // P.V(j - 1): only the (rare) rescale of the output accumulator has to wait for it.  P(j) goes to the buffer of S(j), whose
            // last reader P.V(j - 2) completed before S(j) was written (in-order tensor pipe); that completion is consumed here
            // anyway -- it costs one successful try_wait -- because every waiter must walk every phase of a barrier in order
            // (the chaos build showed wrong-parity passes when this warp inferred a completion instead of observing it)
            // (polling form: the arrival lands within nanoseconds of S_DONE(j); parking the warp for it cost 2 ms per 16384 estimates)
            if (j >= 2) mbar_wait_poll(sbar + S_PV_DONE + 8 * (j & 1), ((((j - 2) & 1) ? 2 * hv : 3 * hv) + ((j - 2) >> 1)) & 1);
            const int jm = j - 1;
            const uint32_t pv_bar = sbar + S_PV_DONE + 8 * (jm & 1), pv_par = (((jm & 1) ? 2 * hv : 3 * hv) + (jm >> 1)) & 1;
            bool pv_waited = false;
            tmem_wait_ld();
            TL3(tl_slot, g * 100 + 10 + j);
            tmem_ld_cols(sbuf + 32, xb);
#ifdef AFT_V3_TIMELINE
            unsigned long long dbgv[8];
            softmax_half3<false, false>(xa, sbuf, tl + 128, j == 0, j > 0, pv_bar, pv_par, pv_waited, m_ref, lsum,
                                        (tl_on && g == 1 && j == 1) ? dbgv : nullptr);
            if (tl_on && g == 1 && j == 1)
              for (int e = 0; e < 7; ++e) p.timeline[tl_slot * 256 + (++tl_n)] = ((unsigned long long)(900 + e) << 48) | (dbgv[e] & 0xFFFFFFFFFFFFull);
#else
            softmax_half3<false, false>(xa, sbuf, tl + 128, j == 0, j > 0, pv_bar, pv_par, pv_waited, m_ref, lsum);
#endif
            TL3(tl_slot, g * 100 + 20 + j);
            tmem_wait_ld();
            {   // scores of the next key tile (tile 4: 32 keys) are complete long before: two score buffers
              const int jn = j + 1;
              const uint32_t sidx = (jn & 1) ? 2 * hv + (jn >> 1) : 3 * hv + (jn >> 1);
              mbar_wait_spin(sbar + S_S_DONE + 8 * (jn & 1), sidx & 1);
              tc_fence_after_sync();
              tmem_ld_cols(tl + 64 * (jn & 1), xa);
              TL3(tl_slot, g * 100 + 30 + j);
            }
            softmax_half3<false, true>(xb, sbuf + 16, tl + 128, false, j > 0, pv_bar, pv_par, pv_waited, m_ref, lsum);
            TL3(tl_slot, g * 100 + 40 + j);
            tmem_wait_st();
            tc_fence_before_sync();
            warp_arrive(sbar + S_P_READY + 8 * (j & 1), lane);
            TL3(tl_slot, g * 100 + 50 + j);
          }
          {
            bool pv_waited = false;
            mbar_wait_poll(sbar + S_PV_DONE, (3 * hv + 1) & 1);   // P.V(2), see above
            tmem_wait_ld();
            softmax_half3<true, false>(xa, tl, tl + 128, false, true, sbar + S_PV_DONE + 8, (2 * hv + 1) & 1, pv_waited, m_ref, lsum);
            tmem_wait_st();
            tc_fence_before_sync();
            warp_arrive(sbar + S_P_READY, lane);
            TL3(tl_slot, g * 100 + 61);
          }
          mbar_wait_spin(sbar + S_PV_DONE + 8, (2 * hv + 1) & 1);   // P.V(3)
          mbar_wait_spin(sbar + S_PV_DONE, (3 * hv + 2) & 1);       // P.V(4)
          tc_fence_after_sync();
          TL3(tl_slot, g * 100 + 62);
          epi_o3(tl + 128, sb, g, r, lsum, valid);
          TL3(tl_slot, g * 100 + 63);
          if (g == 3) {
            tc_fence_before_sync();
            fence_proxy_async_smem();
            warp_arrive(sbar + S_O_READY, lane);
          }
        }
        mbar_wait_spin(bars + B_QKV_FREE, (5 * Lg + 4) & 1);   // head 3 done by every stream: every warp walks every phase of the join
        const uint32_t lv = Lg;
        TL3(tl_slot, 1070);
        mbar_wait_spin(sbar + S_OUT_DONE, lv & 1);
        mbar_wait_spin(bars + B_VEC_FULL, Lg & 1);
        tc_fence_after_sync();
        TL3(tl_slot, 1071);
        epi_ln3(tl, sb, vec, 1, r, valid);
        TL3(tl_slot, 1072);
        tc_fence_before_sync();
        fence_proxy_async_smem();
        warp_arrive(sbar + S_X1_READY, lane);
#pragma unroll 1
        for (int c = 0; c < 8; ++c) {
          mbar_wait_spin(sbar + S_F1_DONE, c & 1);
          tc_fence_after_sync();
          TL3(tl_slot, 1080 + c);
          uint32_t a[32];
          tmem_ld_cols(tl + 128, a);
          tmem_wait_ld();
          tc_fence_before_sync();
          warp_arrive(sbar + S_F1_FREE, lane);
          uint32_t pk[16];
          act_chunk3(a, vec, c, p.activation, pk);
          const int ch = c >> 1;
          // hidden buffer ch & 1 still feeds the FFN2 partial products of chunk ch - 2
          if (c >= 4 && (c & 1) == 0) mbar_wait_spin(sbar + S_F2_DONE + 8 * (ch & 1), 0);
          if (valid) {
            const uint32_t row = sb + OFF_O + (ch & 1) * kXChunkBytes + r * 128;
#pragma unroll
            for (int u = 0; u < 4; ++u)
              st_shared_v4(row + ((((c & 1) * 4 + u) ^ (r & 7)) << 4), pk[4 * u], pk[4 * u + 1], pk[4 * u + 2], pk[4 * u + 3]);
          }
          fence_proxy_async_smem();
          warp_arrive(sbar + S_HID_READY + 8 * (c & 1), lane);
          TL3(tl_slot, 1090 + c);
        }
        mbar_wait_spin(sbar + S_F2_DONE, 1);
        mbar_wait_spin(sbar + S_F2_DONE + 8, 1);
        tc_fence_after_sync();
        TL3(tl_slot, 1098);
        epi_ln3(tl, sb, vec, 2, r, valid);
        TL3(tl_slot, 1099);
        tc_fence_before_sync();
        fence_proxy_async_smem();
        warp_arrive(sbar + S_X2_READY, lane);
      }
      warp_arrive(bars + B_X_DONE, lane);
    }
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == kMmaWarp0) tmem_dealloc(tmem, 512);
}

}  // namespace

bool tc_encoder3_launch(char* x_images, const TcLayer* layers_dev, int num_layers, int activation, int64_t nseq, int sm_count,
                        cudaStream_t st) {
  if (cudaFuncSetAttribute(encoder3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem3) != cudaSuccess) {
    set_error("encoder3_kernel: cannot opt in to %u bytes of shared memory: %s", kSmem3, cudaGetErrorString(cudaGetLastError()));
    return false;
  }
  Enc3Params ep;
  ep.x_images = x_images;
  ep.layers = layers_dev;
  ep.num_layers = num_layers;
  ep.activation = activation;
  ep.nseq = nseq;
  ep.timeline = nullptr;
  const unsigned grid = (unsigned)(nseq < sm_count ? nseq : sm_count);
#ifdef AFT_V3_TIMELINE
  static unsigned long long* tl_dev = nullptr;
  if (!tl_dev) cudaMalloc(&tl_dev, 3 * 256 * 8);
  cudaMemsetAsync(tl_dev, 0, 3 * 256 * 8, st);
  ep.timeline = tl_dev;
#endif
  encoder3_kernel<<<grid, kThreads3, kSmem3, st>>>(ep);
  count_launch();
#ifdef AFT_V3_TIMELINE
  if (getenv("AFT_V3_TIMELINE_DUMP")) {
    static unsigned long long host[3 * 256];
    cudaStreamSynchronize(st);
    cudaMemcpy(host, tl_dev, sizeof(host), cudaMemcpyDeviceToHost);
    for (int sl = 0; sl < 3; ++sl)
      for (int i = 1; i < 256 && host[sl * 256 + i]; ++i)
        fprintf(stderr, "TL3 %d %llu %llu\n", sl, host[sl * 256 + i] >> 48, host[sl * 256 + i] & 0xFFFFFFFFFFFFull);
  }
#endif
  return check_launch("encoder3_kernel");
}

}  // namespace aft
