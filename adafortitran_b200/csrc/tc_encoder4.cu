// AFT_BF16 path, encoder kernel v4: the 6-layer post-norm transformer encoder (reference src/models/blocks/encoders.py:44-55,69
// -> torch _transformer_encoder_layer_fwd) as one persistent kernel: THREE ROW-TILE STREAMS per CTA, TWO WARP SETS per stream.
//
// What v3 (tc_encoder3.cu) measured: three asynchronous row-tile streams with one thread per token row do not slow each
// other down -- the layer time IS the length of one stream's serial chain (in-kernel timeline: ~1.9 k clk per 64-key
// tile, ~13 k clk per head, 80 k clk per layer), because a single warp issues its dependent epilogue code at ~3-4 clk
// per instruction.  v4 halves that chain: every stream is worked by two warp sets that split each token row's work,
//   attention : key tiles of 32 keys alternate between the sets (set 0: tiles 0,2,4,6,8; set 1: tiles 1,3,5,7); each set
//               keeps its own running maximum / sum and its own output accumulator (the two partial softmaxes are merged
//               once per head),
//   QKV / O   : the sets split the columns,
//   LayerNorm : 64 columns each, one exchange of (sum, sum of squares),
//   FFN       : the 32-unit hidden chunks alternate between the sets.
//
//   stream 0 : token rows   0..127   warps 0..3 (set 0), 4..7 (set 1)    MMA issuer warp 21
//   stream 1 : token rows 128..255   warps 8..11 (set 0), 12..15 (set 1) MMA issuer warp 22
//   stream 2 : token rows 256..279   warps 16..19                        MMA issuer warp 23
//              the 24-row tail: TWO of its four warps are active at a time, as set 0 and set 1 on two different TMEM lane
//              quadrants (the A operand of the tail's M = 128 MMAs starts 32 q rows early, which puts the tail rows on
//              the lanes of quadrant q); the pair of quadrants rotates with the head / the layer.  Each set of the tail
//              gets its own MMAs (its own shift), so the per-thread code is the one of the main streams; only the
//              exchanges between the two sets go through shared memory instead of TMEM.
//   warp 20  : producer (bulk copies of weights / the sequence image, result image back to global memory)
//
// Tensor memory: stream s owns columns [160 s, 160 s + 160):  O_0 [0,32) O_1 [32,64) S_0 [64,96) P_0 [96,112) S_1 [112,144)
// P_1 [144,160)  |  QKV accumulators [64,160)  |  out_proj / FFN2 accumulator [0,128), FFN1 chunk [128,160).
// Columns [480,512): (m, l) / LayerNorm partial sums exchanged between the two sets of a main stream (same lanes).
// Shared memory map: as tc_encoder.cu (O | X | QKV / weight ring | W | MISC).
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

#ifndef AFT_V4_POLY_MASK
#define AFT_V4_POLY_MASK 0x88   // bit jj set: pair jj of every 8 pairs of exponentials runs on the FMA pipe (packed Cody-Waite + cubic)
#endif
#ifndef AFT_V4_REGS_COMPUTE
#define AFT_V4_REGS_COMPUTE 88
#endif
#ifndef AFT_V4_REGS_CTRL
#define AFT_V4_REGS_CTRL 40
#endif

constexpr int kThreads4 = 768;
constexpr int kProducerWarp4 = 20, kMmaWarp4 = 21;
constexpr float kRescale4 = 8.0f;

constexpr int kVecBlock4 = 384;
constexpr int kVBOut = 0, kVBL1 = 128, kVBL2 = 384, kVN1W = 512, kVN1B = 640, kVN2W = 768, kVN2B = 896;
constexpr uint32_t kBiasBytes4 = 96 * 4, kVecBytes4 = 1024 * 4;

constexpr uint32_t OFF_O = 0, OFF_X = 73728, OFF_QKV = 147456, OFF_W = 202752, OFF_MISC = 227328;
constexpr uint32_t kQkvPart = 18432, kSlot = 16384, kWInSlice = 24576;
constexpr uint32_t OFF_VEC = OFF_QKV + 3 * kSlot;
constexpr uint32_t kSmem4 = OFF_MISC + 5120;   // 232,448
// MISC: in_proj bias double buffer | mbarriers | TMEM base | tail exchange of partial output accumulators [2 sets][24 rows][16 floats].
// The tail's (m, l) / LayerNorm-sum exchange [2 uses][2 sets][32 rows] float2 lives in the padding rows 280..287 of the O
// image's first chunk (1 KB that no epilogue writes; the tensor core only reads them into padding rows of its results).
constexpr uint32_t MISC_BIAS = 0, MISC_BARS = 768, MISC_TMEM = 1456, MISC_TX_O = 1472;
constexpr uint32_t OFF_TX_ML = OFF_O + 280 * 128;
static_assert(MISC_TX_O + 2 * 24 * 64 <= 5120 && 2 * 2 * 32 * 8 <= 8 * 128, "exchange areas overflow");

// mbarriers.  Protocol rule (tc_ptx.cuh / DESIGN.md): a waiter tests phase parity, so completion k + 1 of a barrier must
// causally depend on every waiter having passed its wait for completion k - 1; every waiter waits for every completion
// in order (or a preceding join implies the completions it skips).
enum : uint32_t {
  B_X_FULL = 0,        // commit : sequence image landed
  B_X_DONE = 8,        // warps(20): last LayerNorm of the sequence written
  B_ATTN_DONE = 16,    // 3 commits: all P.V of the layer complete (producer: ring / vector block may overwrite Q/K/V)
  B_QKV_READY = 24,    // warps(18): Q/K/V rows of head g written by every stream
  B_QKV_FREE = 32,     // 3 arrivals: Q/K/V region free -- 5 completions per layer: [layer start], head 0..3 done
  B_VEC_FULL = 40,     // commit
  B_BIAS_FULL = 48,    // 2 x commit
  B_W_FULL = 64,       // 4 x commit : [0..2] ring slots, [3] in_proj slot
  B_W_EMPTY = 96,      // 4 x 3 commits
  B_STREAM = 128,      // per-stream blocks of 176 bytes
};
constexpr uint32_t kStreamBars = 176;
enum : uint32_t {
  S_QKV_DONE = 0,       // commit
  S_S_DONE = 8,         // 2 x commit (per set): score tile complete
  S_S_LOADED = 24,      // 2 x warps  : score tile in registers (the set's score buffer may take the next tile)
  S_P_READY = 40,       // 2 x warps  : P tile stored
  S_PV_DONE = 56,       // 2 x commit : P.V of the set's tile complete (P buffer free, accumulator valid)
  S_O_READY = 72,       // warps(8)   : O image rows of the layer complete
  S_OUT_DONE = 80,      // commit
  S_X1_READY = 88,      // warps      : LayerNorm1 rows written
  S_F1_DONE = 96,       // 2 x commit (chunk parity)
  S_F1_FREE = 112,      // 2 x warps
  S_HID_READY = 128,    // 2 x warps
  S_F2_DONE = 144,      // 2 x commit (hidden buffer)
  S_X2_READY = 160,     // warps
};
static_assert(MISC_BARS + B_STREAM + 3 * kStreamBars <= MISC_TMEM, "barrier block overflow");
constexpr uint32_t kSlotIn = 3;

// TMEM columns inside a stream's 160
constexpr uint32_t T_O = 0, T_S0 = 64, T_P0 = 96, T_S1 = 112, T_P1 = 144, T_QKV = 64, T_ACC = 0, T_F1 = 128, T_XC = 480;
__device__ __forceinline__ uint32_t t_s(int x) { return x ? T_S1 : T_S0; }
__device__ __forceinline__ uint32_t t_p(int x) { return x ? T_P1 : T_P0; }

constexpr uint32_t kIdQkv96 = make_idesc_bf16(128, 96, false, false), kIdQkv48 = make_idesc_bf16(128, 48, false, false);
constexpr uint32_t kIdS32 = make_idesc_bf16(128, 32, false, false);
constexpr uint32_t kIdPV = make_idesc_bf16(128, 32, false, true);
constexpr uint32_t kIdN128 = make_idesc_bf16(128, 128, false, false), kIdN64 = make_idesc_bf16(128, 64, false, false),
                   kIdN32 = make_idesc_bf16(128, 32, false, false);

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
__device__ __forceinline__ void tmem_st2(uint32_t taddr, uint32_t a, uint32_t b) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(taddr), "r"(a), "r"(b));
}
__device__ __forceinline__ void tmem_ld2(uint32_t taddr, uint32_t& a, uint32_t& b) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(a), "=r"(b) : "r"(taddr));
}
__device__ __forceinline__ void st_shared_f2(uint32_t addr, float a, float b) {
  asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(a), "f"(b));
}
__device__ __forceinline__ void ld_shared_f2(uint32_t addr, float& a, float& b) {
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(a), "=f"(b) : "r"(addr));
}

struct Enc4Params {
  char* x_images;
  const TcLayer* layers;
  int num_layers;
  int activation;
  int64_t nseq;
};

// what a compute thread knows about itself
struct Me {
  uint32_t sb, miscb;
  uint32_t tl;       // TMEM address: this thread's lane, this stream's first column
  uint32_t txc;      // TMEM address: this thread's lane, exchange columns of this main stream (2 uses x 2 sets x 2 columns)
  int s, x, r, lane;
  bool valid;        // r < kS
  int pair_bar;      // named barrier shared with the partner warp (the other set's warp that owns the same rows)
};

// (a, b) -> partner, partner's (a, b) <- ; `use` alternates the exchange slots (a slot is rewritten two exchanges later, behind
// the barrier of the exchange in between, which the partner only reaches after it has read this one)
__device__ __forceinline__ void exchange2(const Me& me, uint32_t use, float a, float b, float& pa, float& pb) {
  if (me.s < 2) {
    tmem_st2(me.txc + (use & 1) * 4 + me.x * 2, __float_as_uint(a), __float_as_uint(b));
    tmem_wait_st();
    tc_fence_before_sync();
    named_bar_sync(me.pair_bar, 64);
    tc_fence_after_sync();
    uint32_t ua, ub;
    tmem_ld2(me.txc + (use & 1) * 4 + (me.x ^ 1) * 2, ua, ub);
    tmem_wait_ld();
    pa = __uint_as_float(ua);
    pb = __uint_as_float(ub);
  } else {
    const uint32_t base = me.sb + OFF_TX_ML + (use & 1) * 512;
    st_shared_f2(base + (me.x * 32 + me.lane) * 8, a, b);
    named_bar_sync(me.pair_bar, 64);
    ld_shared_f2(base + ((me.x ^ 1) * 32 + me.lane) * 8, pa, pb);
  }
}

// ---------------------------------------------------------------------------------------------
// compute-warp epilogues: the thread owns token row r together with its partner of the other set
// ---------------------------------------------------------------------------------------------
// QKV accumulators [q_g | k_g | v_g]: this set's 48 columns (6 units of 8) + bias -> bf16 -> row r of the Q / K / V images
__device__ __forceinline__ void epi_qkv4(const Me& me, uint32_t bias96) {
  uint32_t acc[48];
  tmem_ld_cols(me.tl + T_QKV + 48 * me.x, acc);
  tmem_wait_ld();
  const int sw = (me.r >> 1) & 3;
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    const int u8 = 6 * me.x + i, mat = u8 >> 2, u = u8 & 3;
    const float4 b0 = lds_f4(bias96 + u8 * 32), b1 = lds_f4(bias96 + u8 * 32 + 16);
    const uint32_t* a = acc + i * 8;
    auto sum = [](uint32_t x0, uint32_t x1, float y0, float y1) {
      return pack_bf16_pair(add2(pack2(__uint_as_float(x0), __uint_as_float(x1)), pack2(y0, y1)));
    };
    st_shared_v4(me.sb + OFF_QKV + mat * kQkvPart + me.r * 64 + ((u ^ sw) << 4), sum(a[0], a[1], b0.x, b0.y), sum(a[2], a[3], b0.z, b0.w),
                 sum(a[4], a[5], b1.x, b1.y), sum(a[6], a[7], b1.z, b1.w));
  }
}

// One key tile (32 keys) of this set's streaming softmax: scores x -> running maximum -> exponentials against the reference
// maximum -> bf16 pairs pk (the caller stores them) -> row sum.  The set's output accumulator is rescaled only when some
// row's maximum outgrew its reference by 2^8 (rare): P.V of the set's previous tile must have completed then.
template <bool kMask>
__device__ __forceinline__ void softmax_tile4(const uint32_t (&x)[32], uint32_t (&pk)[16], uint32_t o_addr, bool first, uint32_t pv_bar,
                                              uint32_t pv_par, bool& pv_waited, float& m_ref, float& lsum) {
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
  if (__any_sync(0xFFFFFFFFu, mt > m_ref + kRescale4)) {
    const float mn = fmaxf(m_ref, mt);
    const float f = ex2(m_ref - mn);   // first tile: exp2(-inf) = 0
    if (!first) {
      if (!pv_waited) {
        mbar_wait(pv_bar, pv_par);
        tc_fence_after_sync();
        pv_waited = true;
      }
      uint32_t a[32];
      tmem_ld_cols(o_addr, a);
      tmem_wait_ld();
#pragma unroll
      for (int i = 0; i < 32; ++i) a[i] = __float_as_uint(__uint_as_float(a[i]) * f);
#pragma unroll
      for (int i = 0; i < 4; ++i) tmem_st8p(o_addr + i * 8, a + 8 * i);
    }
    lsum *= f;
    m_ref = mn;
  }
  const f32x2 negm2 = pack2(-m_ref, -m_ref);
  f32x2 s2a = pack2(0.f, 0.f), s2b = pack2(0.f, 0.f);
#pragma unroll
  for (int jj = 0; jj < 16; ++jj) {
    const f32x2 x2 = add2(pack2(v[2 * jj], v[2 * jj + 1]), negm2);
    f32x2 e2;
    if (((AFT_V4_POLY_MASK) >> (jj & 7)) & 1) {
      e2 = ex2_poly2(x2);
    } else {
      float a, b;
      unpack2(x2, a, b);
      e2 = pack2(ex2(a), ex2(b));
    }
    if (jj & 1) s2b = add2(s2b, e2); else s2a = add2(s2a, e2);
    pk[jj] = pack_bf16_pair(e2);
  }
  float sa, sb2, sc, sd;
  unpack2(s2a, sa, sb2);
  unpack2(s2b, sc, sd);
  lsum += (sa + sb2) + (sc + sd);
}

// merge of the two sets' partial softmaxes of head g: this set writes output dims 16 x .. 16 x + 15 of O image row r
__device__ __forceinline__ void epi_merge4(const Me& me, int g, uint32_t use, float m_ref, float lsum) {
  const int d0 = 16 * me.x;
  uint32_t own[16], oth[16];
  float pm, pl;
  if (me.s < 2) {
    exchange2(me, use, m_ref, lsum, pm, pl);
    tmem_ld16(me.tl + T_O + 32 * me.x + d0, own);
    tmem_ld16(me.tl + T_O + 32 * (me.x ^ 1) + d0, oth);
    tmem_wait_ld();
  } else {
    // the partner sits on another lane quadrant: its accumulator columns travel through shared memory
    uint32_t send[16];
    tmem_ld16(me.tl + T_O + 32 * me.x + 16 * (me.x ^ 1), send);   // the 16 dims the partner merges
    tmem_ld16(me.tl + T_O + 32 * me.x + d0, own);
    tmem_wait_ld();
    if (me.lane < 24) {
      const uint32_t dst = me.miscb + MISC_TX_O + (me.x * 24 + me.lane) * 64;
#pragma unroll
      for (int u = 0; u < 4; ++u) st_shared_v4(dst + u * 16, send[4 * u], send[4 * u + 1], send[4 * u + 2], send[4 * u + 3]);
    }
    exchange2(me, use, m_ref, lsum, pm, pl);   // its barrier also publishes the accumulator columns
    const uint32_t src = me.miscb + MISC_TX_O + ((me.x ^ 1) * 24 + (me.lane < 24 ? me.lane : 0)) * 64;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const uint4 t = ld_shared_v4(src + u * 16);
      oth[4 * u] = t.x; oth[4 * u + 1] = t.y; oth[4 * u + 2] = t.z; oth[4 * u + 3] = t.w;
    }
    named_bar_sync(me.pair_bar, 64);   // both have read: the area may be rewritten by the next head
  }
  const float m = fmaxf(m_ref, pm);
  const float ws = ex2(m_ref - m), wo = ex2(pm - m);
  const float inv = rcp_approx(fmaf(lsum, ws, pl * wo));
  const f32x2 a2 = pack2(ws * inv, ws * inv), b2 = pack2(wo * inv, wo * inv);
  if (me.valid) {
    const uint32_t row = me.sb + OFF_O + (g >> 1) * kXChunkBytes + me.r * 128;
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      auto mix = [&](int j) {
        const f32x2 o = pack2(__uint_as_float(own[8 * u + j]), __uint_as_float(own[8 * u + j + 1]));
        const f32x2 p = pack2(__uint_as_float(oth[8 * u + j]), __uint_as_float(oth[8 * u + j + 1]));
        return pack_bf16_pair(fma2(o, a2, mul2(p, b2)));
      };
      st_shared_v4(row + ((((g & 1) * 4 + 2 * me.x + u) ^ (me.r & 7)) << 4), mix(0), mix(2), mix(4), mix(6));
    }
  }
}

// accumulator columns 64 x .. 64 x + 63 + bias + residual (X row r, K-chunk x) -> LayerNorm over the row (partial sums
// exchanged with the partner) -> X row r in place.  The pre-norm values go back to TMEM between the two passes.
__device__ __forceinline__ void epi_ln4(const Me& me, uint32_t vec, int which, uint32_t use) {
  const int c0 = 64 * me.x;
  const uint32_t acc_addr = me.tl + T_ACC + c0;
  const uint32_t bias = vec + 4 * ((which == 1 ? kVBOut : kVBL2) + c0);
  const uint32_t gam = vec + 4 * ((which == 1 ? kVN1W : kVN2W) + c0);
  const uint32_t bet = vec + 4 * ((which == 1 ? kVN1B : kVN2B) + c0);
  const uint32_t xrow = me.sb + OFF_X + me.x * kXChunkBytes + me.r * 128;
  f32x2 s2 = pack2(0.f, 0.f), q2 = pack2(0.f, 0.f);
#pragma unroll
  for (int blk = 0; blk < 2; ++blk) {
    uint32_t acc[32];
    tmem_ld_cols(acc_addr + blk * 32, acc);
    uint4 xr[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) xr[u] = ld_shared_v4(xrow + (((blk * 4 + u) ^ (me.r & 7)) << 4));
    tmem_wait_ld();
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
  float sa, sb2, qa, qb, ps, pq;
  unpack2(s2, sa, sb2);
  unpack2(q2, qa, qb);
  exchange2(me, use, sa + sb2, qa + qb, ps, pq);   // (main streams: its tcgen05.wait::st also completes the stores above)
  if (me.s == 2) tmem_wait_st();
  const float mean = ((sa + sb2) + ps) * (1.0f / 128.0f);
  const float var = fmaxf(fmaf(-mean, mean, ((qa + qb) + pq) * (1.0f / 128.0f)), 0.f);
  const float rstd = rsqrtf(var + 1e-5f);
  const f32x2 rstd2 = pack2(rstd, rstd), shift2 = pack2(-mean * rstd, -mean * rstd);
#pragma unroll
  for (int blk = 0; blk < 2; ++blk) {
    uint32_t y[32];
    tmem_ld_cols(acc_addr + blk * 32, y);
    tmem_wait_ld();
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
      if (me.valid) st_shared_v4(xrow + (((blk * 4 + u) ^ (me.r & 7)) << 4), pk[0], pk[1], pk[2], pk[3]);
    }
  }
}

// FFN1 chunk c (32 hidden units) + bias -> GELU / ReLU -> bf16 pairs
__device__ __forceinline__ void act_chunk4(const uint32_t (&a)[32], uint32_t vec, int c, int act, uint32_t (&pk)[16]) {
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
__global__ void __launch_bounds__(kThreads4, 1) encoder4_kernel(Enc4Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t sb = smem_u32(smem_raw);
  if ((sb & 1023u) != 0) __trap();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t miscb = sb + OFF_MISC, bars = miscb + MISC_BARS;

  if (threadIdx.x == 0) {
    mbar_init(bars + B_X_FULL, 1);
    mbar_init(bars + B_X_DONE, 20);
    mbar_init(bars + B_ATTN_DONE, 3);
    mbar_init(bars + B_QKV_READY, 18);
    mbar_init(bars + B_QKV_FREE, 3);
    mbar_init(bars + B_VEC_FULL, 1);
    mbar_init(bars + B_BIAS_FULL, 1);
    mbar_init(bars + B_BIAS_FULL + 8, 1);
    for (int i = 0; i < 4; ++i) { mbar_init(bars + B_W_FULL + 8 * i, 1); mbar_init(bars + B_W_EMPTY + 8 * i, 3); }
    for (int s = 0; s < 3; ++s) {
      const uint32_t b = bars + B_STREAM + kStreamBars * s, nset = s < 2 ? 4 : 1;
      mbar_init(b + S_QKV_DONE, 1);
      for (int x = 0; x < 2; ++x) {
        mbar_init(b + S_S_DONE + 8 * x, 1);
        mbar_init(b + S_S_LOADED + 8 * x, nset);
        mbar_init(b + S_P_READY + 8 * x, nset);
        mbar_init(b + S_PV_DONE + 8 * x, 1);
        mbar_init(b + S_F1_DONE + 8 * x, 1);
        mbar_init(b + S_F1_FREE + 8 * x, nset);
        mbar_init(b + S_HID_READY + 8 * x, nset);
        mbar_init(b + S_F2_DONE + 8 * x, 1);
      }
      mbar_init(b + S_O_READY, 8);
      mbar_init(b + S_OUT_DONE, 1);
      mbar_init(b + S_X1_READY, 2 * nset);
      mbar_init(b + S_X2_READY, 2 * nset);
    }
    fence_mbar_init();
  }
  if (warp == kMmaWarp4) { tmem_alloc(miscb + MISC_TMEM, 512); tmem_relinquish(); }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  uint32_t tmem;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(miscb + MISC_TMEM));

  const int L = p.num_layers;
  const int nseq = (int)p.nseq;

  if (warp >= 20) {
    setmaxnreg_dec<AFT_V4_REGS_CTRL>();
    if (warp == kProducerWarp4) {
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
              mbar_arrive_expect_tx(bars + B_BIAS_FULL + 8 * (g & 1), kBiasBytes4);
              bulk_g2s(miscb + MISC_BIAS + (g & 1) * kBiasBytes4, W.b_in + g * 96, kBiasBytes4, bars + B_BIAS_FULL + 8 * (g & 1));
            }
            mbar_wait_relaxed(bars + B_ATTN_DONE, n_attn & 1);   // Q/K/V images dead
            ++n_attn;
            mbar_arrive_expect_tx(bars + B_VEC_FULL, kVecBytes4);
            bulk_g2s(sb + OFF_VEC, W.b_in + kVecBlock4, kVecBytes4, bars + B_VEC_FULL);
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
      const int s = warp - kMmaWarp4;
      const bool el = elect_one();
      const uint32_t sbar = bars + B_STREAM + kStreamBars * s;
      const uint32_t tm = tmem + 160 * s;
      uint32_t n_seq = 0, Lg = 0, hg = 0, n_in = 0, ring_base = 0;
      auto ring_wait = [&](uint32_t idx) -> uint32_t {
        mbar_wait(bars + B_W_FULL + 8 * (idx % 3), (idx / 3) & 1);
        tc_fence_after_sync();
        return sb + OFF_QKV + (idx % 3) * kSlot;
      };
      auto ring_release = [&](uint32_t idx) { mma_commit(bars + B_W_EMPTY + 8 * (idx % 3), el); };
      for (int seq = blockIdx.x; seq < nseq; seq += gridDim.x, ++n_seq) {
        mbar_wait(bars + B_X_FULL, n_seq & 1);
        for (int l = 0; l < L; ++l, ++Lg, ring_base += 8) {
          // layer start: the stream's LayerNorm2 of the previous layer is written (X rows, accumulator columns and the
          // vector block are free as far as this stream is concerned)
          if (Lg > 0) mbar_wait(sbar + S_X2_READY, (Lg - 1) & 1);
          tc_fence_after_sync();
          if (lane == 0) mbar_arrive(bars + B_QKV_FREE);
          __syncwarp();
          // first token row of the A operand for set x: main streams: the row tile; tail: 32 q rows early (quadrant q = (k + 2 x) & 3)
          auto arow_of = [&](int k, int x) -> int { return s < 2 ? 128 * s : 256 - 32 * ((k + 2 * x) & 3); };
          auto issue_qkv = [&](int g) {
            mbar_wait(bars + B_W_FULL + 8 * kSlotIn, n_in & 1);
            tc_fence_after_sync();
            if (s < 2) {
              const uint32_t a0 = sb + OFF_X + arow_of(g, 0) * 128;
              gemm_k128(tm + T_QKV, a0, a0 + kXChunkBytes, sb + OFF_W, sb + OFF_W + 96 * 128, kIdQkv96, el);
            } else {
#pragma unroll
              for (int x = 0; x < 2; ++x) {   // each set of the tail: its 48 columns on its own lane quadrant
                const uint32_t a0 = sb + OFF_X + arow_of(g, x) * 128;
                gemm_k128(tm + T_QKV + 48 * x, a0, a0 + kXChunkBytes, sb + OFF_W + 48 * 128 * x, sb + OFF_W + 96 * 128 + 48 * 128 * x, kIdQkv48, el);
              }
            }
            mma_commit(bars + B_W_EMPTY + 8 * kSlotIn, el);
            mma_commit(sbar + S_QKV_DONE, el);
            ++n_in;
          };
          issue_qkv(0);
          for (int g = 0; g < 4; ++g, ++hg) {
            mbar_wait(bars + B_QKV_READY, hg & 1);
            tc_fence_after_sync();
            const uint32_t kd = lo_k64(sb + OFF_QKV + kQkvPart);
            const uint32_t vd = lo_mn64(sb + OFF_QKV + 2 * kQkvPart);
            // tile i of set x = key tile T = 2 i + x (32 keys)
            auto issue_s = [&](int x, int i) {
              const int T = 2 * i + x;
              const uint32_t qd = lo_k64(sb + OFF_QKV + arow_of(g, x) * 64);
#pragma unroll
              for (int ks = 0; ks < 2; ++ks) mma_ss(tm + t_s(x), d64(qd + ks * 2), d64(kd + T * 128 + ks * 2), kIdS32, ks > 0, el);
              mma_commit(sbar + S_S_DONE + 8 * x, el);
            };
            auto issue_pv = [&](int x, int i) {
              const int T = 2 * i + x;
#pragma unroll
              for (int ks = 0; ks < 2; ++ks)
                mma_ts(tm + T_O + 32 * x, tm + t_p(x) + ks * 8, d64(vd + T * 128 + ks * 64), kIdPV, i > 0 || ks > 0, el);
              mma_commit(sbar + S_PV_DONE + 8 * x, el);
            };
            issue_s(0, 0);
            issue_s(1, 0);
#pragma unroll 1
            for (int i = 0; i < 5; ++i) {
              mbar_wait(sbar + S_S_LOADED, (5 * hg + i) & 1);
              tc_fence_after_sync();
              if (i + 1 < 5) issue_s(0, i + 1);
              if (i < 4) {
                mbar_wait(sbar + S_S_LOADED + 8, (4 * hg + i) & 1);
                tc_fence_after_sync();
                if (i + 1 < 4) issue_s(1, i + 1);
              }
              mbar_wait(sbar + S_P_READY, (5 * hg + i) & 1);
              tc_fence_after_sync();
              issue_pv(0, i);
              if (i < 4) {
                mbar_wait(sbar + S_P_READY + 8, (4 * hg + i) & 1);
                tc_fence_after_sync();
                issue_pv(1, i);
              }
            }
            // every P.V of this head by this stream has been issued: arrival when they complete
            mma_commit(bars + B_QKV_FREE, el);
            if (g == 3) mma_commit(bars + B_ATTN_DONE, el);
            if (g < 3) {
              // the next head's projection overwrites the score / P columns: every MMA of this head must have completed
              mbar_wait(sbar + S_PV_DONE, (5 * hg + 4) & 1);
              mbar_wait(sbar + S_PV_DONE + 8, (4 * hg + 3) & 1);
              tc_fence_after_sync();
              issue_qkv(g + 1);
            }
          }
          // ---- linear part of the layer on this stream's row tile
          mbar_wait(sbar + S_O_READY, Lg & 1);
          tc_fence_after_sync();
          {
            const uint32_t w0 = ring_wait(ring_base + 0), w1 = ring_wait(ring_base + 1);
            if (s < 2) {
              const uint32_t a0 = sb + OFF_O + arow_of(l, 0) * 128;
              gemm_k128(tm + T_ACC, a0, a0 + kXChunkBytes, w0, w1, kIdN128, el);
            } else {
#pragma unroll
              for (int x = 0; x < 2; ++x) {
                const uint32_t a0 = sb + OFF_O + arow_of(l, x) * 128;
                gemm_k128(tm + T_ACC + 64 * x, a0, a0 + kXChunkBytes, w0 + 64 * 128 * x, w1 + 64 * 128 * x, kIdN64, el);
              }
            }
            mma_commit(sbar + S_OUT_DONE, el);
            ring_release(ring_base + 0);
            ring_release(ring_base + 1);
          }
          mbar_wait(sbar + S_X1_READY, Lg & 1);
          tc_fence_after_sync();
          {
            uint32_t w1a = 0, w1b = 0, w2 = 0;
            auto issue_f2 = [&](int c) {   // FFN2 partial product over hidden units 32 c .. 32 c + 31
              mbar_wait(sbar + S_HID_READY + 8 * (c & 1), (c >> 1) & 1);
              tc_fence_after_sync();
              const int ch = c >> 1;   // 64-unit chunk = W2 K-chunk = hidden buffer ch & 1
              if ((c & 1) == 0) w2 = ring_wait(ring_base + (ch < 2 ? 3 + ch : 4 + ch));
              if (s < 2) {
                const uint32_t ha = sb + OFF_O + (ch & 1) * kXChunkBytes + arow_of(l, 0) * 128;
#pragma unroll
                for (int i = 0; i < 2; ++i) mma_ss(tm + T_ACC, d128(ha, 2 * (c & 1) + i), d128(w2, 2 * (c & 1) + i), kIdN128, c > 0 || i > 0, el);
              } else {
#pragma unroll
                for (int x = 0; x < 2; ++x) {
                  const uint32_t ha = sb + OFF_O + (ch & 1) * kXChunkBytes + arow_of(l, x) * 128;
#pragma unroll
                  for (int i = 0; i < 2; ++i)
                    mma_ss(tm + T_ACC + 64 * x, d128(ha, 2 * (c & 1) + i), d128(w2 + 64 * 128 * x, 2 * (c & 1) + i), kIdN64, c > 0 || i > 0, el);
                }
              }
              if (c & 1) {
                mma_commit(sbar + S_F2_DONE + 8 * (ch & 1), el);
                ring_release(ring_base + (ch < 2 ? 3 + ch : 4 + ch));
              }
            };
#pragma unroll 1
            for (int c = 0; c < 8; ++c) {
              if ((c & 3) == 0) {
                w1a = ring_wait(ring_base + (c == 0 ? 2 : 5));
                mbar_wait(bars + B_W_FULL + 8 * kSlotIn, n_in & 1);
                tc_fence_after_sync();
                w1b = sb + OFF_W;
              }
              // the single FFN1 buffer: chunk c - 1 (the other set's) must be in registers
              if (c > 0) mbar_wait(sbar + S_F1_FREE + 8 * ((c - 1) & 1), ((c - 1) >> 1) & 1);
              tc_fence_after_sync();
              const uint32_t xa0 = sb + OFF_X + arow_of(l, c & 1) * 128;   // the chunk belongs to set c & 1
              gemm_k128(tm + T_F1, xa0, xa0 + kXChunkBytes, w1a + 4096 * (c & 3), w1b + 4096 * (c & 3), kIdN32, el);
              mma_commit(sbar + S_F1_DONE + 8 * (c & 1), el);
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
    setmaxnreg_inc<AFT_V4_REGS_COMPUTE>();
    Me me;
    me.sb = sb;
    me.miscb = miscb;
    me.lane = lane;
    const int q = warp & 3;
    me.s = warp < 16 ? warp >> 3 : 2;
    me.x = warp < 16 ? (warp >> 2) & 1 : 0;   // tail: set per head / layer below
    me.r = me.s < 2 ? 128 * me.s + 32 * q + lane : 256 + lane;
    me.valid = me.r < kS;
    me.tl = tmem + ((uint32_t)(q * 32) << 16) + 160 * me.s;
    me.txc = tmem + ((uint32_t)(q * 32) << 16) + T_XC + 8 * me.s;
    me.pair_bar = me.s < 2 ? 1 + 4 * me.s + q : 9 + (q & 1);   // tail pairs are quadrants {q, q + 2}
    const int s = me.s;
    const uint32_t sbar = bars + B_STREAM + kStreamBars * s;
    const uint32_t vec = sb + OFF_VEC;
    uint32_t n_seq = 0, Lg = 0, hg = 0;   // exchanges with the partner are numbered 6 Lg + {head, 4, 5}: the slot parity every warp agrees on
#pragma unroll 1
    for (int seq = blockIdx.x; seq < nseq; seq += gridDim.x, ++n_seq) {
      mbar_wait(bars + B_X_FULL, n_seq & 1);
#pragma unroll 1
      for (int l = 0; l < L; ++l, ++Lg) {
#pragma unroll 1
        for (int g = 0; g < 4; ++g, ++hg) {
          // Q/K/V region free: [layer start: every stream's LayerNorm2 of the previous layer], head g - 1 done by every stream
          mbar_wait(bars + B_QKV_FREE, (5 * Lg + g) & 1);
          if (s == 2) {   // tail stream: quadrant g & 3 is set 0 of this head, quadrant (g + 2) & 3 is set 1
            if (q == (g & 3)) me.x = 0; else if (q == ((g + 2) & 3)) me.x = 1; else continue;
          }
          const int x = me.x;
          const uint32_t nt = x ? 4 : 5;   // key tiles of this set
          mbar_wait(bars + B_BIAS_FULL + 8 * (g & 1), (hg >> 1) & 1);
          mbar_wait(sbar + S_QKV_DONE, hg & 1);
          tc_fence_after_sync();
          epi_qkv4(me, miscb + MISC_BIAS + (g & 1) * kBiasBytes4);
          tc_fence_before_sync();
          fence_proxy_async_smem();
          warp_arrive(bars + B_QKV_READY, lane);
          float m_ref = -INFINITY, lsum = 0.f;
          const uint32_t s_addr = me.tl + t_s(x), p_addr = me.tl + t_p(x), o_addr = me.tl + T_O + 32 * x;
          const uint32_t b_sdone = sbar + S_S_DONE + 8 * x, b_sload = sbar + S_S_LOADED + 8 * x, b_pready = sbar + S_P_READY + 8 * x,
                         b_pvdone = sbar + S_PV_DONE + 8 * x;
#pragma unroll 1
          for (uint32_t i = 0; i < nt; ++i) {
            const uint32_t idx = nt * hg + i;
            mbar_wait(b_sdone, idx & 1);
            tc_fence_after_sync();
            uint32_t sc[32], pk[16];
            tmem_ld_cols(s_addr, sc);
            tmem_wait_ld();
            tc_fence_before_sync();
            warp_arrive(b_sload, lane);
            bool pv_waited = i == 0;
            if (x == 0 && i == 4) softmax_tile4<true>(sc, pk, o_addr, false, b_pvdone, (idx - 1) & 1, pv_waited, m_ref, lsum);
            else softmax_tile4<false>(sc, pk, o_addr, i == 0, b_pvdone, (idx - 1) & 1, pv_waited, m_ref, lsum);
            if (!pv_waited) {   // P.V of the set's previous tile has read the P buffer
              mbar_wait(b_pvdone, (idx - 1) & 1);
              tc_fence_after_sync();
            }
            tmem_st8p(p_addr, pk);
            tmem_st8p(p_addr + 8, pk + 8);
            tmem_wait_st();
            tc_fence_before_sync();
            warp_arrive(b_pready, lane);
          }
          mbar_wait(b_pvdone, (nt * hg + nt - 1) & 1);
          tc_fence_after_sync();
          epi_merge4(me, g, 6 * Lg + g, m_ref, lsum);
          if (s == 2 || g == 3) {
            tc_fence_before_sync();
            fence_proxy_async_smem();
            warp_arrive(sbar + S_O_READY, lane);
          }
        }
        mbar_wait(bars + B_QKV_FREE, (5 * Lg + 4) & 1);   // head 3 done by every stream (keeps every warp in step with the barrier)
        if (s == 2) {   // tail stream: quadrants l & 3 / (l + 2) & 3 own the linear part of this layer
          if (q == (l & 3)) me.x = 0; else if (q == ((l + 2) & 3)) me.x = 1; else continue;
        }
        const int x = me.x;
        mbar_wait(bars + B_VEC_FULL, Lg & 1);
        mbar_wait(sbar + S_OUT_DONE, Lg & 1);
        tc_fence_after_sync();
        epi_ln4(me, vec, 1, 6 * Lg + 4);
        tc_fence_before_sync();
        fence_proxy_async_smem();
        warp_arrive(sbar + S_X1_READY, lane);
#pragma unroll 1
        for (int k = 0; k < 4; ++k) {
          const int c = 2 * k + x;   // this set's chunks
          mbar_wait(sbar + S_F1_DONE + 8 * x, k & 1);
          tc_fence_after_sync();
          uint32_t a[32];
          tmem_ld_cols(me.tl + T_F1, a);
          tmem_wait_ld();
          tc_fence_before_sync();
          warp_arrive(sbar + S_F1_FREE + 8 * x, lane);
          uint32_t pk[16];
          act_chunk4(a, vec, c, p.activation, pk);
          const int ch = c >> 1;
          // hidden buffer ch & 1 still feeds the FFN2 partial products of chunk ch - 2
          if (c >= 4) mbar_wait(sbar + S_F2_DONE + 8 * (ch & 1), 0);
          if (me.valid) {
            const uint32_t row = sb + OFF_O + (ch & 1) * kXChunkBytes + me.r * 128;
#pragma unroll
            for (int u = 0; u < 4; ++u)
              st_shared_v4(row + ((((c & 1) * 4 + u) ^ (me.r & 7)) << 4), pk[4 * u], pk[4 * u + 1], pk[4 * u + 2], pk[4 * u + 3]);
          }
          fence_proxy_async_smem();
          warp_arrive(sbar + S_HID_READY + 8 * x, lane);
        }
        mbar_wait(sbar + S_F2_DONE, 1);
        mbar_wait(sbar + S_F2_DONE + 8, 1);
        tc_fence_after_sync();
        epi_ln4(me, vec, 2, 6 * Lg + 5);
        tc_fence_before_sync();
        fence_proxy_async_smem();
        warp_arrive(sbar + S_X2_READY, lane);
      }
      warp_arrive(bars + B_X_DONE, lane);
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == kMmaWarp4) tmem_dealloc(tmem, 512);
}

}  // namespace

bool tc_encoder4_launch(char* x_images, const TcLayer* layers_dev, int num_layers, int activation, int64_t nseq, int sm_count,
                        cudaStream_t st) {
  if (cudaFuncSetAttribute(encoder4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem4) != cudaSuccess) {
    set_error("encoder4_kernel: cannot opt in to %u bytes of shared memory: %s", kSmem4, cudaGetErrorString(cudaGetLastError()));
    return false;
  }
  Enc4Params ep;
  ep.x_images = x_images;
  ep.layers = layers_dev;
  ep.num_layers = num_layers;
  ep.activation = activation;
  ep.nseq = nseq;
  const unsigned grid = (unsigned)(nseq < sm_count ? nseq : sm_count);
  encoder4_kernel<<<grid, kThreads4, kSmem4, st>>>(ep);
  count_launch();
  return check_launch("encoder4_kernel");
}

}  // namespace aft
