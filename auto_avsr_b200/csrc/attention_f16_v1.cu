// Fused rel-pos flash attention, fp16 operands (tcgen05 kind::f16, fp32 accumulate) -- the product path.
//
//   scores[i,j] = ((q_i+u).k_j + (q_i+v).p[rel=i-j]) / 8 ;  keys j >= len[b] masked ;  ctx = softmax_j(scores) @ v
// (transformer/attention.py:174-189 + :59-82 of the reference).  Same algorithm as attention_tc.cu (the TF32
// bring-up version documents the band / skew derivation); with 2-byte operands a head row (d_k = 64) is exactly one
// 128-byte swizzle atom, which doubles the key tile to 128 and halves every tile's shared-memory footprint:
//
// One CTA = one (utterance b, head h, 128-query tile).  Per 128-key tile:
//   MMA  S [128 x 128] = Qu . K_tile^T        4 x tcgen05.mma M128 N128 K16      TMEM cols   0..127
//   MMA  G [128 x 256] = Qv . Pband^T         4 x tcgen05.mma M128 N256 K16      TMEM cols 128..383
//        Pband = the 255 rel-pos table rows m = j-i+T-1 the (query tile, key tile) pair touches (one 3-D TMA box);
//        score (r, c) = S[r][c] + G[r][c + 127 - r]  -- the reference's rel_shift as a per-row accumulator skew.
//   softmax: EIGHT warps, two threads per query row (64 keys each): skew = tcgen05.ld column offset (warp-uniform
//        part) + 5-stage barrel shifter (per-lane part), scale, key mask, online max (halves exchanged through
//        shared memory + a 64-thread named barrier), exp2, P -> shared memory as fp16 in UMMA SWIZZLE_128B layout
//   MMA  O'[128 x 64] = P . V_tile            8 x tcgen05.mma M128 N64 K16       TMEM cols 384..447
//        V stays in its natural (B,H,T,64) layout: the V tile [128 keys][64 d] is the B operand in MN-major form
//        (N = d contiguous), so the QKV projection is ONE GEMM and no transpose of V is ever materialised
//        rescale-accumulated into registers (each thread owns 32 of the 64 output channels of its row)
// Warp roles: warp 0 TMA producer, warp 1 MMA issuer + TMEM owner, warps 2-9 softmax/epilogue.
// Padded QUERY rows are computed like any other row (the reference masks keys only).
#include "common.cuh"
#include "sm100.cuh"

namespace avsr {

using namespace sm100;

constexpr int AH_BQ = 128;    // queries per CTA (= UMMA M)
constexpr int AH_BKV = 128;   // keys per tile
constexpr int AH_BAND = 256;  // rel-pos rows per tile (>= BQ + BKV - 1)
constexpr int AH_THREADS = 320;

// shared memory map (bytes; every tile 1024-aligned; rows of 128 B = 64 halves, SWIZZLE_128B)
constexpr int AH_QU = 0;                          // [128][128B]
constexpr int AH_QV = AH_QU + AH_BQ * 128;        // 16384
constexpr int AH_K = AH_QV + AH_BQ * 128;         // 32768   [128 keys][128B]
constexpr int AH_V = AH_K + AH_BKV * 128;         // 49152   [128 keys][128B = 64 d]  (MN-major B operand)
constexpr int AH_PB = AH_V + AH_BKV * 128;        // 65536   [256][128B]
constexpr int AH_P = AH_PB + AH_BAND * 128;       // 98304   2 atoms x [128 rows][128B = 64 keys]
constexpr int AH_XCH = AH_P + 2 * AH_BQ * 128;    // 131072  float [2 slots][2 halves][128 rows]
constexpr int AH_BARS = AH_XCH + 2 * 2 * 128 * 4; // 133120
constexpr int AH_SMEM = AH_BARS + 128 + 1024;

constexpr uint32_t TH_S = 0, TH_G = 128, TH_O = 384;

// 2^x for x <= 0 (softmax exponents): one MUFU.EX2; results below 2^-126 flush to zero, which is what a probability
// that small contributes anyway.  exp2f() wraps the same instruction in a denormal-range fix-up (FSETP + 2 FMUL).
static __device__ __forceinline__ float ex2_neg(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

static __device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

__global__ void __launch_bounds__(AH_THREADS, 1)
attention_f16_v1_kernel(const __grid_constant__ CUtensorMap tmQu, const __grid_constant__ CUtensorMap tmQv,
                     const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                     const __grid_constant__ CUtensorMap tmP, const int32_t* __restrict__ lengths,
                     __half* __restrict__ ctx, int T, int H) {
  extern __shared__ uint8_t ah_smem_raw[];
  const uint32_t raw = smem_u32(ah_smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gen = ah_smem_raw + (base - raw);
  const uint32_t bars = base + AH_BARS;
  const uint32_t q_full = bars, kp_full = bars + 8, v_full = bars + 16, s_full = bars + 24, p_full = bars + 32,
                 o_full = bars + 40;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gen + AH_BARS + 64);
  float* xch = reinterpret_cast<float*>(gen + AH_XCH);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // grid = (H, B, query tiles): the query tile is the SLOWEST block index, so the last tile of every utterance
  // (mostly rows >= T, e.g. 16 valid of 128 at T = 400) is scheduled in the final, partially filled wave
  const int h = blockIdx.x, b = blockIdx.y, i0 = blockIdx.z * AH_BQ;
  const int bh = b * H + h;
  pdl_launch_dependents();
  int L = T;
  if (lengths) { L = lengths[b]; L = L < 0 ? 0 : (L > T ? T : L); }   // lengths: written before the graph, not by the predecessor
  const int nkt = (L + AH_BKV - 1) / AH_BKV;
#ifdef AVSR_TRACE
  // phase marks: 0 prologue done, 1 dependency resolved, 2 first S/G MMAs issued, 3 softmax sees S/G(0),
  // 4 softmax published P(0), 5 softmax sees O(0), 6 softmax warp done with the last tile, 7 CTA drained
  unsigned long long** trc = reinterpret_cast<unsigned long long**>(gen + AH_BARS + 96);
  if (threadIdx.x == 0) AVSR_TRACE_OPEN(trc, 301, (unsigned)nkt | ((unsigned)blockIdx.z << 8));
#endif

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQu); tma_prefetch_desc(&tmQv); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
    tma_prefetch_desc(&tmP);
    mbar_init(q_full, 1); mbar_init(kp_full, 1); mbar_init(v_full, 1); mbar_init(s_full, 1);
    mbar_init(p_full, 256); mbar_init(o_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(smem_u32(tmem_slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  AVSR_TRACE_MARK(threadIdx.x == 0, trc, 0);
  pdl_wait();
  AVSR_TRACE_MARK(threadIdx.x == 0, trc, 1);
  AVSR_TRACE_STAMP(threadIdx.x == 0, trc, 10);

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0 && nkt > 0) {
      mbar_expect_tx(q_full, 2 * AH_BQ * 128);
      tma_load_2d(base + AH_QU, &tmQu, 0, bh * T + i0, q_full);
      tma_load_2d(base + AH_QV, &tmQv, 0, bh * T + i0, q_full);
      for (int it = 0; it < nkt; ++it) {
        const int j0 = it * AH_BKV;
        if (it > 0) mbar_wait(s_full, (it - 1) & 1);          // S/G MMAs of the previous tile retired: K, Pband free
        mbar_expect_tx(kp_full, AH_BKV * 128 + AH_BAND * 128);
        tma_load_2d(base + AH_K, &tmK, 0, bh * T + j0, kp_full);
        const int m_lo = j0 - i0 - (AH_BQ - 1) + T - 1;       // first table row of the band (may be < 0: zero fill)
        tma_load_3d(base + AH_PB, &tmP, 0, m_lo, h, kp_full);
        if (it > 0) mbar_wait(o_full, (it - 1) & 1);          // P.V of the previous tile retired: V free
        mbar_expect_tx(v_full, AH_BKV * 128);
        tma_load_2d(base + AH_V, &tmV, 0, bh * T + j0, v_full);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (lane == 0 && nkt > 0) {
      constexpr uint32_t idesc_s = umma_idesc_f16(AH_BQ, AH_BKV);
      constexpr uint32_t idesc_g = umma_idesc_f16(AH_BQ, AH_BAND);
      constexpr uint32_t idesc_o = umma_idesc_f16_bmn(AH_BQ, 64);
      auto issue_scores = [&]() {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)   // d_k = 64 halves = 4 MMA-K steps of 32 bytes inside one atom
          mma_f16(tmem + TH_S, umma_desc_sw128(base + AH_QU + ks * 32), umma_desc_sw128(base + AH_K + ks * 32),
                  idesc_s, ks != 0);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
          mma_f16(tmem + TH_G, umma_desc_sw128(base + AH_QV + ks * 32), umma_desc_sw128(base + AH_PB + ks * 32),
                  idesc_g, ks != 0);
        tc_commit(s_full);
      };
      mbar_wait(q_full, 0);
      mbar_wait(kp_full, 0);
      tc_fence_after();
      issue_scores();
      AVSR_TRACE_MARK(true, trc, 2);
      for (int it = 0; it < nkt; ++it) {
        mbar_wait(p_full, it & 1);   // softmax consumed S/G(it) and published P(it)
        mbar_wait(v_full, it & 1);
        tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)   // A = P: 2 K-major atoms x 4 steps; B = V: 16 key rows (2 KB) per step
          mma_f16(tmem + TH_O, umma_desc_sw128(base + AH_P + (ks >> 2) * (AH_BQ * 128) + (ks & 3) * 32),
                  umma_desc_sw128(base + AH_V + ks * (16 * 128)), idesc_o, ks != 0);
        tc_commit(o_full);
        if (it + 1 < nkt) {
          mbar_wait(kp_full, (it + 1) & 1);
          tc_fence_after();
          issue_scores();
        }
      }
    }
  } else {
    // ------------------------------------------------------------ softmax + epilogue: two threads per query row
    const int q = warp & 3;                       // TMEM lane quarter
    const int hf = (warp - 2) >> 2;               // which 64-key half of the tile / which 32 output channels
    const int r = q * 32 + lane;                  // row inside the query tile
    const int i = i0 + r;
    const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16);
    const int gbase = 96 - 32 * q;                // warp-uniform part of the skew 127 - r = gbase + (31 - lane)
    const int sh = 31 - lane;
    const float kScale = 0.125f * 1.4426950408889634f;   // 1/sqrt(64) * log2(e)
    float o[32];
#pragma unroll
    for (int d = 0; d < 32; ++d) o[d] = 0.f;
    float m_run = -INFINITY, l_run = 0.f;
    uint8_t* prow = gen + AH_P + hf * (AH_BQ * 128) + r * 128;
    if (i0 + q * 32 >= T) {
      // all 32 query rows of this warp lie beyond T (last query tile): nothing to compute or store -- only keep the
      // barrier protocol going (its P rows / accumulator rows are never read by anyone)
      for (int it = 0; it < nkt; ++it) {
        mbar_arrive(p_full);
        mbar_wait(o_full, it & 1);
      }
    } else {

    for (int it = 0; it < nkt; ++it) {
      const int j0 = it * AH_BKV + hf * 64;       // first key this thread scores
      mbar_wait(s_full, it & 1);
      tc_fence_after();
      AVSR_TRACE_MARK(it == 0 && threadIdx.x == 64, trc, 3);
      float s[64];   // raw (unscaled) scores of this thread's 64 keys
#pragma unroll
      for (int c0 = 0; c0 < 64; c0 += 32) {
        const int jc = j0 + c0;                   // first key of this 32-key chunk (warp-uniform)
        if (jc >= L) {                            // chunk entirely beyond the utterance: no loads, no skew, no exp
#pragma unroll
          for (int c = 0; c < 32; ++c) s[c0 + c] = -INFINITY;
          continue;
        }
        float x[64];
        tmem_ld32(trow + TH_G + gbase + hf * 64 + c0, x);
        tmem_ld32(trow + TH_G + gbase + hf * 64 + c0 + 32, x + 32);
        tmem_ld32(trow + TH_S + hf * 64 + c0, s + c0);
        tmem_ld_wait();
        // y[c] = x[c + sh], sh in [0,31]: barrel shifter, one stage per bit of sh.  Written as per-element selects
        // (in-place is safe in increasing c: x[c + 2^k] is still the previous stage's value); the if/else form made
        // the compiler emit divergent branches with register copies on both paths (2x the moves, r01 SASS).
        {
          const bool b16 = (sh & 16) != 0, b8 = (sh & 8) != 0, b4 = (sh & 4) != 0, b2 = (sh & 2) != 0, b1 = (sh & 1) != 0;
#pragma unroll
          for (int c = 0; c < 47; ++c) x[c] = b16 ? x[c + 16] : x[c];
#pragma unroll
          for (int c = 0; c < 39; ++c) x[c] = b8 ? x[c + 8] : x[c];
#pragma unroll
          for (int c = 0; c < 35; ++c) x[c] = b4 ? x[c + 4] : x[c];
#pragma unroll
          for (int c = 0; c < 33; ++c) x[c] = b2 ? x[c + 2] : x[c];
#pragma unroll
          for (int c = 0; c < 32; ++c) x[c] = b1 ? x[c + 1] : x[c];
        }
        if (jc + 32 <= L) {                       // fully valid chunk: no per-key mask
#pragma unroll
          for (int c = 0; c < 32; ++c) s[c0 + c] += x[c];
        } else {
#pragma unroll
          for (int c = 0; c < 32; ++c) s[c0 + c] = (jc + c < L) ? s[c0 + c] + x[c] : -INFINITY;
        }
      }
      float mloc = -INFINITY;
#pragma unroll
      for (int c = 0; c < 64; ++c) mloc = fmaxf(mloc, s[c]);
      // row max over both halves: exchange through shared memory (slot it&1), 64-thread named barrier per quarter
      float* slot = xch + (it & 1) * 256;
      slot[hf * 128 + r] = mloc;
      named_bar_sync(1 + q, 64);
      const float mx = fmaxf(m_run, fmaxf(mloc, slot[(hf ^ 1) * 128 + r]));   // finite: key (tile start) < L is valid
      const float alpha = ex2_neg((m_run - mx) * kScale);   // first tile: exp2(-inf) = 0
      m_run = mx;
      const float mxs = mx * kScale;
      float sum = 0.f;
      // P row (64 halves = 128 B = one atom row) -> shared as fp16, 16-byte chunk index XOR (r & 7)
      // p = exp2(s * kScale - mx * kScale): scale folded into one FFMA per element
#pragma unroll
      for (int ch = 0; ch < 8; ++ch) {
        uint32_t pk[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float p0 = ex2_neg(fmaf(s[8 * ch + 2 * e], kScale, -mxs));
          const float p1 = ex2_neg(fmaf(s[8 * ch + 2 * e + 1], kScale, -mxs));
          sum += p0 + p1;
          const __half2 hp = __floats2half2_rn(p0, p1);
          pk[e] = *reinterpret_cast<const uint32_t*>(&hp);
        }
        *reinterpret_cast<uint4*>(prow + ((ch ^ (r & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      }
      l_run = l_run * alpha + sum;
      fence_proxy_async();      // generic-proxy smem writes -> visible to the tensor core (async proxy)
      tc_fence_before();        // our tcgen05.ld of S/G are complete before the MMA warp overwrites them
      mbar_arrive(p_full);
      AVSR_TRACE_MARK(it == 0 && threadIdx.x == 64, trc, 4);
      mbar_wait(o_full, it & 1);
      tc_fence_after();
      AVSR_TRACE_MARK(it == 0 && threadIdx.x == 64, trc, 5);
      {
        float pv[32];
        tmem_ld32(trow + TH_O + hf * 32, pv);
        tmem_ld_wait();
#pragma unroll
        for (int d = 0; d < 32; ++d) o[d] = fmaf(o[d], alpha, pv[d]);
      }
    }
    AVSR_TRACE_MARK(threadIdx.x == 64, trc, 6);
    // total row sum = both halves' partial sums (same running max in both threads)
    float* slot = xch + (nkt & 1) * 256;
    slot[hf * 128 + r] = l_run;
    named_bar_sync(1 + q, 64);
    const float l_tot = l_run + slot[(hf ^ 1) * 128 + r];
    if (i < T) {
      const float inv = l_tot > 0.f ? 1.0f / l_tot : 0.f;   // len == 0: zeros, like softmax(...).masked_fill(mask, 0)
      __half* dst = ctx + ((long)b * T + i) * (H * kHeadDim) + h * kHeadDim + hf * 32;
#pragma unroll
      for (int d = 0; d < 32; d += 4) store_op4<__half>(dst + d, o[d] * inv, o[d + 1] * inv, o[d + 2] * inv, o[d + 3] * inv);
    }
    }  // valid warp
  }
  tc_fence_before();
  __syncthreads();
  AVSR_TRACE_MARK(threadIdx.x == 0, trc, 7);
  AVSR_TRACE_STAMP(threadIdx.x == 0, trc, 11);
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem);
  }
}

AVSR_TRACE_DEFINE_BIND(trace_bind_attention_f16_v1)

int attention_f16_v1(const __half* qu, const __half* qv, const __half* kk, const __half* vv, const __half* pos,
                  const int32_t* lengths, __half* ctx, int B, int T, int H, int Rp, cudaStream_t st) {
  AVSR_REQUIRE(Rp >= 2 * T - 1, "attention_f16: bad Rp=%d for T=%d", Rp, T);
  if (B <= 0 || T <= 0) return AVSR_OK;
  CUtensorMap tmQu, tmQv, tmK, tmV, tmP;
  const uint64_t rows = (uint64_t)B * H * T;
  AVSR_TRY(make_tmap_2d(&tmQu, qu, rows, 64, 64, AH_BQ, 2));
  AVSR_TRY(make_tmap_2d(&tmQv, qv, rows, 64, 64, AH_BQ, 2));
  AVSR_TRY(make_tmap_2d(&tmK, kk, rows, 64, 64, AH_BKV, 2));
  AVSR_TRY(make_tmap_2d(&tmV, vv, rows, 64, 64, AH_BKV, 2));
  AVSR_TRY(make_tmap_3d(&tmP, pos, (uint64_t)H, (uint64_t)Rp, 64, 64, (uint64_t)Rp * 64, AH_BAND, 2));
  AVSR_SET_MAX_SMEM(attention_f16_v1_kernel, AH_SMEM);
  dim3 grid(H, B, cdiv(T, AH_BQ));
  AVSR_LAUNCH(attention_f16_v1_kernel, grid, AH_THREADS, AH_SMEM, st, tmQu, tmQv, tmK, tmV, tmP, lengths, ctx, T, H);
  return AVSR_OK;
}

}  // namespace avsr
