// Fused rel-pos flash attention, fp16 operands (tcgen05 kind::f16, fp32 accumulate) -- the product path, round 2.
//
//   scores[i,j] = ((q_i+u).k_j + (q_i+v).p[rel=i-j]) / 8 ;  keys j >= len[b] masked ;  ctx = softmax_j(scores) @ v
// (transformer/attention.py:174-189 + :59-82 of the reference).
//
// Why this shape (r02 launch timeline, profiles/r02_timeline_S2_base.txt): the round-1 kernel (128-key tiles, one CTA
// per SM; deleted after the r02 A/B: 29.3 vs 21.7 us per launch in situ) is a chain of dependent latencies -- S/G MMA -> TMEM load -> skew -> max -> exp -> P ->
// P.V MMA -> O load -- with nothing to overlap it, and at T = 400 its 192 CTAs need two waves on 148 SMs, the second
// (16 valid query rows per CTA) as long as the first: 28 us per launch for 3 us of instruction issue.  This version is
// built so that TWO CTAs share an SM (one wave of 296 slots; one CTA's latencies hide under the other's work):
//   * 32-key tiles: TMEM = S 32 + G 160 + O 64 = 256 columns, shared memory 80 KB, <= 102 registers per thread;
//   * O accumulates IN TENSOR MEMORY across the key tiles (tcgen05.mma accumulate), so there is no per-tile O load and
//     no 32-register accumulator per thread.  The softmax keeps a REFERENCE maximum per row instead of the running
//     maximum: probabilities are exp2(s - ref), and only when a tile's maximum exceeds ref by more than 2^8 the row's
//     O columns (tcgen05.ld / st) and running sum are rescaled and ref moves -- p stays <= 256 (exact in fp16/fp32
//     terms: the final O / l is the same softmax), and after the first tiles rescales are rare;
//   * the context tile leaves through shared memory and ONE TMA store (3-D map, rows beyond T are clipped) instead of
//     8-byte lane-per-row stores.
//
// One CTA = one (utterance b, head h, 128-query tile).  Per 32-key tile:
//   MMA  S [128 x 32]  = Qu . K_tile^T        4 x tcgen05.mma M128 N32  K16      TMEM cols   0..31
//   MMA  G [128 x 160] = Qv . Pband^T         4 x tcgen05.mma M128 N160 K16      TMEM cols  32..191
//        Pband = the 159 rel-pos table rows m = j-i+T-1 the (query tile, key tile) pair touches (one 3-D TMA box);
//        score (r, c) = S[r][c] + G[r][c + 127 - r]  -- the reference's rel_shift as a per-row accumulator skew.
//   softmax: eight warps, two threads per query row (16 keys each): skew = tcgen05.ld column offset (warp-uniform part)
//        + 5-stage select barrel shifter (per-lane part), key mask, tile max (halves exchanged through shared memory
//        + a 64-thread named barrier), exp2 against the row's reference maximum, P -> shared memory as fp16 (UMMA
//        SWIZZLE_128B rows, first 64 bytes used)
//   MMA  O [128 x 64] += P . V_tile           2 x tcgen05.mma M128 N64 K16       TMEM cols 192..255
//        V stays in its natural (B,H,T,64) layout (MN-major B operand), as in round 1.
// Warp roles: warp 0 TMA producer, warp 1 MMA issuer + TMEM owner, warps 2-9 softmax / epilogue.
// Padded QUERY rows are computed like any other row (the reference masks keys only).
#include "common.cuh"
#include "sm100.cuh"

namespace avsr {

using namespace sm100;

constexpr int A2_BQ = 128;    // queries per CTA (= UMMA M)
constexpr int A2_BKV = 32;    // keys per tile
constexpr int A2_BAND = 160;  // rel-pos rows per tile (>= BQ + BKV - 1 = 159, multiple of 16 for UMMA N)
constexpr int A2_THREADS = 320;

// shared memory map (bytes; every tile 1024-aligned; rows of 128 B = 64 halves, SWIZZLE_128B)
constexpr int A2_QU = 0;                           // [128][128B]   (re-used as the output staging tile at the end)
constexpr int A2_QV = A2_QU + A2_BQ * 128;         // 16384
constexpr int A2_K = A2_QV + A2_BQ * 128;          // 32768   [32 keys][128B]
constexpr int A2_V = A2_K + A2_BKV * 128;          // 36864   [32 keys][128B = 64 d]  (MN-major B operand)
constexpr int A2_PB = A2_V + A2_BKV * 128;         // 40960   [160][128B]
constexpr int A2_P = A2_PB + A2_BAND * 128;        // 61440   [128 rows][128B: 32 keys = first 64 B]
constexpr int A2_XCH = A2_P + A2_BQ * 128;         // 77824   float [2 slots][2 halves][128 rows]
constexpr int A2_BARS = A2_XCH + 2 * 2 * 128 * 4;  // 79872
constexpr int A2_SMEM = A2_BARS + 128 + 1024;      // 81024: two CTAs per SM

constexpr uint32_t T2_S = 0, T2_G = 32, T2_O = 192;
constexpr float kRescaleLog2 = 8.0f;   // a row's reference maximum moves when a tile's maximum exceeds it by 2^8

// 2^x for x <= ~8 (softmax exponents): one MUFU.EX2; results below 2^-126 flush to zero, which is what a probability
// that small contributes anyway.
__device__ __forceinline__ float ex2_fast(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ void a2_named_bar(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

__global__ void __launch_bounds__(A2_THREADS, 2)
attention_f16_kernel(const __grid_constant__ CUtensorMap tmQu, const __grid_constant__ CUtensorMap tmQv,
                     const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                     const __grid_constant__ CUtensorMap tmP, const __grid_constant__ CUtensorMap tmCtx,
                     const int32_t* __restrict__ lengths, int T, int H) {
  extern __shared__ uint8_t a2_smem_raw[];
  const uint32_t raw = smem_u32(a2_smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gen = a2_smem_raw + (base - raw);
  const uint32_t bars = base + A2_BARS;
  const uint32_t q_full = bars, kp_full = bars + 8, v_full = bars + 16, s_full = bars + 24, p_full = bars + 32,
                 o_full = bars + 40;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gen + A2_BARS + 64);
  float* xch = reinterpret_cast<float*>(gen + A2_XCH);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // grid = (H, B, query tiles): the query tile is the SLOWEST block index, so the partially filled last tile of every
  // utterance is scheduled last
  const int h = blockIdx.x, b = blockIdx.y, i0 = blockIdx.z * A2_BQ;
  const int bh = b * H + h;
  pdl_launch_dependents();
  int L = T;
  if (lengths) { L = lengths[b]; L = L < 0 ? 0 : (L > T ? T : L); }   // lengths: written before the graph, not by the predecessor
  const int nkt = (L + A2_BKV - 1) / A2_BKV;
#ifdef AVSR_TRACE
  // phase marks: 0 prologue done, 1 dependency resolved, 2 first S/G MMAs issued, 3 softmax sees S/G(0),
  // 4 softmax published P(0), 6 softmax warp done with the last tile, 7 CTA drained
  unsigned long long** trc = reinterpret_cast<unsigned long long**>(gen + A2_BARS + 96);
  if (threadIdx.x == 0) AVSR_TRACE_OPEN(trc, 300, (unsigned)nkt | ((unsigned)blockIdx.z << 8));
#endif

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQu); tma_prefetch_desc(&tmQv); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
    tma_prefetch_desc(&tmP); tma_prefetch_desc(&tmCtx);
    mbar_init(q_full, 1); mbar_init(kp_full, 1); mbar_init(v_full, 1); mbar_init(s_full, 1);
    mbar_init(p_full, 256); mbar_init(o_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<256>(smem_u32(tmem_slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  AVSR_TRACE_MARK(threadIdx.x == 0, trc, 0);
  pdl_wait();
  AVSR_TRACE_MARK(threadIdx.x == 0, trc, 1);
  AVSR_TRACE_STAMP(threadIdx.x == 0, trc, 10);

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (warp-uniform; one elected lane issues)
    if (nkt > 0) {
      if (elect_one_sync()) {
        mbar_expect_tx(q_full, 2 * A2_BQ * 128);
        tma_load_2d(base + A2_QU, &tmQu, 0, bh * T + i0, q_full);
        tma_load_2d(base + A2_QV, &tmQv, 0, bh * T + i0, q_full);
      }
      for (int it = 0; it < nkt; ++it) {
        const int j0 = it * A2_BKV;
        if (it > 0) mbar_wait(s_full, (it - 1) & 1);          // S/G MMAs of the previous tile retired: K, Pband free
        if (elect_one_sync()) {
          mbar_expect_tx(kp_full, A2_BKV * 128 + A2_BAND * 128);
          tma_load_2d(base + A2_K, &tmK, 0, bh * T + j0, kp_full);
          const int m_lo = j0 - i0 - (A2_BQ - 1) + T - 1;     // first table row of the band (may be < 0: zero fill)
          tma_load_3d(base + A2_PB, &tmP, 0, m_lo, h, kp_full);
        }
        if (it > 0) mbar_wait(o_full, (it - 1) & 1);          // P.V of the previous tile retired: V free
        if (elect_one_sync()) {
          mbar_expect_tx(v_full, A2_BKV * 128);
          tma_load_2d(base + A2_V, &tmV, 0, bh * T + j0, v_full);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (warp-uniform; one elected lane issues)
    if (nkt > 0) {
      constexpr uint32_t idesc_s = umma_idesc_f16(A2_BQ, A2_BKV);
      constexpr uint32_t idesc_g = umma_idesc_f16(A2_BQ, A2_BAND);
      constexpr uint32_t idesc_o = umma_idesc_f16_bmn(A2_BQ, 64);
      // every operand tile sits at a fixed shared-memory address: the descriptors are loop-invariant
      const uint64_t d_qv = umma_desc_sw128(base + A2_QV), d_pb = umma_desc_sw128(base + A2_PB);
      const uint64_t d_qu = umma_desc_sw128(base + A2_QU), d_k = umma_desc_sw128(base + A2_K);
      const uint64_t d_p = umma_desc_sw128(base + A2_P), d_v = umma_desc_sw128(base + A2_V);
      const uint32_t t_s = tmem + T2_S, t_g = tmem + T2_G, t_o = tmem + T2_O;
      auto issue_scores = [&]() {
        if (elect_one_sync()) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)   // d_k = 64 halves = 4 MMA-K steps of 32 bytes (+2 in the 16-byte address field)
            mma_f16(t_g, d_qv + 2 * ks, d_pb + 2 * ks, idesc_g, ks != 0);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            mma_f16(t_s, d_qu + 2 * ks, d_k + 2 * ks, idesc_s, ks != 0);
          tc_commit(s_full);
        }
      };
      mbar_wait(q_full, 0);
      mbar_wait(kp_full, 0);
      tc_fence_after();
      issue_scores();
      AVSR_TRACE_MARK(lane == 0, trc, 2);
      for (int it = 0; it < nkt; ++it) {
        mbar_wait(p_full, it & 1);   // softmax consumed S/G(it), rescaled O if needed and published P(it)
        mbar_wait(v_full, it & 1);
        tc_fence_after();
        if (elect_one_sync()) {
          // A = P: K-major, 32 keys = 2 steps of 32 bytes; B = V: 16 key rows (2 KB = 128 in the address field) per step
          mma_f16(t_o, d_p, d_v, idesc_o, it != 0);
          mma_f16(t_o, d_p + 2, d_v + 128, idesc_o, 1);
          tc_commit(o_full);
        }
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
    const int hf = (warp - 2) >> 2;               // which 16-key half of the tile / which 32 output channels
    const int r = q * 32 + lane;                  // row inside the query tile
    const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16);
    const int gbase = 96 - 32 * q;                // warp-uniform part of the skew 127 - r = gbase + (31 - lane)
    const int sh = 31 - lane;
    const float kScale = 0.125f * 1.4426950408889634f;   // 1/sqrt(64) * log2(e)
    const bool warp_valid = i0 + q * 32 < T;      // else: all 32 query rows of this warp lie beyond T (last query tile)
    float t_ref = -INFINITY, l_run = 0.f;         // reference maximum (log2 domain, scaled), running sum of p
    uint8_t* prow = gen + A2_P + r * 128;

    for (int it = 0; it < nkt; ++it) {
      if (!warp_valid) {            // nothing to compute or store: keep the barrier protocol going, one phase at a time
        mbar_arrive(p_full);
        mbar_wait(o_full, it & 1);
        continue;
      }
      const int jc = it * A2_BKV + hf * 16;       // first key this thread scores (warp-uniform)
      mbar_wait(s_full, it & 1);
      tc_fence_after();
      AVSR_TRACE_MARK(it == 0 && threadIdx.x == 64, trc, 3);
      float s[16];   // raw (unscaled) scores of this thread's 16 keys
      if (jc >= L) {                              // chunk entirely beyond the utterance: no loads, no skew
#pragma unroll
        for (int c = 0; c < 16; ++c) s[c] = -INFINITY;
      } else {
        float x[48];
        tmem_ld32(trow + T2_G + gbase + hf * 16, x);
        tmem_ld16(trow + T2_G + gbase + hf * 16 + 32, x + 32);
        tmem_ld16(trow + T2_S + hf * 16, s);
        tmem_ld_wait();
        // y[c] = x[c + sh], sh in [0,31]: barrel shifter, one stage per bit of sh, as per-element selects (in-place is
        // safe in increasing c: x[c + 2^k] is still the previous stage's value)
        {
          const bool b16 = (sh & 16) != 0, b8 = (sh & 8) != 0, b4 = (sh & 4) != 0, b2 = (sh & 2) != 0, b1 = (sh & 1) != 0;
#pragma unroll
          for (int c = 0; c < 31; ++c) x[c] = b16 ? x[c + 16] : x[c];
#pragma unroll
          for (int c = 0; c < 23; ++c) x[c] = b8 ? x[c + 8] : x[c];
#pragma unroll
          for (int c = 0; c < 19; ++c) x[c] = b4 ? x[c + 4] : x[c];
#pragma unroll
          for (int c = 0; c < 17; ++c) x[c] = b2 ? x[c + 2] : x[c];
#pragma unroll
          for (int c = 0; c < 16; ++c) x[c] = b1 ? x[c + 1] : x[c];
        }
        if (jc + 16 <= L) {                       // fully valid chunk: no per-key mask
#pragma unroll
          for (int c = 0; c < 16; ++c) s[c] += x[c];
        } else {
#pragma unroll
          for (int c = 0; c < 16; ++c) s[c] = (jc + c < L) ? s[c] + x[c] : -INFINITY;
        }
      }
      float mloc = -INFINITY;
#pragma unroll
      for (int c = 0; c < 16; ++c) mloc = fmaxf(mloc, s[c]);
      // tile max over both halves: exchange through shared memory (slot it&1), 64-thread named barrier per quarter
      float* slot = xch + (it & 1) * 256;
      slot[hf * 128 + r] = mloc;
      a2_named_bar(1 + q, 64);
      const float tmax = fmaxf(mloc, slot[(hf ^ 1) * 128 + r]) * kScale;   // finite: the tile's first key is < L
      const bool move = tmax > t_ref + kRescaleLog2;   // same decision in both threads of the row (same tmax, same t_ref)
      const float alpha = move ? ex2_fast(t_ref - tmax) : 1.0f;            // first tile: exp2(-inf) = 0
      if (move) { l_run *= alpha; t_ref = tmax; }
      // tcgen05.ld / st are warp-collective (.sync.aligned): when ANY row of the warp moves its reference, the whole
      // warp rescales its 32 x 32 block of O in tensor memory (alpha = 1 for the rows that stay).  P.V(it-1) has
      // retired: S/G(it), which we just observed, were issued after it.
      if (it > 0 && __any_sync(0xffffffffu, move)) {
        float ov[32];
        tmem_ld32(trow + T2_O + hf * 32, ov);
        tmem_ld_wait();
#pragma unroll
        for (int d = 0; d < 32; ++d) ov[d] *= alpha;
        tmem_st32(trow + T2_O + hf * 32, ov);
        tmem_st_wait();
      }
      float sum = 0.f;
      // P row: 16 halves = 32 B = 16-byte chunks 2*hf, 2*hf+1 of the row, chunk index XOR (r & 7) (SWIZZLE_128B)
#pragma unroll
      for (int ch = 0; ch < 2; ++ch) {
        uint32_t pk[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float p0 = ex2_fast(fmaf(s[8 * ch + 2 * e], kScale, -t_ref));
          const float p1 = ex2_fast(fmaf(s[8 * ch + 2 * e + 1], kScale, -t_ref));
          sum += p0 + p1;
          const __half2 hp = __floats2half2_rn(p0, p1);
          pk[e] = *reinterpret_cast<const uint32_t*>(&hp);
        }
        *reinterpret_cast<uint4*>(prow + (((2 * hf + ch) ^ (r & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      }
      l_run += sum;
      fence_proxy_async();      // generic-proxy smem writes -> visible to the tensor core (async proxy)
      tc_fence_before();        // our tcgen05.ld / st are complete before the MMA warp touches S, G, O again
      mbar_arrive(p_full);
      AVSR_TRACE_MARK(it == 0 && threadIdx.x == 64, trc, 4);
    }
    AVSR_TRACE_MARK(threadIdx.x == 64, trc, 6);
    // ---- epilogue: O / l -> fp16 -> staging tile (the Qu buffer: its last reader, S(nkt-1), retired long ago)
    uint8_t* orow = gen + A2_QU + r * 128;
    if (warp_valid) {
      float ov[32];
      float inv = 0.f;
      if (nkt > 0) {
        float* slot = xch + (nkt & 1) * 256;      // total row sum = both halves' partial sums (same reference maximum)
        slot[hf * 128 + r] = l_run;
        a2_named_bar(1 + q, 64);
        const float l_tot = l_run + slot[(hf ^ 1) * 128 + r];
        inv = l_tot > 0.f ? 1.0f / l_tot : 0.f;
        mbar_wait(o_full, (nkt - 1) & 1);
        tc_fence_after();
        tmem_ld32(trow + T2_O + hf * 32, ov);
        tmem_ld_wait();
      } else {                                    // len == 0: zeros, like softmax(...).masked_fill(mask, 0)
#pragma unroll
        for (int d = 0; d < 32; ++d) ov[d] = 0.f;
      }
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {            // this thread's 32 channels = chunks 4*hf .. 4*hf+3 of the row
        uint32_t pk[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const __half2 hp = __halves2half2(to_half_sat(ov[8 * ch + 2 * e] * inv), to_half_sat(ov[8 * ch + 2 * e + 1] * inv));
          pk[e] = *reinterpret_cast<const uint32_t*>(&hp);
        }
        *reinterpret_cast<uint4*>(orow + (((4 * hf + ch) ^ (r & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      }
      fence_proxy_async();
    }
    a2_named_bar(5, 256);                         // all eight softmax warps: the staging tile is complete
    if (threadIdx.x == 64) {
      tma_store_3d(&tmCtx, base + A2_QU, h * kHeadDim, i0, b);   // rows >= T are clipped by the tensor map
      tma_store_commit();
      tma_store_wait_read();
    }
  }
  tc_fence_before();
  __syncthreads();
  AVSR_TRACE_MARK(threadIdx.x == 0, trc, 7);
  AVSR_TRACE_STAMP(threadIdx.x == 0, trc, 11);
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<256>(tmem);
  }
}

AVSR_TRACE_DEFINE_BIND(trace_bind_attention_f16)

int attention_f16(const __half* qu, const __half* qv, const __half* kk, const __half* vv, const __half* pos,
                  const int32_t* lengths, __half* ctx, int B, int T, int H, int Rp, cudaStream_t st) {
  AVSR_REQUIRE(Rp >= 2 * T - 1, "attention_f16: bad Rp=%d for T=%d", Rp, T);
  if (B <= 0 || T <= 0) return AVSR_OK;
  CUtensorMap tmQu, tmQv, tmK, tmV, tmP, tmCtx;
  const uint64_t rows = (uint64_t)B * H * T;
  const uint64_t D = (uint64_t)H * kHeadDim;
  AVSR_TRY(make_tmap_2d(&tmQu, qu, rows, 64, 64, A2_BQ, 2));
  AVSR_TRY(make_tmap_2d(&tmQv, qv, rows, 64, 64, A2_BQ, 2));
  AVSR_TRY(make_tmap_2d(&tmK, kk, rows, 64, 64, A2_BKV, 2));
  AVSR_TRY(make_tmap_2d(&tmV, vv, rows, 64, 64, A2_BKV, 2));
  AVSR_TRY(make_tmap_3d(&tmP, pos, (uint64_t)H, (uint64_t)Rp, 64, 64, (uint64_t)Rp * 64, A2_BAND, 2));
  AVSR_TRY(make_tmap_3d(&tmCtx, ctx, (uint64_t)B, (uint64_t)T, D, D, (uint64_t)T * D, A2_BQ, 2));
  AVSR_SET_MAX_SMEM(attention_f16_kernel, A2_SMEM);
  dim3 grid(H, B, cdiv(T, A2_BQ));
  AVSR_LAUNCH(attention_f16_kernel, grid, A2_THREADS, A2_SMEM, st, tmQu, tmQv, tmK, tmV, tmP, tmCtx, lengths, T, H);
  return AVSR_OK;
}

}  // namespace avsr
