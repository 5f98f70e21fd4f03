// EXPERIMENTAL variant of attention_f16.cu (selected with AVSR_B200_ATTN=x4; not the default, not yet run on a B200):
// the same tiles, barriers, TMEM map and MMA schedule, but SIXTEEN softmax warps -- four threads per query row, each
// scoring 32 of the tile's 128 keys and owning 16 of the 64 output channels.
//
// Why (r01 analysis): per 128-key tile the softmax warps issue ~900 instructions per thread =
// ~1.8 K issue cycles per scheduler, and the TMEM reads cost ~3.6 K cycles, but a tile takes ~9.5 K cycles: with two
// warps per scheduler the tcgen05.ld -> skew -> exp2 -> st.shared chain is latency-bound.  Four warps per scheduler
// halve the per-thread chain and double the warps available to hide it; the bytes read from TMEM do not change
// (each 32-key chunk still needs a 64-column window of G for the per-lane skew).
// Register budget: 576 threads -> 112 registers per thread, so the G window (64 registers) is loaded and skewed
// BEFORE the 32 S values are loaded (peak = window + 16 output accumulators + bookkeeping).
#include "common.cuh"
#include "sm100.cuh"

namespace avsr {

using namespace sm100;

constexpr int AX_BQ = 128;    // queries per CTA (= UMMA M)
constexpr int AX_BKV = 128;   // keys per tile
constexpr int AX_BAND = 256;  // rel-pos rows per tile (>= BQ + BKV - 1)
constexpr int AX_SOFTMAX_WARPS = 16;
constexpr int AX_THREADS = 64 + 32 * AX_SOFTMAX_WARPS;   // 576

// shared memory map (bytes; every tile 1024-aligned; rows of 128 B = 64 halves, SWIZZLE_128B)
constexpr int AX_QU = 0;                          // [128][128B]
constexpr int AX_QV = AX_QU + AX_BQ * 128;        // 16384
constexpr int AX_K = AX_QV + AX_BQ * 128;         // 32768   [128 keys][128B]
constexpr int AX_V = AX_K + AX_BKV * 128;         // 49152   [128 keys][128B = 64 d]  (MN-major B operand)
constexpr int AX_PB = AX_V + AX_BKV * 128;        // 65536   [256][128B]
constexpr int AX_P = AX_PB + AX_BAND * 128;       // 98304   2 atoms x [128 rows][128B = 64 keys]
constexpr int AX_XCH = AX_P + 2 * AX_BQ * 128;    // 131072  float [2 slots][4 parts][128 rows]
constexpr int AX_BARS = AX_XCH + 2 * 4 * 128 * 4; // 135168
constexpr int AX_SMEM = AX_BARS + 128 + 1024;

constexpr uint32_t TX_S = 0, TX_G = 128, TX_O = 384;

__device__ __forceinline__ float ex2_neg_x(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void named_bar_sync_x(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
// 32 lanes x 16 consecutive fp32 columns
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

__global__ void __launch_bounds__(AX_THREADS, 1)
attention_f16x_kernel(const __grid_constant__ CUtensorMap tmQu, const __grid_constant__ CUtensorMap tmQv,
                      const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                      const __grid_constant__ CUtensorMap tmP, const int32_t* __restrict__ lengths,
                      __half* __restrict__ ctx, int T, int H) {
  extern __shared__ uint8_t ax_smem_raw[];
  const uint32_t raw = smem_u32(ax_smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gen = ax_smem_raw + (base - raw);
  const uint32_t bars = base + AX_BARS;
  const uint32_t q_full = bars, kp_full = bars + 8, v_full = bars + 16, s_full = bars + 24, p_full = bars + 32,
                 o_full = bars + 40;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gen + AX_BARS + 64);
  float* xch = reinterpret_cast<float*>(gen + AX_XCH);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.x, b = blockIdx.y, i0 = blockIdx.z * AX_BQ;
  const int bh = b * H + h;
  pdl_launch_dependents();
  int L = T;
  if (lengths) { L = lengths[b]; L = L < 0 ? 0 : (L > T ? T : L); }
  const int nkt = (L + AX_BKV - 1) / AX_BKV;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQu); tma_prefetch_desc(&tmQv); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
    tma_prefetch_desc(&tmP);
    mbar_init(q_full, 1); mbar_init(kp_full, 1); mbar_init(v_full, 1); mbar_init(s_full, 1);
    mbar_init(p_full, 32 * AX_SOFTMAX_WARPS); mbar_init(o_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(smem_u32(tmem_slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  pdl_wait();

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (identical to attention_f16.cu)
    if (lane == 0 && nkt > 0) {
      mbar_expect_tx(q_full, 2 * AX_BQ * 128);
      tma_load_2d(base + AX_QU, &tmQu, 0, bh * T + i0, q_full);
      tma_load_2d(base + AX_QV, &tmQv, 0, bh * T + i0, q_full);
      for (int it = 0; it < nkt; ++it) {
        const int j0 = it * AX_BKV;
        if (it > 0) mbar_wait(s_full, (it - 1) & 1);
        mbar_expect_tx(kp_full, AX_BKV * 128 + AX_BAND * 128);
        tma_load_2d(base + AX_K, &tmK, 0, bh * T + j0, kp_full);
        const int m_lo = j0 - i0 - (AX_BQ - 1) + T - 1;
        tma_load_3d(base + AX_PB, &tmP, 0, m_lo, h, kp_full);
        if (it > 0) mbar_wait(o_full, (it - 1) & 1);
        mbar_expect_tx(v_full, AX_BKV * 128);
        tma_load_2d(base + AX_V, &tmV, 0, bh * T + j0, v_full);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (identical to attention_f16.cu)
    if (lane == 0 && nkt > 0) {
      constexpr uint32_t idesc_s = umma_idesc_f16(AX_BQ, AX_BKV);
      constexpr uint32_t idesc_g = umma_idesc_f16(AX_BQ, AX_BAND);
      constexpr uint32_t idesc_o = umma_idesc_f16_bmn(AX_BQ, 64);
      auto issue_scores = [&]() {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
          mma_f16(tmem + TX_S, umma_desc_sw128(base + AX_QU + ks * 32), umma_desc_sw128(base + AX_K + ks * 32),
                  idesc_s, ks != 0);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
          mma_f16(tmem + TX_G, umma_desc_sw128(base + AX_QV + ks * 32), umma_desc_sw128(base + AX_PB + ks * 32),
                  idesc_g, ks != 0);
        tc_commit(s_full);
      };
      mbar_wait(q_full, 0);
      mbar_wait(kp_full, 0);
      tc_fence_after();
      issue_scores();
      for (int it = 0; it < nkt; ++it) {
        mbar_wait(p_full, it & 1);
        mbar_wait(v_full, it & 1);
        tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)
          mma_f16(tmem + TX_O, umma_desc_sw128(base + AX_P + (ks >> 2) * (AX_BQ * 128) + (ks & 3) * 32),
                  umma_desc_sw128(base + AX_V + ks * (16 * 128)), idesc_o, ks != 0);
        tc_commit(o_full);
        if (it + 1 < nkt) {
          mbar_wait(kp_full, (it + 1) & 1);
          tc_fence_after();
          issue_scores();
        }
      }
    }
  } else {
    // ------------------------------------------------------------ softmax + epilogue: FOUR threads per query row
    const int q = warp & 3;                       // TMEM lane quarter this warp may read (= warp id mod 4)
    const int part = (warp - 2) >> 2;             // 0..3: which 32 keys of the tile / which 16 output channels
    const int r = q * 32 + lane;                  // row inside the query tile
    const int i = i0 + r;
    const int kc = part * 32;                     // first key column of this thread's chunk inside the tile
    const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16);
    const int gbase = 96 - 32 * q;                // warp-uniform part of the skew 127 - r = gbase + (31 - lane)
    const int sh = 31 - lane;
    const float kScale = 0.125f * 1.4426950408889634f;   // 1/sqrt(64) * log2(e)
    float o[16];
#pragma unroll
    for (int d = 0; d < 16; ++d) o[d] = 0.f;
    float m_run = -INFINITY, l_run = 0.f;
    // P row chunk: atom = part / 2, 16-byte chunks (part & 1) * 4 + [0, 4) of the 128-byte atom row
    uint8_t* prow = gen + AX_P + (part >> 1) * (AX_BQ * 128) + r * 128;
    const int pch = (part & 1) * 4;
    if (i0 + q * 32 >= T) {
      for (int it = 0; it < nkt; ++it) {          // rows beyond T: keep the barrier protocol going only
        mbar_arrive(p_full);
        mbar_wait(o_full, it & 1);
      }
    } else {
      for (int it = 0; it < nkt; ++it) {
        const int jc = it * AX_BKV + kc;          // first key of this thread's chunk (warp-uniform)
        mbar_wait(s_full, it & 1);
        tc_fence_after();
        float s[32];
        if (jc >= L) {
#pragma unroll
          for (int c = 0; c < 32; ++c) s[c] = -INFINITY;
        } else {
          {
            float x[64];
            tmem_ld32(trow + TX_G + gbase + kc, x);
            tmem_ld32(trow + TX_G + gbase + kc + 32, x + 32);
            tmem_ld_wait();
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
            tmem_ld32(trow + TX_S + kc, s);       // only now: the window's upper half is dead
            tmem_ld_wait();
            if (jc + 32 <= L) {
#pragma unroll
              for (int c = 0; c < 32; ++c) s[c] += x[c];
            } else {
#pragma unroll
              for (int c = 0; c < 32; ++c) s[c] = (jc + c < L) ? s[c] + x[c] : -INFINITY;
            }
          }
        }
        float mloc = -INFINITY;
#pragma unroll
        for (int c = 0; c < 32; ++c) mloc = fmaxf(mloc, s[c]);
        // row max over the four parts: shared-memory slots (it & 1) + a 128-thread named barrier per lane quarter
        float* slot = xch + (it & 1) * 512;
        slot[part * 128 + r] = mloc;
        named_bar_sync_x(1 + q, 128);
        float mx = m_run;
#pragma unroll
        for (int p = 0; p < 4; ++p) mx = fmaxf(mx, slot[p * 128 + r]);   // finite: the tile's first key is < L
        const float alpha = ex2_neg_x((m_run - mx) * kScale);
        m_run = mx;
        const float mxs = mx * kScale;
        float sum = 0.f;
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          uint32_t pk[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float p0 = ex2_neg_x(fmaf(s[8 * ch + 2 * e], kScale, -mxs));
            const float p1 = ex2_neg_x(fmaf(s[8 * ch + 2 * e + 1], kScale, -mxs));
            sum += p0 + p1;
            const __half2 hp = __floats2half2_rn(p0, p1);
            pk[e] = *reinterpret_cast<const uint32_t*>(&hp);
          }
          *reinterpret_cast<uint4*>(prow + (((pch + ch) ^ (r & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
        l_run = l_run * alpha + sum;
        fence_proxy_async();
        tc_fence_before();
        mbar_arrive(p_full);
        mbar_wait(o_full, it & 1);
        tc_fence_after();
        {
          float pv[16];
          tmem_ld16(trow + TX_O + part * 16, pv);
          tmem_ld_wait();
#pragma unroll
          for (int d = 0; d < 16; ++d) o[d] = fmaf(o[d], alpha, pv[d]);
        }
      }
      // total row sum = the four parts' sums (all four threads hold the same running max); fixed order
      float* slot = xch + (nkt & 1) * 512;
      slot[part * 128 + r] = l_run;
      named_bar_sync_x(1 + q, 128);
      const float l_tot = (slot[r] + slot[128 + r]) + (slot[256 + r] + slot[384 + r]);
      if (i < T) {
        const float inv = l_tot > 0.f ? 1.0f / l_tot : 0.f;
        __half* dst = ctx + ((long)b * T + i) * (H * kHeadDim) + h * kHeadDim + part * 16;
#pragma unroll
        for (int d = 0; d < 16; d += 4) store_op4<__half>(dst + d, o[d] * inv, o[d + 1] * inv, o[d + 2] * inv, o[d + 3] * inv);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem);
  }
}

int attention_f16x(const __half* qu, const __half* qv, const __half* kk, const __half* vv, const __half* pos,
                   const int32_t* lengths, __half* ctx, int B, int T, int H, int Rp, cudaStream_t st) {
  AVSR_REQUIRE(Rp >= 2 * T - 1, "attention_f16x: bad Rp=%d for T=%d", Rp, T);
  if (B <= 0 || T <= 0) return AVSR_OK;
  CUtensorMap tmQu, tmQv, tmK, tmV, tmP;
  const uint64_t rows = (uint64_t)B * H * T;
  AVSR_TRY(make_tmap_2d(&tmQu, qu, rows, 64, 64, AX_BQ, 2));
  AVSR_TRY(make_tmap_2d(&tmQv, qv, rows, 64, 64, AX_BQ, 2));
  AVSR_TRY(make_tmap_2d(&tmK, kk, rows, 64, 64, AX_BKV, 2));
  AVSR_TRY(make_tmap_2d(&tmV, vv, rows, 64, 64, AX_BKV, 2));
  AVSR_TRY(make_tmap_3d(&tmP, pos, (uint64_t)H, (uint64_t)Rp, 64, 64, (uint64_t)Rp * 64, AX_BAND, 2));
  AVSR_SET_MAX_SMEM(attention_f16x_kernel, AX_SMEM);
  dim3 grid(H, B, cdiv(T, AX_BQ));
  AVSR_LAUNCH(attention_f16x_kernel, grid, AX_THREADS, AX_SMEM, st, tmQu, tmQv, tmK, tmV, tmP, lengths, ctx, T, H);
  return AVSR_OK;
}

}  // namespace avsr
