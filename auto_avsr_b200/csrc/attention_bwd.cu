// Backward of the rel-pos attention core: fp32 CUDA-core kernels, correctness first -- the training
// counterpart of attention_simt.cu.  Forward (transformer/attention.py:174-189 + :59-82):
//   s_ij = ((q_i + u).k_j + (q_i + v).p[m_ij]) / 8,  m_ij = j - i + T - 1 ;  keys j >= len[b] masked ;
//   A = softmax_j(s) ;  ctx_i = sum_j A_ij v_j .
// Backward, with Delta_i = dctx_i . ctx_i and g_ij = A_ij (dctx_i . v_j - Delta_i) / 8 :
//   dq_i = sum_j g_ij (k_j + p[m_ij])        (kept as two parts: their column sums are d pos_bias_u / d pos_bias_v)
//   dk_j = sum_i g_ij (q_i + u) ;  dv_j = sum_i A_ij dctx_i ;  dp[m] = sum_{b, i} g_{i, j = m + i - (T-1)} (q_i + v) .
// Nothing of size T^2 is stored: three passes recompute the scores from q, k, p --
//   rows kernel  (one warp per query row)      : log-sum-exp and Delta of the row, then the two parts of dq
//   keys kernel  (one warp per key)            : dk, dv
//   pos kernel   (one warp per (table row, h)) : dp
// A lane owns one (row, key) pair at a time: its 64-element dot products read the pair's vectors straight from global
// memory (256 contiguous bytes per lane, L1-resident across the 16 float4 steps) against the warp's own vector in
// shared memory (broadcast reads).  Every output element is written by exactly one warp: no atomics, deterministic.
// q, k, v, ctx, dctx and the outputs are (B, T, H*64); p / dp are (2T-1, H*64); u, v (H, 64).
#include <math.h>

#include "common.cuh"

namespace avsr {

constexpr int kAbWarps = 4;

__device__ __forceinline__ float dot64(const float* __restrict__ smem_vec, const float* __restrict__ row) {
  float acc = 0.f;
#pragma unroll
  for (int t = 0; t < 16; ++t) {
    const float4 a = reinterpret_cast<const float4*>(smem_vec)[t];
    const float4 b = reinterpret_cast<const float4*>(row)[t];
    acc = fmaf(a.x, b.x, acc); acc = fmaf(a.y, b.y, acc); acc = fmaf(a.z, b.z, acc); acc = fmaf(a.w, b.w, acc);
  }
  return acc;
}
// dot of two global rows where the first one gets a bias vector (shared memory) added: (a + bias) . c
__device__ __forceinline__ float dot64_biased(const float* __restrict__ a, const float* __restrict__ bias,
                                              const float* __restrict__ c) {
  float acc = 0.f;
#pragma unroll
  for (int t = 0; t < 16; ++t) {
    const float4 x = reinterpret_cast<const float4*>(a)[t];
    const float4 b = reinterpret_cast<const float4*>(bias)[t];
    const float4 y = reinterpret_cast<const float4*>(c)[t];
    acc = fmaf(x.x + b.x, y.x, acc); acc = fmaf(x.y + b.y, y.y, acc); acc = fmaf(x.z + b.z, y.z, acc); acc = fmaf(x.w + b.w, y.w, acc);
  }
  return acc;
}
// acc[0..63] += s * (row[0..63] (+ bias))
__device__ __forceinline__ void axpy64(float (&acc)[64], float s, const float* __restrict__ row, const float* bias) {
#pragma unroll
  for (int t = 0; t < 16; ++t) {
    float4 b = reinterpret_cast<const float4*>(row)[t];
    if (bias) {
      const float4 c = reinterpret_cast<const float4*>(bias)[t];
      b.x += c.x; b.y += c.y; b.z += c.z; b.w += c.w;
    }
    acc[4 * t] = fmaf(s, b.x, acc[4 * t]); acc[4 * t + 1] = fmaf(s, b.y, acc[4 * t + 1]);
    acc[4 * t + 2] = fmaf(s, b.z, acc[4 * t + 2]); acc[4 * t + 3] = fmaf(s, b.w, acc[4 * t + 3]);
  }
}
// sum acc[d] over the warp's lanes; lane (d & 31) keeps element d: returns this lane's two elements (d = lane, lane + 32)
__device__ __forceinline__ float2 reduce64(float (&acc)[64], int lane) {
  float lo = 0.f, hi = 0.f;
#pragma unroll
  for (int d = 0; d < 64; ++d) {
    const float t = warp_sum(acc[d]);
    if ((d & 31) == lane) { if (d < 32) lo = t; else hi = t; }
  }
  return make_float2(lo, hi);
}

// ---------------------------------------------------------------- rows: lse, Delta, dq (k part / p part)
__global__ void __launch_bounds__(32 * kAbWarps) attn_bwd_rows_kernel(
    const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v, const float* __restrict__ p,
    const float* __restrict__ u, const float* __restrict__ vb, const int32_t* __restrict__ lengths,
    const float* __restrict__ ctx, const float* __restrict__ dctx, float* __restrict__ lse, float* __restrict__ delta,
    float* __restrict__ dq_k, float* __restrict__ dq_p, int T, int H) {
  __shared__ __align__(16) float sm[kAbWarps][3][64];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i = blockIdx.x * kAbWarps + warp, h = blockIdx.y, b = blockIdx.z;
  if (i >= T) return;
  const int D = H * kHeadDim;
  int L = T;
  if (lengths) { L = lengths[b]; L = L < 0 ? 0 : (L > T ? T : L); }
  const long ro = ((long)b * T + i) * D + h * kHeadDim;
  float* qu = sm[warp][0];
  float* qv = sm[warp][1];
  float* dc = sm[warp][2];
  float dl = 0.f;
  for (int d = lane; d < 64; d += 32) {
    const float qq = q[ro + d];
    qu[d] = qq + u[h * 64 + d];
    qv[d] = qq + vb[h * 64 + d];
    dc[d] = dctx[ro + d];
    dl += dctx[ro + d] * ctx[ro + d];
  }
  dl = warp_sum(dl);
  __syncwarp();
  const float* kb = k + (long)b * T * D + h * kHeadDim;
  const float* vbase = v + (long)b * T * D + h * kHeadDim;
  const float* pb = p + h * kHeadDim;
  // pass 1: online log-sum-exp of the row
  float m = -INFINITY, ssum = 0.f;
  for (int j0 = 0; j0 < L; j0 += 32) {
    const int j = j0 + lane;
    float s = -INFINITY;
    if (j < L) s = 0.125f * (dot64(qu, kb + (long)j * D) + dot64(qv, pb + (long)(j - i + T - 1) * D));
    const float mn = fmaxf(m, warp_max(s));
    ssum = ssum * __expf(m - mn) + warp_sum(j < L ? __expf(s - mn) : 0.f);
    m = mn;
  }
  const float row_lse = L > 0 ? m + logf(ssum) : INFINITY;     // len 0: exp(s - inf) = 0 below
  if (lane == 0) {
    lse[((long)b * H + h) * T + i] = row_lse;
    delta[((long)b * H + h) * T + i] = dl;
  }
  // pass 2: g_ij and the two parts of dq
  float ak[64], ap[64];
#pragma unroll
  for (int d = 0; d < 64; ++d) ak[d] = ap[d] = 0.f;
  for (int j0 = 0; j0 < L; j0 += 32) {
    const int j = j0 + lane;
    if (j < L) {
      const float* kr = kb + (long)j * D;
      const float* pr = pb + (long)(j - i + T - 1) * D;
      const float s = 0.125f * (dot64(qu, kr) + dot64(qv, pr));
      const float a = __expf(s - row_lse);
      const float g = a * (dot64(dc, vbase + (long)j * D) - dl) * 0.125f;
      axpy64(ak, g, kr, nullptr);
      axpy64(ap, g, pr, nullptr);
    }
  }
  const float2 rk = reduce64(ak, lane), rp = reduce64(ap, lane);
  dq_k[ro + lane] = rk.x; dq_k[ro + lane + 32] = rk.y;
  dq_p[ro + lane] = rp.x; dq_p[ro + lane + 32] = rp.y;
}

// ---------------------------------------------------------------- keys: dk, dv
__global__ void __launch_bounds__(32 * kAbWarps) attn_bwd_keys_kernel(
    const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v, const float* __restrict__ p,
    const float* __restrict__ u, const float* __restrict__ vb, const int32_t* __restrict__ lengths,
    const float* __restrict__ dctx, const float* __restrict__ lse, const float* __restrict__ delta,
    float* __restrict__ dk, float* __restrict__ dv, int T, int H) {
  __shared__ __align__(16) float sm[kAbWarps][4][64];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j = blockIdx.x * kAbWarps + warp, h = blockIdx.y, b = blockIdx.z;
  if (j >= T) return;
  const int D = H * kHeadDim;
  int L = T;
  if (lengths) { L = lengths[b]; L = L < 0 ? 0 : (L > T ? T : L); }
  const long ko = ((long)b * T + j) * D + h * kHeadDim;
  if (j >= L) {                                   // masked key: it received no probability
    dk[ko + lane] = 0.f; dk[ko + lane + 32] = 0.f;
    dv[ko + lane] = 0.f; dv[ko + lane + 32] = 0.f;
    return;
  }
  float* kj = sm[warp][0];
  float* vj = sm[warp][1];
  float* su = sm[warp][2];
  float* sv = sm[warp][3];
  for (int d = lane; d < 64; d += 32) { kj[d] = k[ko + d]; vj[d] = v[ko + d]; su[d] = u[h * 64 + d]; sv[d] = vb[h * 64 + d]; }
  __syncwarp();
  const float* qb = q + (long)b * T * D + h * kHeadDim;
  const float* db_ = dctx + (long)b * T * D + h * kHeadDim;
  const float* pb = p + h * kHeadDim;
  const float* lrow = lse + ((long)b * H + h) * T;
  const float* drow = delta + ((long)b * H + h) * T;
  float ak[64], av[64];
#pragma unroll
  for (int d = 0; d < 64; ++d) ak[d] = av[d] = 0.f;
  for (int i0 = 0; i0 < T; i0 += 32) {
    const int i = i0 + lane;
    if (i < T) {
      const float* qr = qb + (long)i * D;
      const float* dr = db_ + (long)i * D;
      const float s = 0.125f * (dot64_biased(qr, su, kj) + dot64_biased(qr, sv, pb + (long)(j - i + T - 1) * D));
      const float a = __expf(s - lrow[i]);
      const float g = a * (dot64(vj, dr) - drow[i]) * 0.125f;
      axpy64(av, a, dr, nullptr);
      axpy64(ak, g, qr, su);
    }
  }
  const float2 rk = reduce64(ak, lane), rv = reduce64(av, lane);
  dk[ko + lane] = rk.x; dk[ko + lane + 32] = rk.y;
  dv[ko + lane] = rv.x; dv[ko + lane + 32] = rv.y;
}

// ---------------------------------------------------------------- rel-pos table rows: dp
__global__ void __launch_bounds__(32 * kAbWarps) attn_bwd_pos_kernel(
    const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v, const float* __restrict__ p,
    const float* __restrict__ u, const float* __restrict__ vb, const int32_t* __restrict__ lengths,
    const float* __restrict__ dctx, const float* __restrict__ lse, const float* __restrict__ delta,
    float* __restrict__ dp, int B, int T, int H) {
  __shared__ __align__(16) float sm[kAbWarps][3][64];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m = blockIdx.x * kAbWarps + warp, h = blockIdx.y;
  const int R = 2 * T - 1;
  if (m >= R) return;
  const int D = H * kHeadDim;
  float* pm = sm[warp][0];
  float* su = sm[warp][1];
  float* sv = sm[warp][2];
  for (int d = lane; d < 64; d += 32) { pm[d] = p[(long)m * D + h * 64 + d]; su[d] = u[h * 64 + d]; sv[d] = vb[h * 64 + d]; }
  __syncwarp();
  float ap[64];
#pragma unroll
  for (int d = 0; d < 64; ++d) ap[d] = 0.f;
  const int shift = m - (T - 1);                  // j = i + shift
  for (int b = 0; b < B; ++b) {
    int L = T;
    if (lengths) { L = lengths[b]; L = L < 0 ? 0 : (L > T ? T : L); }
    const float* qb = q + (long)b * T * D + h * kHeadDim;
    const float* kb = k + (long)b * T * D + h * kHeadDim;
    const float* vbase = v + (long)b * T * D + h * kHeadDim;
    const float* db_ = dctx + (long)b * T * D + h * kHeadDim;
    const float* lrow = lse + ((long)b * H + h) * T;
    const float* drow = delta + ((long)b * H + h) * T;
    const int ilo = shift < 0 ? -shift : 0;       // j >= 0
    int ihi = L - shift;                          // j < L
    if (ihi > T) ihi = T;
    for (int i0 = ilo; i0 < ihi; i0 += 32) {
      const int i = i0 + lane;
      if (i < ihi) {
        const int j = i + shift;
        const float* qr = qb + (long)i * D;
        const float s = 0.125f * (dot64_biased(qr, su, kb + (long)j * D) + dot64_biased(qr, sv, pm));
        const float a = __expf(s - lrow[i]);
        float da = 0.f;
        {
          const float* dr = db_ + (long)i * D;
          const float* vr = vbase + (long)j * D;
#pragma unroll
          for (int t = 0; t < 16; ++t) {
            const float4 x = reinterpret_cast<const float4*>(dr)[t];
            const float4 y = reinterpret_cast<const float4*>(vr)[t];
            da = fmaf(x.x, y.x, da); da = fmaf(x.y, y.y, da); da = fmaf(x.z, y.z, da); da = fmaf(x.w, y.w, da);
          }
        }
        const float gg = a * (da - drow[i]) * 0.125f;
        axpy64(ap, gg, qr, sv);
      }
    }
  }
  const float2 rp = reduce64(ap, lane);
  dp[(long)m * D + h * 64 + lane] = rp.x;
  dp[(long)m * D + h * 64 + lane + 32] = rp.y;
}

}  // namespace avsr

using namespace avsr;

extern "C" {

size_t avsr_relpos_attention_bwd_workspace_bytes(int B, int T, int H) {
  if (B <= 0 || T <= 0 || H <= 0) return 256;
  return align_up((size_t)2 * B * H * T * sizeof(float), 256) + 256;
}

int avsr_relpos_attention_bwd(const float* q, const float* k, const float* v, const float* p, const float* pos_bias_u,
                              const float* pos_bias_v, const int32_t* lengths, const float* ctx, const float* dctx,
                              float* dq_k, float* dq_p, float* dk, float* dv, float* dp, int B, int T, int H,
                              void* workspace, size_t workspace_bytes, void* stream) {
  AVSR_REQUIRE(q && k && v && p && pos_bias_u && pos_bias_v && ctx && dctx && dq_k && dq_p && dk && dv && dp && workspace,
               "NULL argument");
  if (B <= 0 || T <= 0) return AVSR_OK;
  if (avsr_relpos_attention_bwd_workspace_bytes(B, T, H) > workspace_bytes) {
    set_error("attention backward workspace too small");
    return AVSR_E_WORKSPACE;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  float* lse = reinterpret_cast<float*>(workspace);
  float* delta = lse + (size_t)B * H * T;
  dim3 block(32 * kAbWarps);
  attn_bwd_rows_kernel<<<dim3(cdiv(T, kAbWarps), H, B), block, 0, st>>>(q, k, v, p, pos_bias_u, pos_bias_v, lengths, ctx, dctx, lse,
                                                                         delta, dq_k, dq_p, T, H);
  AVSR_CHECK_LAUNCH();
  attn_bwd_keys_kernel<<<dim3(cdiv(T, kAbWarps), H, B), block, 0, st>>>(q, k, v, p, pos_bias_u, pos_bias_v, lengths, dctx, lse, delta,
                                                                         dk, dv, T, H);
  AVSR_CHECK_LAUNCH();
  attn_bwd_pos_kernel<<<dim3(cdiv(2 * T - 1, kAbWarps), H), block, 0, st>>>(q, k, v, p, pos_bias_u, pos_bias_v, lengths, dctx, lse,
                                                                             delta, dp, B, T, H);
  AVSR_CHECK_LAUNCH();
  return AVSR_OK;
}

}  // extern "C"
