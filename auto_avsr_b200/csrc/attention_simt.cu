// fp32 CUDA-core rel-pos attention -- the on-device EXACT reference path (AVSR_PREC_FP32).
//   scores[i,j] = (qu_i . k_j + qv_i . p[j-i+T-1]) / 8,  keys j >= len masked, softmax, ctx = attn @ v
// restating transformer/attention.py:174-189 (rel_shift folded into the index j-i+T-1) and
// :72-82.  One warp per query row; scores of the row live in shared memory; warp-shuffle max / sum.
// Padded QUERY rows are computed like any other (only keys are masked).  Not the product path.
#include "common.cuh"

namespace avsr {

constexpr int kAsWarps = 4;

__global__ void __launch_bounds__(kAsWarps * 32) attention_simt_kernel(
    const float* __restrict__ qu, const float* __restrict__ qv, const float* __restrict__ kk,
    const float* __restrict__ vt, const float* __restrict__ pos, const int32_t* __restrict__ lengths,
    float* __restrict__ ctx, int T, int H, int Tp, int Rp, int round_out) {
  extern __shared__ float as_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i = blockIdx.x * kAsWarps + warp;
  const int h = blockIdx.y, b = blockIdx.z;
  if (i >= T) return;  // warp-uniform; only __syncwarp below
  const int T4 = (T + 3) & ~3;
  float* qu_s = as_smem + warp * (2 * kHeadDim + T4);
  float* qv_s = qu_s + kHeadDim;
  float* sc = qv_s + kHeadDim;

  int L = T;
  if (lengths) { L = lengths[b]; L = L < 0 ? 0 : (L > T ? T : L); }
  const long bh = (long)b * H + h;
  const float* qur = qu + (bh * T + i) * kHeadDim;
  const float* qvr = qv + (bh * T + i) * kHeadDim;
  qu_s[lane] = qur[lane]; qu_s[lane + 32] = qur[lane + 32];
  qv_s[lane] = qvr[lane]; qv_s[lane + 32] = qvr[lane + 32];
  __syncwarp();

  float mx = -INFINITY;
  for (int j = lane; j < L; j += 32) {
    const float4* kr = reinterpret_cast<const float4*>(kk + (bh * T + j) * kHeadDim);
    const float4* pr = reinterpret_cast<const float4*>(pos + ((long)h * Rp + (j - i + T - 1)) * kHeadDim);
    float ac = 0.f, bd = 0.f;
#pragma unroll
    for (int d = 0; d < kHeadDim / 4; ++d) {
      const float4 kv = kr[d], pv = pr[d];
      const float4 a = *reinterpret_cast<const float4*>(qu_s + 4 * d);
      const float4 c = *reinterpret_cast<const float4*>(qv_s + 4 * d);
      ac = fmaf(a.x, kv.x, ac); ac = fmaf(a.y, kv.y, ac); ac = fmaf(a.z, kv.z, ac); ac = fmaf(a.w, kv.w, ac);
      bd = fmaf(c.x, pv.x, bd); bd = fmaf(c.y, pv.y, bd); bd = fmaf(c.z, pv.z, bd); bd = fmaf(c.w, pv.w, bd);
    }
    const float s = (ac + bd) * 0.125f;  // 1/sqrt(64)
    sc[j] = s;
    mx = fmaxf(mx, s);
  }
  mx = warp_max(mx);
  float sum = 0.f;
  for (int j = lane; j < T4; j += 32) {
    float e = 0.f;
    if (j < L) e = expf(sc[j] - mx);
    sc[j] = e;
    sum += e;
  }
  sum = warp_sum(sum);
  __syncwarp();
  // all keys masked -> zeros (softmax(...).masked_fill(mask, 0), attention.py:75-77)
  const float inv = (L > 0) ? 1.0f / sum : 0.f;

  float o[2] = {0.f, 0.f};
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const float* vr = vt + (bh * kHeadDim + lane + 32 * r) * Tp;
    float acc = 0.f;
    const int L4 = (L + 3) & ~3;
    for (int j = 0; j < L4; j += 4) {
      const float4 pj = *reinterpret_cast<const float4*>(sc + j);
      const float4 vv = *reinterpret_cast<const float4*>(vr + j);
      acc = fmaf(pj.x, vv.x, acc); acc = fmaf(pj.y, vv.y, acc); acc = fmaf(pj.z, vv.z, acc); acc = fmaf(pj.w, vv.w, acc);
    }
    o[r] = acc * inv;
    if (round_out) o[r] = round_tf32(o[r]);
  }
  float* out = ctx + ((long)b * T + i) * (H * kHeadDim) + h * kHeadDim;
  out[lane] = o[0];
  out[lane + 32] = o[1];
}

int attention_simt(const float* qu, const float* qv, const float* kk, const float* vt, const float* pos,
                   const int32_t* lengths, float* ctx, int B, int T, int H, int Tp, int Rp, int round_out,
                   cudaStream_t st) {
  AVSR_REQUIRE(Tp % 4 == 0 && Tp >= T, "attention_simt: Tp=%d must be a multiple of 4 and >= T=%d", Tp, T);
  if (B <= 0 || T <= 0) return AVSR_OK;
  const int T4 = (T + 3) & ~3;
  const size_t smem = (size_t)kAsWarps * (2 * kHeadDim + T4) * sizeof(float);
  AVSR_REQUIRE(smem <= 200 * 1024, "attention_simt: T=%d too long for the fp32 reference kernel", T);
  if (smem > 48 * 1024)
    AVSR_CUDA_TRY(cudaFuncSetAttribute(attention_simt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(cdiv(T, kAsWarps), H, B);
  attention_simt_kernel<<<grid, kAsWarps * 32, smem, st>>>(qu, qv, kk, vt, pos, lengths, ctx, T, H, Tp, Rp, round_out);
  AVSR_CHECK_LAUNCH();
  return AVSR_OK;
}

}  // namespace avsr
