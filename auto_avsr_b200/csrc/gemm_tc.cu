// Tensor-core GEMM for the Conformer layer's seven projections (FFN x2, QK, V^T, out, pw1+GLU, pw2) and the
// stacked linear_pos:   acc[m][n] = sum_k A[m][k] * Bw[n][k]   (both K-major = torch.nn.Linear layout).
//
// Blackwell-native pipeline (sm_100a), one 128 x BN output tile per CTA, two CTAs resident per SM:
//   warp 0  : TMA producer   -- cp.async.bulk.tensor 2-D boxes (128 x 32 fp32 = 128 B rows, SWIZZLE_128B) of A and
//                               Bw into a STAGES-deep shared-memory ring, completion on `full` mbarriers
//   warp 1  : MMA issuer     -- one thread issues tcgen05.mma.kind::tf32 (M=128, N=BN, K=8) x4 per stage straight
//                               from shared memory into a TMEM accumulator; tcgen05.commit releases the stage
//                               (`empty` mbarrier) and finally signals `tmem_full`
//   warps 2-5: epilogue      -- tcgen05.ld 32 lanes x 32 columns, fused epilogue (bias / ReLU / residual / GLU /
//                               head-major scatter / TF32 rounding of the next operand), direct global stores
// Out-of-bounds rows of the last M / N tile are zero-filled by TMA and masked in the epilogue.
// Operands are fp32 in memory, pre-rounded to TF32 by their producers (avsr_prepare_weights, the LayerNorm /
// epilogue that wrote them), so the tensor core's mantissa truncation is exact.
#include "common.cuh"
#include "sm100.cuh"

namespace avsr {

using namespace sm100;

// ---------------------------------------------------------------- host: tensor maps
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

int make_tmap_2d(CUtensorMap* map, const float* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows) {
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) { set_error("cuTensorMapEncodeTiled unavailable (no CUDA driver?)"); return AVSR_E_CUDA; }
  AVSR_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0 && (ld * 4) % 16 == 0 && box_rows <= 256,
               "tensor map: base/stride must be 16-byte aligned (ld=%llu)", (unsigned long long)ld);
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstr[1] = {ld * sizeof(float)};
  cuuint32_t box[2] = {32, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(2d) failed: CUresult %d", (int)r); return AVSR_E_CUDA; }
  return AVSR_OK;
}

int make_tmap_3d(CUtensorMap* map, const float* base, uint64_t planes, uint64_t rows, uint64_t cols, uint64_t ld_row,
                 uint64_t ld_plane, uint32_t box_rows) {
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) { set_error("cuTensorMapEncodeTiled unavailable (no CUDA driver?)"); return AVSR_E_CUDA; }
  AVSR_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0 && (ld_row * 4) % 16 == 0 && (ld_plane * 4) % 16 == 0 &&
                   box_rows <= 256,
               "tensor map 3d: base/strides must be 16-byte aligned");
  cuuint64_t gdim[3] = {cols, rows, planes};
  cuuint64_t gstr[2] = {ld_row * sizeof(float), ld_plane * sizeof(float)};
  cuuint32_t box[3] = {32, box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(3d) failed: CUresult %d", (int)r); return AVSR_E_CUDA; }
  return AVSR_OK;
}

// ---------------------------------------------------------------- vectorised epilogues: one row, 32 columns
// `v` holds acc[m][n .. n+31]; n is a multiple of 32 inside the tile.
template <int MODE>
__device__ __forceinline__ void epi_chunk32(const EpiParams& p, int m, int n, const float* v) {
  if (m >= p.M || n >= p.N) return;
  if constexpr (MODE == EPI_LINEAR) {
    if (n + 32 <= p.N && (p.ldo & 3) == 0) {
      float* dst = p.out + (long)m * p.ldo + n;
      const float* res = p.resid ? p.resid + (long)m * p.ldo + n : nullptr;
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        float4 o = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        if (p.bias) {
          const float4 b = *reinterpret_cast<const float4*>(p.bias + n + j);
          o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
        }
        if (p.relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
        if (res) {
          const float4 r = *reinterpret_cast<const float4*>(res + j);
          o.x = r.x + p.alpha * o.x; o.y = r.y + p.alpha * o.y; o.z = r.z + p.alpha * o.z; o.w = r.w + p.alpha * o.w;
        }
        if (p.round_out) { o.x = round_tf32(o.x); o.y = round_tf32(o.y); o.z = round_tf32(o.z); o.w = round_tf32(o.w); }
        *reinterpret_cast<float4*>(dst + j) = o;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) epi_store<EPI_LINEAR>(p, m, n + j, v[j]);
    }
  } else if constexpr (MODE == EPI_QK) {
    const int D = p.H * kHeadDim;                 // multiple of 64, so a 32-chunk never straddles q|k or a head
    const int b = m / p.T, t = m - b * p.T;
    const int nn = n < D ? n : n - D;
    const int h = nn / kHeadDim, d0 = nn - h * kHeadDim;
    const long idx = (((long)b * p.H + h) * p.T + t) * kHeadDim + d0;
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      const float4 bq = *reinterpret_cast<const float4*>(p.bias + n + j);
      float4 o = make_float4(v[j] + bq.x, v[j + 1] + bq.y, v[j + 2] + bq.z, v[j + 3] + bq.w);
      if (n < D) {
        const float4 u = *reinterpret_cast<const float4*>(p.pos_u + nn + j);
        const float4 w = *reinterpret_cast<const float4*>(p.pos_v + nn + j);
        float4 a = make_float4(o.x + u.x, o.y + u.y, o.z + u.z, o.w + u.w);
        float4 c = make_float4(o.x + w.x, o.y + w.y, o.z + w.z, o.w + w.w);
        if (p.round_out) {
          a.x = round_tf32(a.x); a.y = round_tf32(a.y); a.z = round_tf32(a.z); a.w = round_tf32(a.w);
          c.x = round_tf32(c.x); c.y = round_tf32(c.y); c.z = round_tf32(c.z); c.w = round_tf32(c.w);
        }
        *reinterpret_cast<float4*>(p.qu + idx + j) = a;
        *reinterpret_cast<float4*>(p.qv + idx + j) = c;
      } else {
        if (p.round_out) { o.x = round_tf32(o.x); o.y = round_tf32(o.y); o.z = round_tf32(o.z); o.w = round_tf32(o.w); }
        *reinterpret_cast<float4*>(p.kk + idx + j) = o;
      }
    }
  } else if constexpr (MODE == EPI_VT) {
    // m = v feature (h*64+d), n.. = 32 consecutive frames: contiguous in v^T unless an utterance boundary intervenes
    const int h = m / kHeadDim, d = m - h * kHeadDim;
    const float bias = p.bias[m];
    const int b0 = n / p.T, t0 = n - b0 * p.T;
    if (t0 + 32 <= p.T && n + 32 <= p.N && (t0 & 3) == 0) {
      float* dst = p.vt + (((long)b0 * p.H + h) * kHeadDim + d) * p.Tp + t0;
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        float4 o = make_float4(v[j] + bias, v[j + 1] + bias, v[j + 2] + bias, v[j + 3] + bias);
        if (p.round_out) { o.x = round_tf32(o.x); o.y = round_tf32(o.y); o.z = round_tf32(o.z); o.w = round_tf32(o.w); }
        *reinterpret_cast<float4*>(dst + j) = o;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) epi_store<EPI_VT>(p, m, n + j, v[j]);
    }
  } else if constexpr (MODE == EPI_POS) {
    const int D = p.H * kHeadDim;
    const int l = n / D, r = n - l * D;
    const int h = r / kHeadDim, d0 = r - h * kHeadDim;
    float* dst = p.out + (((long)l * p.H + h) * p.Rp + m) * kHeadDim + d0;
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      float4 o = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
      if (p.round_out) { o.x = round_tf32(o.x); o.y = round_tf32(o.y); o.z = round_tf32(o.z); o.w = round_tf32(o.w); }
      *reinterpret_cast<float4*>(dst + j) = o;
    }
  }
}

// GLU: value columns n .. n+31 (n % 128 < 64), gates 64 columns further; out column = (n/128)*64 + n%128
__device__ __forceinline__ void epi_chunk32_glu(const EpiParams& p, int m, int n, const float* val, const float* gate) {
  if (m >= p.M || n + 64 >= p.N) return;
  const int c0 = (n >> 7) * 64 + (n & 127);
  float* dst = p.out + (long)m * p.ldo + c0;
#pragma unroll
  for (int j = 0; j < 32; j += 4) {
    const float4 ba = *reinterpret_cast<const float4*>(p.bias + n + j);
    const float4 bg = *reinterpret_cast<const float4*>(p.bias + n + 64 + j);
    float4 o;
    o.x = (val[j] + ba.x) * sigmoidf_acc(gate[j] + bg.x);
    o.y = (val[j + 1] + ba.y) * sigmoidf_acc(gate[j + 1] + bg.y);
    o.z = (val[j + 2] + ba.z) * sigmoidf_acc(gate[j + 2] + bg.z);
    o.w = (val[j + 3] + ba.w) * sigmoidf_acc(gate[j + 3] + bg.w);
    *reinterpret_cast<float4*>(dst + j) = o;
  }
}

// ---------------------------------------------------------------- the kernel
constexpr int TC_BM = 128;
constexpr int TC_BK = 32;  // fp32 elements = one 128-byte swizzle row
constexpr int TC_THREADS = 192;

template <int BN>
struct TcCfg {
  static constexpr int kStages = BN == 128 ? 3 : 4;
  static constexpr int kABytes = TC_BM * TC_BK * 4;
  static constexpr int kBBytes = BN * TC_BK * 4;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kSmem = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
};

template <int MODE, int BN>
__global__ void __launch_bounds__(TC_THREADS, 2)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int K, EpiParams ep) {
  using Cfg = TcCfg<BN>;
  constexpr int S = Cfg::kStages;
  extern __shared__ uint8_t tc_smem_raw[];
  const uint32_t raw = smem_u32(tc_smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;       // SWIZZLE_128B tiles need 1024-byte alignment
  uint8_t* gen = tc_smem_raw + (base - raw);
  const uint32_t bars = base + S * Cfg::kStageBytes;  // full[S], empty[S], tmem_full
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gen + S * Cfg::kStageBytes + (2 * S + 1) * 8);
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (S + s); };
  const uint32_t tmem_full_bar = bars + 8u * (2 * S);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * TC_BM, n0 = blockIdx.x * BN;
  const int nkb = K / TC_BK;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < S; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<BN>(smem_u32(tmem_slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % S;
        const uint32_t ph = (kb / S) & 1;
        mbar_wait(empty_bar(s), ph ^ 1);
        mbar_expect_tx(full_bar(s), Cfg::kStageBytes);
        const uint32_t a_dst = base + s * Cfg::kStageBytes;
        tma_load_2d(a_dst, &tmA, kb * TC_BK, m0, full_bar(s));
        tma_load_2d(a_dst + Cfg::kABytes, &tmB, kb * TC_BK, n0, full_bar(s));
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_tf32(TC_BM, BN);
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % S;
        const uint32_t ph = (kb / S) & 1;
        mbar_wait(full_bar(s), ph);
        tc_fence_after();
        const uint32_t a_addr = base + s * Cfg::kStageBytes;
        const uint64_t a_desc = umma_desc_sw128(a_addr);
        const uint64_t b_desc = umma_desc_sw128(a_addr + Cfg::kABytes);
#pragma unroll
        for (int k = 0; k < TC_BK / 8; ++k)  // 8 tf32 = 32 bytes = +2 in the descriptor's 16-byte address field
          mma_tf32(tmem_base, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0);
        tc_commit(empty_bar(s));
      }
      tc_commit(tmem_full_bar);
    }
  } else {
    const int q = warp & 3;                      // TMEM lane quarter this warp may read
    const int m = m0 + q * 32 + lane;
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
    if constexpr (MODE == EPI_GLU) {
      static_assert(MODE != EPI_GLU || BN == 128, "GLU pairs live 64 columns apart inside a 128-wide tile");
#pragma unroll 1
      for (int c = 0; c < 64; c += 32) {
        float val[32], gate[32];
        tmem_ld32(trow + c, val);
        tmem_ld32(trow + 64 + c, gate);
        tmem_ld_wait();
        epi_chunk32_glu(ep, m, n0 + c, val, gate);
      }
    } else {
#pragma unroll 1
      for (int c = 0; c < BN; c += 32) {
        float v[32];
        tmem_ld32(trow + c, v);
        tmem_ld_wait();
        epi_chunk32<MODE>(ep, m, n0 + c, v);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<BN>(tmem_base);
  }
}

template <int MODE, int BN>
static int launch_tc(const float* A, const float* Bw, int M, int N, int K, const EpiParams& ep, cudaStream_t st) {
  using Cfg = TcCfg<BN>;
  CUtensorMap tmA, tmB;
  AVSR_TRY(make_tmap_2d(&tmA, A, (uint64_t)M, (uint64_t)K, (uint64_t)K, TC_BM));
  AVSR_TRY(make_tmap_2d(&tmB, Bw, (uint64_t)N, (uint64_t)K, (uint64_t)K, BN));
  static bool attr_done = false;  // per instantiation
  if (!attr_done) {
    AVSR_CUDA_TRY(cudaFuncSetAttribute(gemm_tc_kernel<MODE, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmem));
    attr_done = true;
  }
  dim3 grid(cdiv(N, BN), cdiv(M, TC_BM));
  gemm_tc_kernel<MODE, BN><<<grid, TC_THREADS, Cfg::kSmem, st>>>(tmA, tmB, K, ep);
  AVSR_CHECK_LAUNCH();
  return AVSR_OK;
}

int gemm_tc(int mode, const float* A, const float* Bw, int M, int N, int K, const EpiParams& ep, cudaStream_t st) {
  AVSR_REQUIRE(K >= TC_BK && K % TC_BK == 0, "gemm_tc: K=%d must be a multiple of %d", K, TC_BK);
  if (M <= 0 || N <= 0) return AVSR_OK;
  // 128-wide tiles when they already fill the 148 SMs, otherwise 64-wide for more CTAs
  const bool wide = (long)cdiv(M, TC_BM) * cdiv(N, 128) >= 148;
  switch (mode) {
    case EPI_LINEAR:
      return wide ? launch_tc<EPI_LINEAR, 128>(A, Bw, M, N, K, ep, st) : launch_tc<EPI_LINEAR, 64>(A, Bw, M, N, K, ep, st);
    case EPI_QK:
      AVSR_REQUIRE(N % 128 == 0, "gemm_tc: QK needs N %% 128 == 0 (N=%d)", N);
      return wide ? launch_tc<EPI_QK, 128>(A, Bw, M, N, K, ep, st) : launch_tc<EPI_QK, 64>(A, Bw, M, N, K, ep, st);
    case EPI_VT:
      AVSR_REQUIRE(M % 64 == 0, "gemm_tc: V^T needs M %% 64 == 0 (M=%d)", M);
      return wide ? launch_tc<EPI_VT, 128>(A, Bw, M, N, K, ep, st) : launch_tc<EPI_VT, 64>(A, Bw, M, N, K, ep, st);
    case EPI_GLU:
      AVSR_REQUIRE(N % 128 == 0, "gemm_tc: GLU needs N %% 128 == 0 (N=%d)", N);
      return launch_tc<EPI_GLU, 128>(A, Bw, M, N, K, ep, st);
    case EPI_POS:
      AVSR_REQUIRE(N % 64 == 0, "gemm_tc: POS needs N %% 64 == 0 (N=%d)", N);
      return wide ? launch_tc<EPI_POS, 128>(A, Bw, M, N, K, ep, st) : launch_tc<EPI_POS, 64>(A, Bw, M, N, K, ep, st);
    default:
      AVSR_REQUIRE(false, "gemm_tc: bad epilogue mode %d", mode);
  }
  return AVSR_OK;
}

}  // namespace avsr
