// Shared internals of libavsr_b200: error plumbing, launch accounting, the GEMM epilogue family.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>
#include <string>

#include "../../include/avsr_b200.h"

namespace avsr {

// ---------------------------------------------------------------- errors / accounting
void set_error(const char* fmt, ...);
extern std::atomic<uint64_t> g_launches;

#define AVSR_CUDA_TRY(expr)                                                                   \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess) {                                                                  \
      ::avsr::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return AVSR_E_CUDA;                                                                     \
    }                                                                                         \
  } while (0)

#define AVSR_CHECK_LAUNCH()                      \
  do {                                           \
    ::avsr::g_launches.fetch_add(1);             \
    AVSR_CUDA_TRY(cudaPeekAtLastError());        \
  } while (0)

#define AVSR_REQUIRE(cond, ...)                  \
  do {                                           \
    if (!(cond)) {                               \
      ::avsr::set_error(__VA_ARGS__);            \
      return AVSR_E_INVALID;                     \
    }                                            \
  } while (0)

#define AVSR_TRY(expr)            \
  do {                            \
    int _r = (expr);              \
    if (_r != AVSR_OK) return _r; \
  } while (0)

constexpr int kHeadDim = 64;  // d_k of the reference encoder (768 / 12); the attention kernels are built for it

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
static inline int cdiv(int a, int b) { return (a + b - 1) / b; }

// ---------------------------------------------------------------- device helpers
__device__ __forceinline__ float round_tf32(float x) {
  // round-to-nearest (ties away) to the 10-bit TF32 mantissa; the tensor core then truncates nothing.
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
__device__ __forceinline__ float sigmoidf_acc(float x) { return 1.0f / (1.0f + expf(-x)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---------------------------------------------------------------- GEMM epilogues
// Every GEMM computes acc[m][n] = sum_k A[m][k] * Bw[n][k]  (both operands K-major, i.e. torch Linear).
enum EpiMode : int {
  EPI_LINEAR = 0,  // y = [resid + alpha *] act(acc + bias[n])                      -> out (M, ldo)
  EPI_QK = 1,      // n <  D: q -> qu = q+bias+pos_u, qv = q+bias+pos_v ; n >= D: k   -> (B,H,T,64) each
  EPI_VT = 2,      // A = W_v (m = feature), B = frames (n): v^T                     -> (B,H,64,Tp)
  EPI_GLU = 3,     // interleaved pointwise_cov1: value cols [g*128, +64), gate cols +64 -> out (M, N/2)
  EPI_POS = 4      // linear_pos of all layers: n = l*D + h*64 + d, m = table row      -> (L,H,Rp,64)
};

struct EpiParams {
  int M, N;           // logical extents of acc
  const float* bias;  // LINEAR/QK/GLU: [N]; VT: [M]; may be null (LINEAR, POS)
  float* out;         // LINEAR / GLU / POS destination
  long ldo;           // LINEAR / GLU row stride of out
  const float* resid; // LINEAR: optional residual (same layout as out; may alias out)
  float alpha;        // LINEAR: scale of the residual branch
  int relu;           // LINEAR
  int round_out;      // round the stored value to TF32 (it feeds a tensor-core operand)
  int T, H, Tp, Rp;   // frames per utterance, heads, padded T of v^T, padded rows of the pos table
  const float* pos_u; // QK: [H*64]
  const float* pos_v;
  float* qu;
  float* qv;
  float* kk;
  float* vt;
};

template <int MODE>
__device__ __forceinline__ void epi_store(const EpiParams& p, int m, int n, float acc) {
  if (m >= p.M || n >= p.N) return;
  if constexpr (MODE == EPI_LINEAR) {
    float v = acc + (p.bias ? p.bias[n] : 0.0f);
    if (p.relu) v = fmaxf(v, 0.0f);
    if (p.resid) v = p.resid[(long)m * p.ldo + n] + p.alpha * v;
    if (p.round_out) v = round_tf32(v);
    p.out[(long)m * p.ldo + n] = v;
  } else if constexpr (MODE == EPI_QK) {
    const int D = p.H * kHeadDim;
    const int b = m / p.T, t = m - b * p.T;
    const int nn = n < D ? n : n - D;
    const int h = nn / kHeadDim, d = nn - h * kHeadDim;
    const long idx = (((long)b * p.H + h) * p.T + t) * kHeadDim + d;
    const float v = acc + p.bias[n];
    if (n < D) {
      float a = v + p.pos_u[nn], c = v + p.pos_v[nn];
      if (p.round_out) { a = round_tf32(a); c = round_tf32(c); }
      p.qu[idx] = a;
      p.qv[idx] = c;
    } else {
      p.kk[idx] = p.round_out ? round_tf32(v) : v;
    }
  } else if constexpr (MODE == EPI_VT) {
    const int b = n / p.T, t = n - b * p.T;
    const int h = m / kHeadDim, d = m - h * kHeadDim;
    float v = acc + p.bias[m];
    if (p.round_out) v = round_tf32(v);
    p.vt[(((long)b * p.H + h) * kHeadDim + d) * p.Tp + t] = v;
  } else if constexpr (MODE == EPI_POS) {
    const int D = p.H * kHeadDim;
    const int l = n / D, r = n - l * D;
    const int h = r / kHeadDim, d = r - h * kHeadDim;
    float v = acc;
    if (p.round_out) v = round_tf32(v);
    p.out[(((long)l * p.H + h) * p.Rp + m) * kHeadDim + d] = v;
  }
}

// GLU pair: n_val is the interleaved column of the value half, the gate sits 64 columns further.
__device__ __forceinline__ void epi_store_glu(const EpiParams& p, int m, int n_val, float acc_val, float acc_gate) {
  if (m >= p.M || n_val + 64 >= p.N) return;
  const float a = acc_val + p.bias[n_val];
  const float g = acc_gate + p.bias[n_val + 64];
  const int c = (n_val >> 7) * 64 + (n_val & 63);
  p.out[(long)m * p.ldo + c] = a * sigmoidf_acc(g);
}

// ---------------------------------------------------------------- kernel launchers (defined in the .cu files)
// fp32 CUDA-core GEMM (gemm_simt.cu)
int gemm_simt(int mode, const float* A, const float* Bw, int M, int N, int K, const EpiParams& ep, cudaStream_t st);
// tcgen05 TF32 GEMM (gemm_tc.cu)
int gemm_tc(int mode, const float* A, const float* Bw, int M, int N, int K, const EpiParams& ep, cudaStream_t st);

// elementwise.cu
int launch_embed_scale(const float* xs, float* x, long n, float scale, cudaStream_t st);
int launch_layernorm(const float* x, const float* g, const float* b, float* y, int rows, int d, int round_out,
                     cudaStream_t st);
int launch_sinusoid(float* pe, int T, int d, int round_out, cudaStream_t st);
int launch_dwconv_bn_silu(const float* x, const float* wt /*(K,C)*/, const float* scale, const float* shift, float* y,
                          int B, int T, int C, int K, int round_out, cudaStream_t st);

// attention: q-side tensors (B,H,T,64), vt (B,H,64,Tp), pos (H,Rp,64) for one layer, ctx (B*T, H*64)
int attention_simt(const float* qu, const float* qv, const float* kk, const float* vt, const float* pos,
                   const int32_t* lengths, float* ctx, int B, int T, int H, int Tp, int Rp, int round_out,
                   cudaStream_t st);
int attention_tc(const float* qu, const float* qv, const float* kk, const float* vt, const float* pos,
                 const int32_t* lengths, float* ctx, int B, int T, int H, int Tp, int Rp, int round_out,
                 cudaStream_t st);

}  // namespace avsr
