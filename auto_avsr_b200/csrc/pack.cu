// On-device collate: the reference pads a max-frames bucket on the host (datamodule/data_module.py:10-41,
// `pad` / `collate_pad`: zero padding to the longest utterance) and ships B*Tmax frames; here the utterances of a bucket
// sit back to back in ONE flat device buffer (sum of lengths rows: what actually crosses PCIe) and the padded
// (B, Tmax, d) batch the encoder consumes is formed on the GPU -- coalesced float4 row copies, pad rows written as zeros.
#include "common.cuh"

namespace avsr {

// flat: (sum_len, d) rows of utterance b at offsets[b] .. offsets[b+1]; out: (B, Tmax, d); lengths_out[b] = its length
__global__ void pack_padded_kernel(const float4* __restrict__ flat, const int64_t* __restrict__ offsets, float4* __restrict__ out,
                                   int32_t* __restrict__ lengths_out, int Tmax, int d4, float pad) {
  const int t = blockIdx.x, b = blockIdx.y;
  const long o0 = offsets[b], len = offsets[b + 1] - o0;
  if (t == 0 && threadIdx.x == 0 && lengths_out) lengths_out[b] = (int32_t)(len < Tmax ? len : Tmax);
  float4* dst = out + ((long)b * Tmax + t) * d4;
  if (t < len) {
    const float4* src = flat + (o0 + t) * d4;
    for (int c = threadIdx.x; c < d4; c += blockDim.x) dst[c] = src[c];
  } else {
    const float4 z = make_float4(pad, pad, pad, pad);
    for (int c = threadIdx.x; c < d4; c += blockDim.x) dst[c] = z;
  }
}
// the inverse: valid rows of (B, Tmax, d) -> flat (sum_len, d)
__global__ void unpack_padded_kernel(const float4* __restrict__ padded, const int64_t* __restrict__ offsets, float4* __restrict__ flat,
                                     int Tmax, int d4) {
  const int t = blockIdx.x, b = blockIdx.y;
  const long o0 = offsets[b], len = offsets[b + 1] - o0;
  if (t >= len) return;
  const float4* src = padded + ((long)b * Tmax + t) * d4;
  float4* dst = flat + (o0 + t) * d4;
  for (int c = threadIdx.x; c < d4; c += blockDim.x) dst[c] = src[c];
}

}  // namespace avsr

using namespace avsr;

extern "C" {

int avsr_pack_padded(const float* flat, const int64_t* offsets, float* out, int32_t* lengths_out, int B, int Tmax, int d,
                     float pad_value, void* stream) {
  AVSR_REQUIRE(flat && offsets && out, "NULL argument");
  AVSR_REQUIRE(d > 0 && d % 4 == 0, "pack: d=%d must be a multiple of 4", d);
  if (B <= 0 || Tmax <= 0) return AVSR_OK;
  pack_padded_kernel<<<dim3(Tmax, B), 192, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const float4*>(flat), offsets, reinterpret_cast<float4*>(out), lengths_out, Tmax, d / 4, pad_value);
  AVSR_CHECK_LAUNCH();
  return AVSR_OK;
}

int avsr_unpack_padded(const float* padded, const int64_t* offsets, float* flat, int B, int Tmax, int d, void* stream) {
  AVSR_REQUIRE(padded && offsets && flat, "NULL argument");
  AVSR_REQUIRE(d > 0 && d % 4 == 0, "unpack: d=%d must be a multiple of 4", d);
  if (B <= 0 || Tmax <= 0) return AVSR_OK;
  unpack_padded_kernel<<<dim3(Tmax, B), 192, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const float4*>(padded), offsets, reinterpret_cast<float4*>(flat), Tmax, d / 4);
  AVSR_CHECK_LAUNCH();
  return AVSR_OK;
}

}  // extern "C"
