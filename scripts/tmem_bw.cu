// Micro-benchmark: tcgen05.ld (TMEM -> registers) throughput per SM as a function of the number of reading warps and
// of the load width.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 scripts/tmem_bw.cu -o scripts/bin/tmem_bw
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int X>
__device__ __forceinline__ void ld(uint32_t taddr, uint32_t* r);
template <>
__device__ __forceinline__ void ld<32>(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
template <>
__device__ __forceinline__ void ld<16>(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

// each of `nwarps` warps issues `iters` x (DEPTH loads of X columns, then one wait::ld); returns cycles of warp 0
template <int X, int DEPTH>
__global__ void tmem_read_kernel(int iters, unsigned long long* out, unsigned* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = slot;
  const uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 64 % 256);
  uint32_t r[DEPTH][X];
  unsigned acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int d = 0; d < DEPTH; ++d) ld<X>(trow + (uint32_t)(d * X % 192), r[d]);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int d = 0; d < DEPTH; ++d) acc ^= r[d][0] ^ r[d][X - 1];   // static indices: the arrays stay in registers
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = (unsigned long long)(t1 - t0);
  if (acc == 0x12345678u) sink[0] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
  }
}

template <int X, int DEPTH>
static void run(int nwarps, int blocks) {
  unsigned long long* out;
  unsigned* sink;
  cudaMalloc(&out, blocks * sizeof(unsigned long long));
  cudaMalloc(&sink, 4);
  const int iters = 2000;
  tmem_read_kernel<X, DEPTH><<<blocks, nwarps * 32>>>(iters, out, sink);
  tmem_read_kernel<X, DEPTH><<<blocks, nwarps * 32>>>(iters, out, sink);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return; }
  unsigned long long h;
  cudaMemcpy(&h, out, sizeof h, cudaMemcpyDeviceToHost);
  const double bytes = (double)nwarps * iters * DEPTH * X * 32 * 4;
  printf("x%-2d depth %d warps %2d blocks %3d: %8llu cycles  -> %7.1f B/cycle/SM  (%.0f cycles per warp-load)\n", X, DEPTH,
         nwarps, blocks, h, bytes / (double)h, (double)h / (iters * DEPTH));
  cudaFree(out); cudaFree(sink);
}

int main() {
  for (int w : {1, 2, 4, 8, 16}) run<32, 1>(w, 1);
  for (int w : {1, 4, 8, 16}) run<32, 2>(w, 1);
  for (int w : {4, 8}) run<32, 4>(w, 1);
  for (int w : {1, 4, 8}) run<16, 1>(w, 1);
  for (int w : {4, 8}) run<16, 2>(w, 1);
  for (int w : {4, 8}) run<16, 4>(w, 1);
  for (int w : {8}) run<32, 2>(w, 148);
  return 0;
}
