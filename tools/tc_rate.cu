// Issue-rate probe: how fast can one thread feed tcgen05.mma when descriptors are base + constant?
#include <cstdio>
#include "../deeptreeattention_b200/csrc/dta_tc.cuh"
using namespace dta::tc;

__device__ __forceinline__ uint64_t mk(uint32_t lo, uint32_t hi) {
  uint64_t d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(hi));
  return d;
}

template <int N, int MODE>
__global__ void rate_kernel(int reps, long long* cycles) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  fence_async_smem();
  if (warp == 0) tmem_alloc(&tmem_base, 512);
  if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_base;
  if (tid == 0) {
    constexpr uint32_t idesc = make_idesc_bf16(128, N, MODE, MODE);
    constexpr uint32_t ROWS = 1040;
    const uint64_t a0 = MODE == 0 ? sdesc_kmajor(smem_u32(smem), ROWS) : sdesc_mnmajor(smem_u32(smem), ROWS);
    const uint64_t b0 = MODE == 0 ? sdesc_kmajor(smem_u32(smem) + 40 * 1024, 64) : sdesc_mnmajor(smem_u32(smem) + 40 * 1024, 64);
    const uint32_t a_lo = (uint32_t)a0, a_hi = (uint32_t)(a0 >> 32), b_lo = (uint32_t)b0, b_hi = (uint32_t)(b0 >> 32);
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
#pragma unroll
      for (int s = 0; s < 4; ++s) {
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          const uint32_t shift = (t / 3) * 12 + (t % 3);
          mma_bf16(tmem + s * N, mk(a_lo + s * 128 + shift, a_hi), mk(b_lo + t * 8, b_hi), idesc, 1u);
        }
      }
    }
    mma_commit(&bar);
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    cycles[0] = t1 - t0;
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred;
}

template <int N, int MODE>
__global__ void rate_kernel_uniform(int reps, long long* cycles) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  for (int i = tid; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  fence_async_smem();
  if (warp == 0) tmem_alloc(&tmem_base, 512);
  if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_base;
  if (warp == 0) {
    constexpr uint32_t idesc = make_idesc_bf16(128, N, MODE, MODE);
    constexpr uint32_t ROWS = 1040;
    const uint64_t a0 = MODE == 0 ? sdesc_kmajor(smem_u32(smem), ROWS) : sdesc_mnmajor(smem_u32(smem), ROWS);
    const uint64_t b0 = MODE == 0 ? sdesc_kmajor(smem_u32(smem) + 40 * 1024, 64) : sdesc_mnmajor(smem_u32(smem) + 40 * 1024, 64);
    const uint32_t a_lo = (uint32_t)a0, a_hi = (uint32_t)(a0 >> 32), b_lo = (uint32_t)b0, b_hi = (uint32_t)(b0 >> 32);
    const uint32_t leader = elect_one();
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
#pragma unroll
      for (int s = 0; s < 4; ++s) {
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          const uint32_t shift = (t / 3) * 12 + (t % 3);
          if (leader) mma_bf16(tmem + s * N, mk(a_lo + s * 128 + shift, a_hi), mk(b_lo + t * 8, b_hi), idesc, 1u);
        }
      }
      __syncwarp();
    }
    if (leader) { mma_commit(&bar); }
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    if (leader) cycles[0] = t1 - t0;
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

template <int N, int MODE>
void run(int reps) {
  long long* dc; cudaMalloc(&dc, 8);
  cudaFuncSetAttribute(rate_kernel<N, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024);
  rate_kernel<N, MODE><<<1, 128, 48 * 1024>>>(reps, dc);
  cudaError_t e = cudaDeviceSynchronize();
  long long c = 0; cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
  printf("N %3d mode %d: %s %.1f cycles per MMA (floor %d)\n", N, MODE, cudaGetErrorString(e), (double)c / (reps * 36), 128 * N / 256);
  cudaFuncSetAttribute(rate_kernel_uniform<N, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024);
  rate_kernel_uniform<N, MODE><<<1, 128, 48 * 1024>>>(reps, dc);
  e = cudaDeviceSynchronize();
  cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
  printf("N %3d mode %d uniform-warp: %s %.1f cycles per MMA\n", N, MODE, cudaGetErrorString(e), (double)c / (reps * 36));
  cudaFree(dc);
}
int main() {
  run<32, 0>(200); run<48, 0>(200); run<64, 0>(200); run<128, 0>(200); run<256, 0>(200);
  run<32, 1>(200); run<48, 1>(200); run<64, 1>(200); run<128, 1>(200);
  return 0;
}
