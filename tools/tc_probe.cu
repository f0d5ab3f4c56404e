// Hardware probe for the tcgen05 operand layouts the convolution kernels rely on
// (deeptreeattention_b200/csrc/dta_tc.cuh): no-swizzle "chunked rows" buffers read as K-major
// and as MN-major operands with an arbitrary row shift, plus the MMA issue rate.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o build/tc_probe tools/tc_probe.cu
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../deeptreeattention_b200/csrc/dta_tc.cuh"

using namespace dta::tc;

// mode 0: K-major A (rows = M, shifted by `shift` rows), K-major B.   D[m][n] = sum_k A[m+shift][k] B[n][k]
// mode 1: MN-major A (rows = K, shifted), MN-major B.                 D[m][n] = sum_k A[k+shift][m] B[k][n]
__global__ void probe_kernel(const __nv_bfloat16* __restrict__ a_buf, const __nv_bfloat16* __restrict__ b_buf, int a_rows,
                             int a_chunks, int b_rows, int b_chunks, int mode, int shift, int N, int ksteps, int reps,
                             float* __restrict__ d_out, long long* __restrict__ cycles) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  __nv_bfloat16* sa = reinterpret_cast<__nv_bfloat16*>(smem);
  __nv_bfloat16* sb = sa + (size_t)a_rows * a_chunks * 8;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < a_rows * a_chunks * 8; i += blockDim.x) sa[i] = a_buf[i];
  for (int i = tid; i < b_rows * b_chunks * 8; i += blockDim.x) sb[i] = b_buf[i];
  fence_async_smem();
  if (warp == 0) tmem_alloc(&tmem_base, 128);
  if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_base;
  if (tid == 0) {
    const uint32_t idesc = make_idesc_bf16(128, N, mode, mode);
    const uint32_t a0 = smem_u32(sa) + shift * 16, b0 = smem_u32(sb);
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      for (int ks = 0; ks < ksteps; ++ks) {
        uint64_t ad, bd;
        if (mode == 0) {  // K advances by 2 chunks per MMA
          ad = sdesc_kmajor(a0 + ks * 2 * a_rows * 16, a_rows);
          bd = sdesc_kmajor(b0 + ks * 2 * b_rows * 16, b_rows);
        } else {          // K advances by 16 rows per MMA
          ad = sdesc_mnmajor(a0 + ks * 16 * 16, a_rows);
          bd = sdesc_mnmajor(b0 + ks * 16 * 16, b_rows);
        }
        mma_bf16(tmem, ad, bd, idesc, (r | ks) ? 1u : 0u);
      }
    }
    mma_commit(&bar);
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    cycles[0] = t1 - t0;
  }
  __syncthreads();
  fence_after_sync();
  for (int c0 = 0; c0 < N; c0 += 16) {
    float v[16];
    tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
    tmem_ld_wait();
    for (int j = 0; j < 16; ++j) d_out[(size_t)(warp * 32 + lane) * N + c0 + j] = v[j];
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 128);
}

static float bf(float v) { return __bfloat162float(__float2bfloat16(v)); }

int run(int mode, int shift, int N, int ksteps, int reps, bool verbose) {
  const int M = 128, K = 16 * ksteps;
  // logical matrices with small integers (exact in bf16 and in fp32 accumulation)
  int a_rows, a_chunks, b_rows, b_chunks;
  if (mode == 0) { a_rows = M + 32; a_chunks = K / 8; b_rows = N; b_chunks = K / 8; }
  else { a_rows = K + 32; a_chunks = M / 8; b_rows = K; b_chunks = N / 8; }
  std::vector<__nv_bfloat16> ha((size_t)a_rows * a_chunks * 8), hb((size_t)b_rows * b_chunks * 8);
  std::vector<float> fa(ha.size()), fb(hb.size());
  srand(1234 + mode * 7 + shift);
  for (size_t i = 0; i < ha.size(); ++i) { fa[i] = (float)(rand() % 17 - 8); ha[i] = __float2bfloat16(fa[i]); }
  for (size_t i = 0; i < hb.size(); ++i) { fb[i] = (float)(rand() % 13 - 6); hb[i] = __float2bfloat16(fb[i]); }
  auto A = [&](int row, int col) { return fa[((size_t)(col / 8) * a_rows + row) * 8 + col % 8]; };   // buf[chunk][row][8]
  auto Bm = [&](int row, int col) { return fb[((size_t)(col / 8) * b_rows + row) * 8 + col % 8]; };
  std::vector<float> ref((size_t)M * N, 0.f);
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      float s = 0.f;
      for (int k = 0; k < K; ++k) s += mode == 0 ? A(m + shift, k) * Bm(n, k) : A(k + shift, m) * Bm(k, n);
      ref[(size_t)m * N + n] = s * reps;
    }
  __nv_bfloat16 *da, *db; float* dd; long long* dc;
  cudaMalloc(&da, ha.size() * 2); cudaMalloc(&db, hb.size() * 2); cudaMalloc(&dd, ref.size() * 4); cudaMalloc(&dc, 8);
  cudaMemcpy(da, ha.data(), ha.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(db, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice);
  cudaMemset(dd, 0, ref.size() * 4);
  size_t smem = (ha.size() + hb.size()) * 2;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  probe_kernel<<<1, 128, smem>>>(da, db, a_rows, a_chunks, b_rows, b_chunks, mode, shift, N, ksteps, reps, dd, dc);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("mode %d shift %d N %d: CUDA error %s\n", mode, shift, N, cudaGetErrorString(e)); return 1; }
  std::vector<float> out(ref.size());
  long long cyc = 0;
  cudaMemcpy(out.data(), dd, out.size() * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(&cyc, dc, 8, cudaMemcpyDeviceToHost);
  int bad = 0; double maxerr = 0;
  for (size_t i = 0; i < out.size(); ++i) {
    double d = fabs((double)out[i] - ref[i]);
    if (d > maxerr) maxerr = d;
    if (d > 1e-3 * (1 + fabs(ref[i]))) ++bad;
  }
  printf("mode %d (%s) shift %2d N %3d K %3d reps %4d: %s  mismatches %d / %zu  maxerr %.3g  cycles %lld (%.1f per MMA)\n", mode,
         mode == 0 ? "K-major " : "MN-major", shift, N, K, reps, bad ? "FAIL" : "ok", bad, out.size(), maxerr, cyc,
         (double)cyc / (reps * ksteps));
  if (bad && verbose)
    for (int m = 0; m < 4; ++m) { for (int n = 0; n < 8; ++n) printf(" %7.1f/%7.1f", out[m * N + n], ref[m * N + n]); printf("\n"); }
  cudaFree(da); cudaFree(db); cudaFree(dd); cudaFree(dc);
  (void)bf;
  return bad != 0;
}

int main() {
  int fails = 0;
  for (int mode = 0; mode < 2; ++mode)
    for (int shift : {0, 1, 5, 13, 26})
      for (int N : {64, 48, 32, 128}) {
        fails += run(mode, shift, N, 4, 1, true);
      }
  // issue rate (results still checked: reps accumulate the same product)
  for (int N : {32, 48, 64}) fails += run(0, 3, N, 4, 256, false);
  for (int N : {32, 48, 64}) fails += run(1, 3, N, 4, 256, false);
  printf(fails ? "PROBE FAILED (%d)\n" : "PROBE OK\n", fails);
  return fails != 0;
}
