// Thin inline-PTX layer for the Blackwell (sm_100a) tensor-core path: tcgen05.mma with
// shared-memory matrix descriptors, tensor-memory (TMEM) allocation and loads, mbarriers and
// 1-D bulk async copies.  Only what the convolution kernels need; no CUTLASS.
//
// Shared-memory operand layout used everywhere in this library ("chunked rows", no swizzle):
//     buf[chunk][row][8 x bf16]            16 bytes per (chunk, row), rows contiguous
// * read as a K-major operand  : row = M/N index, chunk = 8-wide slice of K
//       -> canonical INTERLEAVE K-major  ((8,n),2):((1,SBO),LBO) in 16-byte units with
//          SBO = 128 B (8 rows), LBO = rows*16 B (next K slice)
// * read as an MN-major operand: row = K index, chunk = 8-wide slice of M/N
//       -> canonical INTERLEAVE MN-major ((1,n),(8,k)):((X,SBO),(1,LBO)) with
//          LBO = 128 B (next 8 K rows), SBO = rows*16 B (next M/N slice)
// Because every address is linear in `row`, a 3x3 tap is just `start address += shift*16 B`.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace dta {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---------------------------------------------------------------- descriptors
// Instruction descriptor, kind::f16, BF16 x BF16 -> F32 (bit layout: cute/arch/mma_sm100_desc.hpp
// InstrDescriptor): c_format[4,6)=1 (F32), a_format[7,10)=1 (BF16), b_format[10,13)=1,
// a_major bit 15, b_major bit 16 (0 = K-major, 1 = MN-major), n_dim[17,23) = N>>3, m_dim[24,29) = M>>4.
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// Shared-memory matrix descriptor, SWIZZLE_NONE (layout_type 0), version 1 (Blackwell):
// start[0,14) = addr>>4, LBO[16,30) = bytes>>4, SBO[32,46) = bytes>>4, version[46,48) = 1.
__device__ __forceinline__ uint64_t make_sdesc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
// K-major view of a chunked-rows buffer: `rows` rows per chunk.
__device__ __forceinline__ uint64_t sdesc_kmajor(uint32_t smem_addr, uint32_t rows) {
  return make_sdesc(smem_addr, rows * 16u, 128u);
}
// MN-major view of the same buffer.
__device__ __forceinline__ uint64_t sdesc_mnmajor(uint32_t smem_addr, uint32_t rows) {
  return make_sdesc(smem_addr, 128u, rows * 16u);
}

// ---------------------------------------------------------------- tcgen05
// D[tmem] (+)= A[smem] * B[smem]; issued by ONE thread.
__device__ __forceinline__ void mma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      :
      : "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on an mbarrier once every previously issued MMA of this thread has completed.
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" : : "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" : : : "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" : : : "memory"); }

// One full warp allocates `cols` (power of two >= 32) TMEM columns; the base address lands in *dst (shared).
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" : : "r"(smem_u32(dst_smem)), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" : : : "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" : : "r"(taddr), "r"(cols) : "memory");
}

// Warp-wide TMEM load: lane t of warp w (w = warp index % 4) reads TMEM lane 32*w + t, 16 / 32
// consecutive 32-bit columns starting at the column in taddr.  taddr = (lane << 16) | column.
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" : : : "memory"); }

// ---------------------------------------------------------------- mbarrier / async proxy
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" : : "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" : : : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" : : "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" : : "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// Make this thread's generic-proxy shared-memory writes visible to the async proxy (tcgen05.mma reads).
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" : : : "memory"); }

// 1-D bulk copy global -> shared (TMA engine, no tensor map); completes `bytes` on the mbarrier.
// dst, src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :
               : "r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ---------------------------------------------------------------- split precision
// fp32 -> (hi, lo) bf16 pair with hi + lo = v to ~16 mantissa bits.  Packs two values.
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  const float ra = a - __low2float(h), rb = b - __high2float(h);
  __nv_bfloat162 l = __floats2bfloat162_rn(ra, rb);
  hi = *reinterpret_cast<uint32_t*>(&h);
  lo = *reinterpret_cast<uint32_t*>(&l);
}

}  // namespace tc
}  // namespace dta
