// Gradient exchange over NVLink / NVSwitch peer memory: the one collective of the path, written as
// ONE kernel instead of an NCCL call.  Every rank keeps its flat gradient buffer in symmetric memory (peer-mapped on
// all ranks, optionally also bound to an NVSwitch multicast address).  The kernel
//   1. announces "my gradients are written" to every peer and waits for theirs (flags in the peers' buffers,
//      release/acquire at system scope),
//   2. reads the SUM over ranks of every element -- one multimem.ld_reduce per 16 bytes when a multicast address is
//      available (the switch adds the replicas, NVLS), else plain loads from each peer in rank order -- scales by
//      1/world and parks it in a local scratch buffer,
//   3. announces "done reading" / waits for the peers' same announcement, then copies the scratch back over its own
//      buffer, so p.grad (views of that buffer) holds the average exactly where an NCCL all-reduce would leave it.
// With a multicast mapping (NVSwitch, NVLS) steps 2-3 collapse: every rank OWNS one slice of the buffer, reads that slice's
// sum over all replicas with multimem.ld_reduce, scales it and writes the mean to ALL replicas with multimem.st (the
// switch broadcasts) -- no scratch round trip, 1/world of the traffic per rank, and the second flag round only says
// "my slice has landed everywhere".  Nobody but its owner ever reads a slice, so the in-place update cannot race.
// Summation order is fixed (rank order, or the switch's order for NVLS) and every element is computed exactly once (NVLS)
// or identically on every rank (peer loads): all ranks end with identical bits.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace dta {

constexpr int kMaxPeers = 16;
constexpr int kArThreads = 512;

struct PeerPtrs {
  float* buf[kMaxPeers];   // buf[r] = rank r's symmetric buffer as mapped in THIS process
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" : : "l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 multimem_sum_f32x4(const float* mc) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(mc) : "memory");
  return v;
}
__device__ __forceinline__ void multimem_store_f32x4(float* mc, float4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" : : "l"(mc), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ double multimem_sum_f64(const double* mc) {
  double v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.f64 %0, [%1];" : "=d"(v) : "l"(mc) : "memory");
  return v;
}

// Symmetric buffer layout (bytes): [n4*16: fp32 payload, n4 float4][nd*8: fp64 payload][flags: 2 sets of start[world] | end[world] u32]
// (two flag sets: the exchange of a backward pass runs as two kernels that may overlap, see dta_set_grad_exchange)
__host__ __device__ inline size_t ar_flags_offset(size_t n4, size_t nd) { return (n4 * 16 + nd * 8 + 127) / 128 * 128; }
__host__ __device__ inline size_t ar_buffer_bytes(size_t n4, size_t nd, int world) { return ar_flags_offset(n4, nd) + 4 * (size_t)world * 4 + 128; }

// Which part of the float32 payload a launch exchanges: up to three [lo, hi) ranges in float4 units (n = 0: everything),
// and whether it also carries the float64 tail.  Slice ownership is over the concatenation of the ranges.
struct ArRanges {
  unsigned long long lo[3], hi[3];
  int n;
  int with_doubles;
};

// sync[0] = epoch (starts at 0), sync[1] = CTAs done with phase 2, sync[2] = CTAs finished.  Local, zero-initialised.
__global__ void __launch_bounds__(kArThreads)
grad_allreduce_kernel(PeerPtrs peers, const float* __restrict__ mc /*multicast view of the buffers or null*/, int rank, int world,
                      size_t n4, size_t nd, float* __restrict__ scratch /*n4*4 floats + nd doubles (8-byte aligned tail)*/,
                      uint32_t* __restrict__ sync_base, ArRanges rg, int flag_set) {
  uint32_t* sync = sync_base + 4 * flag_set;
  const uint32_t epoch = sync[0] + 1;
  const size_t flags_off = ar_flags_offset(n4, nd) + (size_t)flag_set * 2 * world * 4;
  if (rg.n == 0) { rg.n = 1; rg.lo[0] = 0; rg.hi[0] = n4; rg.with_doubles = 1; }
  if (!rg.with_doubles) nd = 0;
  uint32_t* my_start = reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(peers.buf[rank]) + flags_off);
  uint32_t* my_end = my_start + world;
  const float inv = 1.0f / (float)world;

  // ---- 1. my gradients are complete (kernel boundary) -> tell every peer; wait for every peer ----
  if (blockIdx.x == 0 && threadIdx.x < world) {
    __threadfence_system();
    uint32_t* their_start = reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(peers.buf[threadIdx.x]) + flags_off);
    st_release_sys(their_start + rank, epoch);
  }
  if (threadIdx.x < world) {
    while (ld_acquire_sys(my_start + threadIdx.x) < epoch) {
    }
  }
  __syncthreads();

  const size_t stride = (size_t)gridDim.x * blockDim.x;
  __shared__ uint32_t s_last;
  if (mc != nullptr) {
    // ---- 2'. NVLS: reduce MY slice in the switch, broadcast its mean to every replica ----
    size_t total = 0;
    for (int k = 0; k < rg.n; ++k) total += rg.hi[k] - rg.lo[k];
    const size_t per = (total + world - 1) / world;
    const size_t lo = (size_t)rank * per, hi = lo + per < total ? lo + per : total;
    float* mcw = const_cast<float*>(mc);
    for (size_t v = lo + (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < hi; v += stride) {
      size_t i = v;                                  // position in the concatenated ranges -> float4 index in the buffer
      int k = 0;
      while (k + 1 < rg.n && i >= rg.hi[k] - rg.lo[k]) { i -= rg.hi[k] - rg.lo[k]; ++k; }
      i += rg.lo[k];
      float4 a = multimem_sum_f32x4(mc + i * 4);
      a.x *= inv; a.y *= inv; a.z *= inv; a.w *= inv;
      multimem_store_f32x4(mcw + i * 4, a);
    }
    if (rank == 0) {   // the float64 tail (alpha): summed in the switch, written to every peer with plain stores
      for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nd; i += stride) {
        const double a = multimem_sum_f64(reinterpret_cast<const double*>(reinterpret_cast<const char*>(mc) + n4 * 16) + i) / (double)world;
        for (int r = 0; r < world; ++r) reinterpret_cast<double*>(reinterpret_cast<char*>(peers.buf[r]) + n4 * 16)[i] = a;
      }
    }
    // ---- 3'. my slice is everywhere -> tell the peers; wait until every owner's slice has landed in MY buffer ----
    // (the CTA barrier orders every thread's stores before thread 0's system-scope fence; its device-scope counter update
    //  then carries them to the CTA that signals: one fence per CTA instead of one per thread)
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence_system();
      s_last = (atomicAdd(&sync[1], 1u) == gridDim.x - 1) ? 1u : 0u;
    }
    __syncthreads();
    if (s_last && threadIdx.x < world) {
      __threadfence_system();
      uint32_t* their_end = reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(peers.buf[threadIdx.x]) + flags_off) + world;
      st_release_sys(their_end + rank, epoch);
    }
    if (threadIdx.x < world) {
      while (ld_acquire_sys(my_end + threadIdx.x) < epoch) {
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      if (atomicAdd(&sync[2], 1u) == gridDim.x - 1) {
        sync[1] = 0;
        sync[2] = 0;
        __threadfence();
        sync[0] = epoch;
      }
    }
    return;
  }

  // ---- 2. sum over ranks, scale, park in scratch ----
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 a;
    if (mc != nullptr) {
      a = multimem_sum_f32x4(mc + i * 4);
    } else {
      a = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int r = 0; r < world; ++r) {
        const float4 v = __ldcv(reinterpret_cast<const float4*>(peers.buf[r]) + i);   // volatile: never a stale cached line
        a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
      }
    }
    a.x *= inv; a.y *= inv; a.z *= inv; a.w *= inv;
    reinterpret_cast<float4*>(scratch)[i] = a;
  }
  double* scratch_d = reinterpret_cast<double*>(scratch + n4 * 4);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nd; i += stride) {
    double a = 0.0;
    if (mc != nullptr) {
      a = multimem_sum_f64(reinterpret_cast<const double*>(reinterpret_cast<const char*>(mc) + n4 * 16) + i);
    } else {
      for (int r = 0; r < world; ++r) a += __ldcv(reinterpret_cast<const double*>(reinterpret_cast<const char*>(peers.buf[r]) + n4 * 16) + i);
    }
    scratch_d[i] = a / (double)world;
  }

  // ---- 3. everybody on this rank done reading -> tell the peers; wait until nobody reads MY buffer any more ----
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(&sync[1], 1u) == gridDim.x - 1) ? 1u : 0u;
  __syncthreads();
  if (s_last && threadIdx.x < world) {
    uint32_t* their_end = reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(peers.buf[threadIdx.x]) + flags_off) + world;
    st_release_sys(their_end + rank, epoch);
  }
  if (threadIdx.x < world) {
    while (ld_acquire_sys(my_end + threadIdx.x) < epoch) {
    }
  }
  __syncthreads();
  float4* mine = reinterpret_cast<float4*>(peers.buf[rank]);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) mine[i] = reinterpret_cast<const float4*>(scratch)[i];
  double* mine_d = reinterpret_cast<double*>(reinterpret_cast<char*>(peers.buf[rank]) + n4 * 16);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nd; i += stride) mine_d[i] = scratch_d[i];

  // ---- bookkeeping for the next launch (the last CTA to finish advances the epoch and clears the counters) ----
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(&sync[2], 1u) == gridDim.x - 1) {
      sync[1] = 0;
      sync[2] = 0;
      __threadfence();
      sync[0] = epoch;
    }
  }
}

}  // namespace dta
