// adaptdl_b200 -- common device helpers for the sm_100a gradient kernels.
//
// Everything here is bandwidth/latency code (no tensor cores): 128-bit
// vector loads/stores, system-scope release/acquire flags over NVLink peer
// mappings, warp-uniform per-group sum-of-squares accumulation.
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#define ADL_MAX_RANKS 16
#define ADL_MAX_CTAS 128
#define ADL_THREADS 512

// ---------------------------------------------------------------------------
// 128-bit vectors of T
// ---------------------------------------------------------------------------
template <typename T> struct VecTraits;
template <> struct VecTraits<float> { static constexpr int N = 4; };
template <> struct VecTraits<__nv_bfloat16> { static constexpr int N = 8; };
template <> struct VecTraits<__half> { static constexpr int N = 8; };

struct __align__(16) Vec16 { uint32_t w[4]; };

__device__ __forceinline__ Vec16 ld_vec(const void* p) {
  Vec16 v;
  asm volatile("ld.global.L1::no_allocate.v4.b32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.w[0]), "=r"(v.w[1]), "=r"(v.w[2]), "=r"(v.w[3])
               : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_vec(void* p, const Vec16& v) {
  asm volatile("st.global.L1::no_allocate.v4.b32 [%0], {%1,%2,%3,%4};"
               :: "l"(p), "r"(v.w[0]), "r"(v.w[1]), "r"(v.w[2]), "r"(v.w[3])
               : "memory");
}

template <typename T> __device__ __forceinline__ void unpack(const Vec16& v, float* f);
template <> __device__ __forceinline__ void unpack<float>(const Vec16& v, float* f) {
#pragma unroll
  for (int i = 0; i < 4; ++i) f[i] = __uint_as_float(v.w[i]);
}
template <> __device__ __forceinline__ void unpack<__nv_bfloat16>(const Vec16& v, float* f) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {   // bf16 -> fp32 is a 16-bit shift
    f[2 * i] = __uint_as_float(v.w[i] << 16);
    f[2 * i + 1] = __uint_as_float(v.w[i] & 0xffff0000u);
  }
}
template <> __device__ __forceinline__ void unpack<__half>(const Vec16& v, float* f) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    __half2 h = *reinterpret_cast<const __half2*>(&v.w[i]);
    float2 x = __half22float2(h);
    f[2 * i] = x.x; f[2 * i + 1] = x.y;
  }
}
template <typename T> __device__ __forceinline__ Vec16 pack(const float* f);
template <> __device__ __forceinline__ Vec16 pack<float>(const float* f) {
  Vec16 v;
#pragma unroll
  for (int i = 0; i < 4; ++i) v.w[i] = __float_as_uint(f[i]);
  return v;
}
template <> __device__ __forceinline__ Vec16 pack<__nv_bfloat16>(const float* f) {
  Vec16 v;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    v.w[i] = *reinterpret_cast<uint32_t*>(&h);
  }
  return v;
}
template <> __device__ __forceinline__ Vec16 pack<__half>(const float* f) {
  Vec16 v;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    __half2 h = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
    v.w[i] = *reinterpret_cast<uint32_t*>(&h);
  }
  return v;
}

// ---------------------------------------------------------------------------
// system-scope flags (peer-visible signal pads)
// ---------------------------------------------------------------------------
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint64_t globaltimer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// Signal-pad layout (uint32 words), one pad per rank, mapped on every peer:
//   [phase 0|1][cta < ADL_MAX_CTAS][src rank < ADL_MAX_RANKS]
// A rank WRITES slot (phase, cta, my_rank) on every peer's pad and WAITS on
// its own pad's slots (phase, cta, peer). Values are launch epochs (monotonic,
// wrap-safe comparison), so pads never need resetting.
__device__ __forceinline__ uint32_t* pad_slot(uint32_t* pad, int phase, int cta, int src) {
  return pad + ((phase * ADL_MAX_CTAS + cta) * ADL_MAX_RANKS + src);
}

// All threads of the CTA call this. Makes every prior global write of the CTA
// visible system-wide, then exchanges flags with CTA `cta` of every peer.
__device__ __forceinline__ void cta_rank_barrier(uint32_t* const* pads, int rank, int world,
                                                 int phase, int cta, uint32_t epoch) {
  __syncthreads();
  if (threadIdx.x < world) {
    const int peer = threadIdx.x;
    __threadfence_system();
    st_release_sys(pad_slot(pads[peer], phase, cta, rank), epoch);
    const uint32_t* mine = pad_slot(pads[rank], phase, cta, peer);
    while ((int32_t)(ld_acquire_sys(mine) - epoch) < 0) { __nanosleep(20); }
  }
  __syncthreads();
}

// ---------------------------------------------------------------------------
// segment -> statistics-group lookup
// seg_end[i] (exclusive, in vectors, relative to the bucket, strictly
// increasing, last == bucket length) ; group of vector v = seg_group[first i
// with seg_end[i] > v]. Padding vectors are all-zero so their group is moot.
// ---------------------------------------------------------------------------
struct SegTable {
  const int* seg_end;
  const int* seg_group;
  int n_seg;
};

__device__ __forceinline__ int seg_find(const SegTable& t, int v) {
  int lo = 0, hi = t.n_seg - 1;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (__ldg(t.seg_end + mid) > v) hi = mid; else lo = mid + 1;
  }
  return lo;
}

// ---------------------------------------------------------------------------
// Per-group sum-of-squares accumulation, K statistics at once.
//
// Hot path: while all 32 lanes of a warp are inside the same group as the
// previous iteration, each lane just adds into K fp32 registers (one warp vote
// per iteration, no shuffles). On a group change the warp shuffle-reduces and
// adds (fp64) into the CTA's shared-memory table; at the end of the kernel the
// table is flushed with one fp64 atomic per touched (statistic, group).
// ---------------------------------------------------------------------------
template <int K>
struct GroupAccum {
  float acc[K];
  int run_group;
  double* smem;      // [K][n_groups]
  int n_groups;

  __device__ __forceinline__ void init(double* s, int ng) {
    smem = s; n_groups = ng; run_group = -1;
#pragma unroll
    for (int k = 0; k < K; ++k) acc[k] = 0.f;
  }
  __device__ __forceinline__ void flush_warp() {
    if (run_group < 0) return;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      float x = acc[k];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
      if ((threadIdx.x & 31) == 0 && x != 0.f)
        atomicAdd(&smem[k * n_groups + run_group], (double)x);
      acc[k] = 0.f;
    }
    run_group = -1;
  }
  // all 32 lanes must call this together (inactive lanes pass g = -1, sq = 0)
  __device__ __forceinline__ void add(int g, const float* sq) {
    const int g0 = __shfl_sync(0xffffffffu, g, 0);
    const bool uniform = __all_sync(0xffffffffu, g == g0 || g < 0);
    if (uniform && g0 >= 0) {
      if (g0 != run_group) { flush_warp(); run_group = g0; }
#pragma unroll
      for (int k = 0; k < K; ++k) acc[k] += sq[k];
    } else {
      flush_warp();
      if (g >= 0) {
#pragma unroll
        for (int k = 0; k < K; ++k)
          if (sq[k] != 0.f) atomicAdd(&smem[k * n_groups + g], (double)sq[k]);
      }
    }
  }
};

__device__ __forceinline__ void smem_stats_zero(double* s, int n) {
  for (int i = threadIdx.x; i < n; i += blockDim.x) s[i] = 0.0;
  __syncthreads();
}
// out[k] points at a global double[n_groups] (or nullptr)
template <int K>
__device__ __forceinline__ void smem_stats_flush(const double* s, int n_groups, double* const* out) {
  __syncthreads();
  for (int i = threadIdx.x; i < K * n_groups; i += blockDim.x) {
    const double x = s[i];
    const int k = i / n_groups;
    if (x != 0.0 && out[k] != nullptr) atomicAdd(out[k] + (i - k * n_groups), x);
  }
}
