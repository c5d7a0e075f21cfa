// adaptdl_b200 -- sm_100a gradient kernels (C ABI, launched from Python via
// ctypes on torch's CUDA streams).
//
//   adl_fold_acc       A += G ; L += |G/P|^2 ; G = 0
//   adl_fold_final     L += |G/P|^2 ; G += A ; A = 0
//   adl_allreduce_gns  fused two-shot all-reduce over NVLink peer mappings:
//                      G <- s * sum_r G_r (in place on every rank) with
//                      L += sum_r |G_r/P|^2 and T += |G/P|^2 from the same
//                      registers (reference call sites K1-K6)
//   adl_pair_norm      T=|G/P|^2, Pp=|Pv/P|^2, Pa=|(G+Pv)/2P|^2 ; Pv = G
//   adl_finalize_stats sum per-rank partial statistics over ranks through a
//                      peer-mapped pad, publish to a pinned host mailbox with
//                      %globaltimer stamps, reset the partials
//   adl_bcast_pull     rank src's staging buffer -> every rank
//   adl_stamp          write %globaltimer to device memory
//
// No tensor cores: these are bandwidth / latency kernels. sm_100a specifics:
// 128-bit L1-bypassing vector accesses, .sys-scope release/acquire flags on
// NVLink-mapped signal pads, %globaltimer stamps, grids sized to leave SMs to
// the concurrently running backward pass.
#include "adl_common.cuh"

#include <stdio.h>

// error word bits (device -> host, sticky)
#define ADL_ERR_TIMEOUT 1u
#define ADL_MAX_STAT_SMEM (96 * 1024)

struct ReduceArgs {
  void* buf[ADL_MAX_RANKS];        // bucket start in every rank's G arena
  uint32_t* pad[ADL_MAX_RANKS];    // signal pad of every rank
  int rank, world;
  const uint32_t* step_ctr;        // device: optimizer steps finalized so far
  uint32_t site;                   // launch ordinal within the current step
  int n_vec;                       // vectors in the bucket (multiple of world)
  float scale;
  int want_local;
  SegTable segs;
  int n_groups;
  const void* pinv;                // local preconditioner slice or nullptr
  double* L;                       // [n_groups] partial: sum_r |G_r/P|^2
  double* T;                       // [n_groups] partial: |G/P|^2
  uint32_t* err;                   // sticky error word (device)
  unsigned long long timeout_ns;
  void* mc_buf;                    // NVLS: multicast address of the bucket (or nullptr)
};

__device__ __forceinline__ bool wait_flag(const uint32_t* p, uint32_t epoch,
                                          unsigned long long timeout_ns, uint32_t* err) {
  const uint64_t t0 = globaltimer_ns();
  uint32_t spins = 0;
  while ((int32_t)(ld_acquire_sys(p) - epoch) < 0) {
    __nanosleep(32);
    if ((++spins & 1023u) == 0 && globaltimer_ns() - t0 > timeout_ns) {
      atomicOr(err, ADL_ERR_TIMEOUT);
      return false;
    }
  }
  return true;
}

// Flag values ("epochs") are derived on the device: step counter (bumped by
// the finalize kernel once per optimizer step) * ADL_SITES_PER_STEP + the
// launch's ordinal within the step. Nothing launch-specific is baked into
// kernel arguments, so a captured CUDA graph can be replayed step after step.
#define ADL_SITES_PER_STEP 1024u
__device__ __forceinline__ uint32_t launch_epoch(const uint32_t* step_ctr, uint32_t site) {
  return (*reinterpret_cast<const volatile uint32_t*>(step_ctr)) * ADL_SITES_PER_STEP + site;
}

__device__ __forceinline__ void cta_barrier_peers(const ReduceArgs& a, int phase) {
  __syncthreads();
  if ((int)threadIdx.x < a.world) {
    const int peer = threadIdx.x;
    const uint32_t epoch = launch_epoch(a.step_ctr, a.site);
    __threadfence_system();
    st_release_sys(pad_slot(a.pad[peer], phase, blockIdx.x, a.rank), epoch);
    wait_flag(pad_slot(a.pad[a.rank], phase, blockIdx.x, peer), epoch, a.timeout_ns, a.err);
  }
  __syncthreads();
}

// ---------------------------------------------------------------------------
// fused two-shot all-reduce + gradient-noise-scale statistics
// ---------------------------------------------------------------------------
// W > 0: world size known at compile time (2, 4, 8): the per-peer loads are a
// fully unrolled register array and each thread keeps U = 16/W vectors in
// flight (16 independent 16-byte requests per thread; with 32 CTAs x 512
// threads that is ~4 MB outstanding, enough to cover the ~2-3 us NVLink
// round trip at full link bandwidth). W == 0: generic fallback, runtime world.
template <int W> struct ReduceUnroll { static constexpr int U = 16 / W; };
template <> struct ReduceUnroll<0> { static constexpr int U = 1; };
template <> struct ReduceUnroll<1> { static constexpr int U = 8; };

template <typename T, int W, bool HAS_PINV>
__global__ void __launch_bounds__(ADL_THREADS, 1)
allreduce_gns_kernel(const ReduceArgs a) {
  extern __shared__ double s_stats[];                 // [2][n_groups]
  constexpr int N = VecTraits<T>::N;
  constexpr int U = ReduceUnroll<W>::U;
  constexpr int WMAX = (W > 0) ? W : ADL_MAX_RANKS;
  smem_stats_zero(s_stats, 2 * a.n_groups);
  GroupAccum<2> accum;
  accum.init(s_stats, a.n_groups);

  const int world = (W > 0) ? W : a.world;
  if (world > 1) cta_barrier_peers(a, 0);             // every rank's grads are ready

  const int slice = a.n_vec / world;
  const int base = a.rank * slice;
  const int stride = gridDim.x * blockDim.x;
  const int first = blockIdx.x * blockDim.x + threadIdx.x;
  const int iters = (slice + stride * U - 1) / (stride * U);   // same for every lane
  int cur[U];
#pragma unroll
  for (int u = 0; u < U; ++u) cur[u] = -1;

  // peer order rotated so that rank r starts with its own copy and the ranks
  // do not all hammer the same peer at once
  const Vec16* src[WMAX];
  Vec16* dst[WMAX];
#pragma unroll
  for (int p = 0; p < WMAX; ++p) {
    const int q = (p < world) ? (a.rank + p) % world : a.rank;
    src[p] = static_cast<const Vec16*>(a.buf[q]);
    dst[p] = static_cast<Vec16*>(a.buf[q]);
  }

  for (int it = 0; it < iters; ++it) {
    Vec16 in[U][WMAX];
    Vec16 pv[U];
    int idx[U];
    bool active[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      idx[u] = first + (it * U + u) * stride;
      active[u] = idx[u] < slice;
      if (active[u]) {
#pragma unroll
        for (int p = 0; p < WMAX; ++p)
          if (p < world) in[u][p] = ld_vec(src[p] + base + idx[u]);
        if (HAS_PINV) pv[u] = ld_vec(static_cast<const Vec16*>(a.pinv) + base + idx[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      float sq[2] = {0.f, 0.f};
      int g = -1;
      if (active[u]) {
        const int v = base + idx[u];
        float sum[N], pinv[N];
        if (HAS_PINV) {
          unpack<T>(pv[u], pinv);
#pragma unroll
          for (int e = 0; e < N; ++e) pinv[e] = 1.0f / pinv[e];
        }
#pragma unroll
        for (int e = 0; e < N; ++e) sum[e] = 0.f;
#pragma unroll
        for (int p = 0; p < WMAX; ++p) {
          if (p < world) {
            float x[N];
            unpack<T>(in[u][p], x);
#pragma unroll
            for (int e = 0; e < N; ++e) {
              sum[e] += x[e];
              const float y = HAS_PINV ? x[e] * pinv[e] : x[e];
              sq[0] = fmaf(y, y, sq[0]);
            }
          }
        }
#pragma unroll
        for (int e = 0; e < N; ++e) {
          sum[e] *= a.scale;
          const float y = HAS_PINV ? sum[e] * pinv[e] : sum[e];
          sq[1] = fmaf(y, y, sq[1]);
        }
        const Vec16 out = pack<T>(sum);
#pragma unroll
        for (int p = 0; p < WMAX; ++p)
          if (p < world) st_vec(dst[p] + v, out);
        if (cur[u] < 0) cur[u] = seg_find(a.segs, v);
        while (__ldg(a.segs.seg_end + cur[u]) <= v) ++cur[u];
        g = __ldg(a.segs.seg_group + cur[u]);
        if (!a.want_local) sq[0] = 0.f;
      }
      accum.add(g, sq);
    }
  }
  accum.flush_warp();
  double* outs[2] = {a.want_local ? a.L : nullptr, a.T};
  smem_stats_flush<2>(s_stats, a.n_groups, outs);

  if (world > 1) cta_barrier_peers(a, 1);             // every slice has landed everywhere
}

// ---------------------------------------------------------------------------
// NVLS flavour: the NVSwitch reduces. `multimem.ld_reduce` on the multicast
// address returns sum_r g_r of a vector in ONE load (the switch pulls every
// GPU's copy and adds in flight), `multimem.st` writes the mean into every
// GPU's arena with ONE store. Per GPU that is ~B out + ~B(1+1/W) in instead
// of 2(W-1)/W*B each way, and W times fewer load instructions.
// The switch hides the per-replica values, so sum_r |g_r|^2 comes from a local
// pass over this rank's own bucket (HBM speed, before the start barrier --
// peers may only overwrite it after they have seen this rank's start flag);
// the per-rank partials are summed by the finalize kernel like all others.
// ---------------------------------------------------------------------------
template <typename T> struct Multimem;
template <> struct Multimem<float> {
  static __device__ __forceinline__ Vec16 ld_reduce(const void* p) {
    Vec16 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.w[0]), "=r"(v.w[1]), "=r"(v.w[2]), "=r"(v.w[3]) : "l"(p) : "memory");
    return v;
  }
};
template <> struct Multimem<__nv_bfloat16> {
  static __device__ __forceinline__ Vec16 ld_reduce(const void* p) {
    Vec16 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.bf16x2 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.w[0]), "=r"(v.w[1]), "=r"(v.w[2]), "=r"(v.w[3]) : "l"(p) : "memory");
    return v;
  }
};
template <> struct Multimem<__half> {
  static __device__ __forceinline__ Vec16 ld_reduce(const void* p) {
    Vec16 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.f16x2 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.w[0]), "=r"(v.w[1]), "=r"(v.w[2]), "=r"(v.w[3]) : "l"(p) : "memory");
    return v;
  }
};
__device__ __forceinline__ void multimem_st(void* p, const Vec16& v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};"
               :: "l"(p), "r"(v.w[0]), "r"(v.w[1]), "r"(v.w[2]), "r"(v.w[3]) : "memory");
}

template <typename T, bool HAS_PINV>
__global__ void __launch_bounds__(ADL_THREADS, 1)
allreduce_nvls_kernel(const ReduceArgs a) {
  extern __shared__ double s_stats[];                 // [2][n_groups]
  constexpr int N = VecTraits<T>::N;
  constexpr int U = 4;
  smem_stats_zero(s_stats, 2 * a.n_groups);
  GroupAccum<2> accum;
  accum.init(s_stats, a.n_groups);
  const int stride = gridDim.x * blockDim.x;
  const int first = blockIdx.x * blockDim.x + threadIdx.x;
  const Vec16* mine = static_cast<const Vec16*>(a.buf[a.rank]);

  const int slice = a.n_vec / a.world;
  const int iters = (slice + stride * U - 1) / (stride * U);
  if (a.want_local) {
    // L += |g_local / P|^2 over the whole bucket. CTA c of this rank reads,
    // in EVERY slice q, exactly the vectors that CTA c of rank q will later
    // overwrite -- the per-CTA start barrier below is then enough to order
    // these reads before the peers' multicast stores.
    for (int q = 0; q < a.world; ++q) {
      int cur[U];
#pragma unroll
      for (int u = 0; u < U; ++u) cur[u] = -1;
      for (int it = 0; it < iters; ++it) {
        Vec16 in[U], pv[U];
        int idx[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          idx[u] = first + (it * U + u) * stride;
          if (idx[u] < slice) {
            in[u] = ld_vec(mine + q * slice + idx[u]);
            if (HAS_PINV) pv[u] = ld_vec(static_cast<const Vec16*>(a.pinv) + q * slice + idx[u]);
          }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          float sq[2] = {0.f, 0.f};
          int g = -1;
          if (idx[u] < slice) {
            const int v = q * slice + idx[u];
            float x[N], pinv[N];
            unpack<T>(in[u], x);
            if (HAS_PINV) unpack<T>(pv[u], pinv);
#pragma unroll
            for (int e = 0; e < N; ++e) {
              const float y = HAS_PINV ? x[e] / pinv[e] : x[e];
              sq[0] = fmaf(y, y, sq[0]);
            }
            if (cur[u] < 0) cur[u] = seg_find(a.segs, v);
            while (__ldg(a.segs.seg_end + cur[u]) <= v) ++cur[u];
            g = __ldg(a.segs.seg_group + cur[u]);
          }
          accum.add(g, sq);
        }
      }
    }
    accum.flush_warp();
  }

  cta_barrier_peers(a, 0);                            // every rank's grads are ready

  const int base = a.rank * slice;
  Vec16* mc = static_cast<Vec16*>(a.mc_buf);
  int cur[U];
#pragma unroll
  for (int u = 0; u < U; ++u) cur[u] = -1;
  for (int it = 0; it < iters; ++it) {
    Vec16 in[U], pv[U];
    int idx[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      idx[u] = first + (it * U + u) * stride;
      if (idx[u] < slice) {
        in[u] = Multimem<T>::ld_reduce(mc + base + idx[u]);
        if (HAS_PINV) pv[u] = ld_vec(static_cast<const Vec16*>(a.pinv) + base + idx[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      float sq[2] = {0.f, 0.f};
      int g = -1;
      if (idx[u] < slice) {
        const int v = base + idx[u];
        float x[N], pinv[N];
        unpack<T>(in[u], x);
        if (HAS_PINV) unpack<T>(pv[u], pinv);
#pragma unroll
        for (int e = 0; e < N; ++e) {
          x[e] *= a.scale;
          const float y = HAS_PINV ? x[e] / pinv[e] : x[e];
          sq[1] = fmaf(y, y, sq[1]);
        }
        multimem_st(mc + v, pack<T>(x));
        if (cur[u] < 0) cur[u] = seg_find(a.segs, v);
        while (__ldg(a.segs.seg_end + cur[u]) <= v) ++cur[u];
        g = __ldg(a.segs.seg_group + cur[u]);
      }
      accum.add(g, sq);
    }
  }
  accum.flush_warp();
  double* outs[2] = {a.want_local ? a.L : nullptr, a.T};
  smem_stats_flush<2>(s_stats, a.n_groups, outs);

  cta_barrier_peers(a, 1);                            // every slice has landed everywhere
}

// ---------------------------------------------------------------------------
// local folds (gradient accumulation) and the single-replica pair norm
// ---------------------------------------------------------------------------
struct LocalArgs {
  void* g; void* a; void* pv; const void* pinv;
  int n_vec;
  SegTable segs;
  int n_groups;
  double* s0; double* s1; double* s2;   // statistic outputs (see kernels)
  int flag;                             // MODE 2: previous-step stash is valid
  const int* flag_ptr;                  // if non-null, overrides `flag` (device-resident state)
};

// MODE 0: fold_acc   (a += g ; s0 += |g|^2 ; g = 0)
// MODE 1: fold_final (s0 += |g|^2 ; g += a ; a = 0)
// MODE 2: pair       (s0 += |g|^2 ; if flag: s1 += |pv|^2, s2 += |(g+pv)/2|^2 ; pv = g)
template <typename T, int MODE, bool HAS_PINV>
__global__ void __launch_bounds__(ADL_THREADS, 2)
local_kernel(const LocalArgs a) {
  extern __shared__ double s_stats[];                 // [3][n_groups]
  constexpr int N = VecTraits<T>::N;
  constexpr int K = 3;
  smem_stats_zero(s_stats, K * a.n_groups);
  GroupAccum<K> accum;
  accum.init(s_stats, a.n_groups);
  const int stride = gridDim.x * blockDim.x;
  const int first = blockIdx.x * blockDim.x + threadIdx.x;
  const int iters = (a.n_vec + stride - 1) / stride;
  const bool have_prev = a.flag_ptr ? (*a.flag_ptr != 0) : (a.flag != 0);
  int cur = -1;
  for (int it = 0; it < iters; ++it) {
    const int v = first + it * stride;
    const bool active = v < a.n_vec;
    float sq[K] = {0.f, 0.f, 0.f};
    int grp = -1;
    if (active) {
      float g[N], o[N], pinv[N];
      const Vec16 gv = ld_vec(static_cast<const Vec16*>(a.g) + v);
      Vec16 ov;
      if (MODE == 2) ov = ld_vec(static_cast<const Vec16*>(a.pv) + v);
      else ov = ld_vec(static_cast<const Vec16*>(a.a) + v);
      if (HAS_PINV) {
        const Vec16 pvv = ld_vec(static_cast<const Vec16*>(a.pinv) + v);
        unpack<T>(pvv, pinv);
#pragma unroll
        for (int e = 0; e < N; ++e) pinv[e] = 1.0f / pinv[e];
      }
      unpack<T>(gv, g);
      unpack<T>(ov, o);
#pragma unroll
      for (int e = 0; e < N; ++e) {
        const float y = HAS_PINV ? g[e] * pinv[e] : g[e];
        sq[0] = fmaf(y, y, sq[0]);
      }
      if (MODE == 0) {
#pragma unroll
        for (int e = 0; e < N; ++e) o[e] += g[e];
        st_vec(static_cast<Vec16*>(a.a) + v, pack<T>(o));
        Vec16 z; z.w[0] = z.w[1] = z.w[2] = z.w[3] = 0u;
        st_vec(static_cast<Vec16*>(a.g) + v, z);
      } else if (MODE == 1) {
#pragma unroll
        for (int e = 0; e < N; ++e) g[e] += o[e];
        st_vec(static_cast<Vec16*>(a.g) + v, pack<T>(g));
        Vec16 z; z.w[0] = z.w[1] = z.w[2] = z.w[3] = 0u;
        st_vec(static_cast<Vec16*>(a.a) + v, z);
      } else {
        if (have_prev) {
#pragma unroll
          for (int e = 0; e < N; ++e) {
            const float p = HAS_PINV ? o[e] * pinv[e] : o[e];
            const float m = 0.5f * ((HAS_PINV ? g[e] * pinv[e] : g[e]) + p);
            sq[1] = fmaf(p, p, sq[1]);
            sq[2] = fmaf(m, m, sq[2]);
          }
        }
        st_vec(static_cast<Vec16*>(a.pv) + v, gv);
      }
      if (cur < 0) cur = seg_find(a.segs, v);
      while (__ldg(a.segs.seg_end + cur) <= v) ++cur;
      grp = __ldg(a.segs.seg_group + cur);
    }
    accum.add(grp, sq);
  }
  accum.flush_warp();
  double* outs[K] = {a.s0, a.s1, a.s2};
  smem_stats_flush<K>(s_stats, a.n_groups, outs);
}

// ---------------------------------------------------------------------------
// statistics exchange + gradient-noise-scale estimator + host mailbox
// ---------------------------------------------------------------------------
// Device-resident estimator state (doubles): see GNS_* offsets.
//   [0,G) sqr_biased  [G,2G) var_biased  [2G,3G) sqr_avg  [3G,4G) var_avg
//   4G+0 sqr_unbias  4G+1 var_unbias  4G+2 progress  4G+3 biased flag
// Host-written control block (doubles):
//   0 accum_scale  1 smoothing  2 rule id  3 rule arg (LEGW unit)  4 enabled
enum { GNS_SQR_UNBIAS = 0, GNS_VAR_UNBIAS = 1, GNS_PROGRESS = 2, GNS_BIASED = 3, GNS_TAIL = 8 };
enum { CTL_ACCUM_SCALE = 0, CTL_SMOOTHING = 1, CTL_RULE = 2, CTL_RULE_ARG = 3, CTL_ENABLED = 4 };
enum { RULE_ADASCALE = 0, RULE_ADAMSCALE = 1, RULE_LINEAR = 2, RULE_SQRT = 3, RULE_LEGW = 4 };
// Mailbox slot (doubles): header then payload.
//   0 seq  1 finite  2 gain  3 progress  4 sync_ns  5 err  6 scale  7 n_rows
//   host mode   : 8.. raw rows [n_rows][G]
//   device mode : 8.. sqr_avg[G], var_avg[G], lr_factor[G]
#define ADL_MBOX_HDR 8

struct FinalizeArgs {
  double* xchg[ADL_MAX_RANKS];     // every rank's exchange buffer [2][4*n_groups]
  uint32_t* pad[ADL_MAX_RANKS];
  int rank, world;
  uint32_t* step_ctr;              // device; bumped at the end of this kernel
  uint32_t site;
  int n_rows;                      // statistic rows in use (2, or 4 in pair mode)
  int n_groups;
  double* rows[4];                 // local partial vectors (device), reset after publish
  int sum_mask;                    // bit r set: row r is a per-rank partial to be summed
  int micro_steps;                 // k: backward passes folded into this step
  int pair_mode;                   // single replica, no accumulation
  int pair_flag;                   // host-mode: stash was valid (rows 2,3 meaningful)
  int* pair_state;                 // device-mode: stash validity (read, then updated)
  double* mailbox;                 // pinned host ring: [ring][slot_doubles]
  int ring, slot_doubles;
  double* result;                  // device copy of the summed rows (may be nullptr)
  unsigned long long* t_start;     // device: %globaltimer at end of local backward
  double* gns_state;               // device estimator state or nullptr (host mode)
  const double* gns_ctrl;          // host-written control block (device memory)
  float* lr_factor;                // [n_groups] out (device mode)
  uint32_t* err;
  unsigned long long timeout_ns;
};

__device__ __forceinline__ double block_sum(double x, double* scratch) {
  // blockDim.x == 256
  __syncthreads();
  scratch[threadIdx.x] = x;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) scratch[threadIdx.x] += scratch[threadIdx.x + o];
    __syncthreads();
  }
  return scratch[0];
}

__global__ void __launch_bounds__(256, 1) finalize_stats_kernel(const FinalizeArgs a) {
  __shared__ double scratch[256];
  const int G = a.n_groups;
  const int n = a.n_rows * G;
  const uint32_t step = *reinterpret_cast<volatile uint32_t*>(a.step_ctr);
  const int parity = step & 1;
  double* mine = a.xchg[a.rank] + (size_t)parity * 4 * G;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int r = i / G;
    mine[i] = a.rows[r][i - r * G];
  }
  if (a.world > 1) {
    __syncthreads();
    if ((int)threadIdx.x < a.world) {
      const int peer = threadIdx.x;
      const uint32_t epoch = step * ADL_SITES_PER_STEP + a.site;
      __threadfence_system();
      st_release_sys(pad_slot(a.pad[peer], 0, ADL_MAX_CTAS - 1, a.rank), epoch);
      wait_flag(pad_slot(a.pad[a.rank], 0, ADL_MAX_CTAS - 1, peer), epoch, a.timeout_ns, a.err);
    }
    __syncthreads();
  }
  double* slot = a.mailbox + (size_t)(step % a.ring) * a.slot_doubles;
  double* summed = a.result;                 // [4][G] device scratch (always provided)
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int r = i / G;
    double x;
    if (a.world > 1 && ((a.sum_mask >> r) & 1)) {
      x = 0.0;
      for (int p = 0; p < a.world; ++p)       // fixed order: identical on all ranks
        x += *reinterpret_cast<volatile double*>(a.xchg[p] + (size_t)parity * 4 * G + i);
    } else {
      x = mine[i];
    }
    summed[i] = x;
    a.rows[r][i - r * G] = 0.0;               // partials restart from zero
  }
  __syncthreads();

  const bool device_mode = a.gns_state != nullptr && a.gns_ctrl[CTL_ENABLED] != 0.0;
  double finite_flag = 1.0, gain = 1.0, progress = 0.0, scale_out = 0.0;
  if (!device_mode) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) slot[ADL_MBOX_HDR + i] = summed[i];
  } else {
    double* st = a.gns_state;
    double* tail = st + 4 * G;
    const double accum_scale = a.gns_ctrl[CTL_ACCUM_SCALE];
    const double smoothing = a.gns_ctrl[CTL_SMOOTHING];
    const int rule = (int)a.gns_ctrl[CTL_RULE];
    const double* L = summed;
    const double* T = summed + G;
    // non-finite gradients: skip the statistics update (and the progress)
    double bad = 0.0;
    for (int g = threadIdx.x; g < G; g += blockDim.x)
      if (!isfinite(T[g]) || !isfinite(L[g])) bad += 1.0;
    bad = block_sum(bad, scratch);
    const bool finite = bad == 0.0;
    finite_flag = finite ? 1.0 : 0.0;
    int count = a.world * a.micro_steps;
    double scale = accum_scale * a.micro_steps;
    const double lr_scale = scale;            // ScalingRuleBase.step uses accum_scale * k
    const bool had_stash = a.pair_mode && a.pair_state && (*a.pair_state != 0);
    bool update = finite;
    bool was_biased = tail[GNS_BIASED] != 0.0;
    __syncthreads();
    if (finite) {
      if (count > 1) {
        if (was_biased) {                     // biased -> unbiased: restart the averages
          for (int g = threadIdx.x; g < G; g += blockDim.x) { st[g] = 0.0; st[G + g] = 0.0; }
          __syncthreads();
          if (threadIdx.x == 0) { tail[GNS_SQR_UNBIAS] = 0.0; tail[GNS_VAR_UNBIAS] = 0.0; }
        }
        if (threadIdx.x == 0) tail[GNS_BIASED] = 0.0;
      } else {
        if (threadIdx.x == 0) tail[GNS_BIASED] = 1.0;
        if (!had_stash) update = false;       // first sample: nothing to difference yet
      }
    }
    __syncthreads();
    if (update) {
      const double theta_scale = (count > 1) ? scale : 2.0 * accum_scale;
      const double theta = pow(smoothing, theta_scale);
      const double su = theta * tail[GNS_SQR_UNBIAS] + (1.0 - theta);
      const double vu = theta * tail[GNS_VAR_UNBIAS] + (1.0 - theta);
      for (int g = threadIdx.x; g < G; g += blockDim.x) {
        double local, total, cnt, sc;
        if (count > 1) {
          local = L[g] / count; total = T[g]; cnt = count; sc = scale;
        } else {
          const double* Pp = summed + 2 * G;
          const double* Pa = summed + 3 * G;
          local = 0.5 * (Pp[g] + T[g]); total = Pa[g]; cnt = 2.0; sc = 2.0 * accum_scale;
        }
        const double grad_sqr = (cnt * total - local) / (cnt - 1.0);
        const double grad_var = (local - total) * sc / (cnt - 1.0);
        const double sb = theta * st[g] + (1.0 - theta) * grad_sqr;
        const double vb = theta * st[G + g] + (1.0 - theta) * grad_var;
        st[g] = sb; st[G + g] = vb;
        st[2 * G + g] = sb / su; st[3 * G + g] = vb / vu;
      }
      __syncthreads();
      if (threadIdx.x == 0) { tail[GNS_SQR_UNBIAS] = su; tail[GNS_VAR_UNBIAS] = vu; }
    }
    __syncthreads();
    if (a.pair_state && threadIdx.x == 0)
      *a.pair_state = (a.pair_mode && finite) ? 1 : 0;
    // learning-rate factors and gain from the (possibly just updated) averages
    double sqr_sum = 0.0, var_sum = 0.0;
    for (int g = threadIdx.x; g < G; g += blockDim.x) {
      const double var = fmax(st[3 * G + g], 1e-6);
      const double sqr = fmax(st[2 * G + g], 0.0);
      sqr_sum += sqr; var_sum += var;
    }
    sqr_sum = block_sum(sqr_sum, scratch);
    var_sum = block_sum(var_sum, scratch);
    gain = (var_sum + sqr_sum) / (var_sum / lr_scale + sqr_sum);
    progress = tail[GNS_PROGRESS];
    for (int g = threadIdx.x; g < G; g += blockDim.x) {
      const double var = fmax(st[3 * G + g], 1e-6);
      const double sqr = fmax(st[2 * G + g], 0.0);
      const double ada = (var + sqr) / (var / lr_scale + sqr);
      double f;
      if (rule == RULE_ADASCALE) f = ada;
      else if (rule == RULE_ADAMSCALE) f = sqrt(ada);
      else if (rule == RULE_LINEAR) f = lr_scale;
      else if (rule == RULE_SQRT) f = sqrt(lr_scale);
      else {                                  // LEGW: sqrt(scale) with progress warm-up
        const double total_steps = a.gns_ctrl[CTL_RULE_ARG] * lr_scale;
        f = sqrt(lr_scale) * ((progress < total_steps) ? progress / total_steps : 1.0);
      }
      a.lr_factor[g] = (float)f;
      slot[ADL_MBOX_HDR + g] = st[2 * G + g];
      slot[ADL_MBOX_HDR + G + g] = st[3 * G + g];
      slot[ADL_MBOX_HDR + 2 * G + g] = f;
    }
    __syncthreads();
    if (threadIdx.x == 0 && finite) {         // progress advances with every update
      progress += gain;
      tail[GNS_PROGRESS] = progress;
    }
    scale_out = lr_scale;
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned long long now = globaltimer_ns();
    const unsigned long long t0 = a.t_start ? *a.t_start : now;
    slot[1] = finite_flag;
    slot[2] = gain;
    slot[3] = progress;
    slot[4] = (double)(now > t0 ? now - t0 : 0ull);
    slot[5] = (double)(*a.err);
    slot[6] = scale_out;
    slot[7] = (double)a.n_rows;
    __threadfence_system();
    *reinterpret_cast<volatile double*>(slot) = (double)(step + 1);   // publish last
    *a.step_ctr = step + 1;                                           // next optimizer step
  }
}

__global__ void stamp_kernel(unsigned long long* dst) { *dst = globaltimer_ns(); }

// ---------------------------------------------------------------------------
// broadcast: every non-source rank pulls the source's staging buffer
// ---------------------------------------------------------------------------
struct BcastArgs {
  void* staging[ADL_MAX_RANKS];
  uint32_t* pad[ADL_MAX_RANKS];
  int rank, world, src;
  const uint32_t* step_ctr;
  uint32_t site;
  void* dst;                       // local destination (may equal staging[rank])
  long long n_vec;
  uint32_t* err;
  unsigned long long timeout_ns;
};

__global__ void __launch_bounds__(ADL_THREADS, 1) bcast_pull_kernel(const BcastArgs a) {
  ReduceArgs b;   // reuse the barrier helper
#pragma unroll
  for (int p = 0; p < ADL_MAX_RANKS; ++p) b.pad[p] = a.pad[p];
  b.rank = a.rank; b.world = a.world; b.step_ctr = a.step_ctr; b.site = a.site; b.err = a.err; b.timeout_ns = a.timeout_ns;
  cta_barrier_peers(b, 0);                            // source staging is complete
  if (a.rank != a.src || a.dst != a.staging[a.rank]) {
    const Vec16* src = static_cast<const Vec16*>(a.staging[a.src]);
    Vec16* dst = static_cast<Vec16*>(a.dst);
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < a.n_vec; v += 2 * stride) {
      const Vec16 x0 = ld_vec(src + v);
      Vec16 x1;
      const bool two = v + stride < a.n_vec;
      if (two) x1 = ld_vec(src + v + stride);
      st_vec(dst + v, x0);
      if (two) st_vec(dst + v + stride, x1);
    }
  }
  cta_barrier_peers(b, 1);                            // source may reuse its staging
}

// ===========================================================================
// C ABI
// ===========================================================================
#define ADL_CHECK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return (int)e_; } while (0)

static int g_device = -1;

// Raise the dynamic shared memory limit of every statistics kernel ONCE (not
// per launch: launches may happen under CUDA-graph capture).
template <typename T>
static int set_attrs_for() {
  const int lim = ADL_MAX_STAT_SMEM;
#define ADL_AR_ATTR(W)                                                                                          \
  ADL_CHECK(cudaFuncSetAttribute(allreduce_gns_kernel<T, W, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim)); \
  ADL_CHECK(cudaFuncSetAttribute(allreduce_gns_kernel<T, W, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
  ADL_AR_ATTR(0) ADL_AR_ATTR(1) ADL_AR_ATTR(2) ADL_AR_ATTR(4) ADL_AR_ATTR(8)
#undef ADL_AR_ATTR
  ADL_CHECK(cudaFuncSetAttribute(allreduce_nvls_kernel<T, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
  ADL_CHECK(cudaFuncSetAttribute(allreduce_nvls_kernel<T, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
  ADL_CHECK(cudaFuncSetAttribute(local_kernel<T, 0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
  ADL_CHECK(cudaFuncSetAttribute(local_kernel<T, 0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
  ADL_CHECK(cudaFuncSetAttribute(local_kernel<T, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
  ADL_CHECK(cudaFuncSetAttribute(local_kernel<T, 1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
  ADL_CHECK(cudaFuncSetAttribute(local_kernel<T, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
  ADL_CHECK(cudaFuncSetAttribute(local_kernel<T, 2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
  return 0;
}

extern "C" {

int adl_set_device(int dev) {
  g_device = dev;
  ADL_CHECK(cudaSetDevice(dev));
  if (int rc = set_attrs_for<float>()) return rc;
  if (int rc = set_attrs_for<__nv_bfloat16>()) return rc;
  if (int rc = set_attrs_for<__half>()) return rc;
  return 0;
}

int adl_bind_thread() {
  if (g_device >= 0) ADL_CHECK(cudaSetDevice(g_device));
  return 0;
}

int adl_max_groups() { return ADL_MAX_STAT_SMEM / (3 * (int)sizeof(double)); }

const char* adl_error_string(int code) { return cudaGetErrorString((cudaError_t)code); }

int adl_sm_count(int dev) {
  int n = 0;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
  return n;
}

// dtype: 0 = fp32, 1 = bf16, 2 = fp16
int adl_allreduce_gns(const ReduceArgs* args, int dtype, int grid, void* stream) {
  if (g_device >= 0) ADL_CHECK(cudaSetDevice(g_device));
  const size_t smem = sizeof(double) * 2 * args->n_groups;
  if (smem > ADL_MAX_STAT_SMEM) return -4;
  const bool pinv = args->pinv != nullptr;
  cudaStream_t s = (cudaStream_t)stream;
#define LAUNCH_AR_W(T, P, W) allreduce_gns_kernel<T, W, P><<<grid, ADL_THREADS, smem, s>>>(*args)
#define LAUNCH_AR(T, P)                                        \
  do {                                                         \
    switch (args->world) {                                     \
      case 1: LAUNCH_AR_W(T, P, 1); break;                     \
      case 2: LAUNCH_AR_W(T, P, 2); break;                     \
      case 4: LAUNCH_AR_W(T, P, 4); break;                     \
      case 8: LAUNCH_AR_W(T, P, 8); break;                     \
      default: LAUNCH_AR_W(T, P, 0); break;                    \
    }                                                          \
  } while (0)
#define LAUNCH_NVLS(T, P) allreduce_nvls_kernel<T, P><<<grid, ADL_THREADS, smem, s>>>(*args)
  if (args->mc_buf != nullptr && args->world > 1) {
    if (dtype == 0) { if (pinv) LAUNCH_NVLS(float, true); else LAUNCH_NVLS(float, false); }
    else if (dtype == 1) { if (pinv) LAUNCH_NVLS(__nv_bfloat16, true); else LAUNCH_NVLS(__nv_bfloat16, false); }
    else if (dtype == 2) { if (pinv) LAUNCH_NVLS(__half, true); else LAUNCH_NVLS(__half, false); }
    else return -2;
    return (int)cudaGetLastError();
  }
#undef LAUNCH_NVLS
  if (dtype == 0) { if (pinv) LAUNCH_AR(float, true); else LAUNCH_AR(float, false); }
  else if (dtype == 1) { if (pinv) LAUNCH_AR(__nv_bfloat16, true); else LAUNCH_AR(__nv_bfloat16, false); }
  else if (dtype == 2) { if (pinv) LAUNCH_AR(__half, true); else LAUNCH_AR(__half, false); }
  else return -2;
#undef LAUNCH_AR
#undef LAUNCH_AR_W
  return (int)cudaGetLastError();
}

// mode: 0 fold_acc, 1 fold_final, 2 pair
int adl_local(const LocalArgs* args, int mode, int dtype, int grid, void* stream) {
  if (g_device >= 0) ADL_CHECK(cudaSetDevice(g_device));
  const size_t smem = sizeof(double) * 3 * args->n_groups;
  if (smem > ADL_MAX_STAT_SMEM) return -4;
  const bool pinv = args->pinv != nullptr;
  cudaStream_t s = (cudaStream_t)stream;
#define LAUNCH_L(T, M, P) local_kernel<T, M, P><<<grid, ADL_THREADS, smem, s>>>(*args)
#define LAUNCH_LM(T)                                                                         \
  do {                                                                                       \
    if (mode == 0) { if (pinv) LAUNCH_L(T, 0, true); else LAUNCH_L(T, 0, false); }           \
    else if (mode == 1) { if (pinv) LAUNCH_L(T, 1, true); else LAUNCH_L(T, 1, false); }      \
    else if (mode == 2) { if (pinv) LAUNCH_L(T, 2, true); else LAUNCH_L(T, 2, false); }      \
    else return -3;                                                                          \
  } while (0)
  if (dtype == 0) LAUNCH_LM(float);
  else if (dtype == 1) LAUNCH_LM(__nv_bfloat16);
  else if (dtype == 2) LAUNCH_LM(__half);
  else return -2;
#undef LAUNCH_LM
#undef LAUNCH_L
  return (int)cudaGetLastError();
}

int adl_finalize_stats(const FinalizeArgs* args, void* stream) {
  if (g_device >= 0) ADL_CHECK(cudaSetDevice(g_device));
  finalize_stats_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(*args);
  return (int)cudaGetLastError();
}

int adl_stamp(unsigned long long* dst, void* stream) {
  if (g_device >= 0) ADL_CHECK(cudaSetDevice(g_device));
  stamp_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(dst);
  return (int)cudaGetLastError();
}

int adl_bcast_pull(const BcastArgs* args, int grid, void* stream) {
  if (g_device >= 0) ADL_CHECK(cudaSetDevice(g_device));
  bcast_pull_kernel<<<grid, ADL_THREADS, 0, (cudaStream_t)stream>>>(*args);
  return (int)cudaGetLastError();
}

int adl_sizeof_reduce_args() { return (int)sizeof(ReduceArgs); }
int adl_sizeof_local_args() { return (int)sizeof(LocalArgs); }
int adl_sizeof_finalize_args() { return (int)sizeof(FinalizeArgs); }
int adl_sizeof_bcast_args() { return (int)sizeof(BcastArgs); }

}  // extern "C"
