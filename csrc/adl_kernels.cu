// adaptdl_b200 -- sm_100a gradient kernels (C ABI, launched from Python via
// ctypes on torch's CUDA streams).
//
//   adl_fold_acc       A += G ; L += |G/P|^2 ; G = 0
//   adl_fold_final     L += |G/P|^2 ; G += A ; A = 0
//   adl_allreduce_gns  fused all-reduce over NVLink peer mappings:
//                      G <- s * sum_r G_r (in place on every rank) with
//                      L += sum_r |G_r/P|^2 and T += |G/P|^2 from the same
//                      registers (reference call sites K1-K6).
//                      Three flavours, picked per bucket by the host:
//                        two-shot P2P  (rank r reduces slice r from the peers'
//                                       arenas and stores it back to all of them)
//                        one-shot push (small buckets: every rank pushes its
//                                       bucket into the peers' staging lanes,
//                                       ONE flag round, local reduction)
//                        NVLS          (multimem.ld_reduce / multimem.st: the
//                                       NVSwitch reduces and multicasts)
//                      The kernel of the LAST bucket of a step also runs the
//                      statistics exchange + estimator (finalize) in its last
//                      CTA: one launch and one peer barrier fewer per step.
//   adl_pair_norm      T=|G/P|^2, Pp=|Pv/P|^2, Pa=|(G+Pv)/2P|^2 ; Pv = G
//   adl_finalize_stats sum per-rank partial statistics over ranks through a
//                      peer-mapped pad, run the gradient-noise-scale estimator,
//                      publish to a pinned host mailbox with %globaltimer
//                      stamps (step and sync durations, max over ranks), reset
//                      the partials
//   adl_step_mark      %globaltimer interval of an accumulation micro-step
//   adl_bcast_pull     rank src's staging buffer -> every rank
//   adl_stamp          write %globaltimer to device memory
//
// No tensor cores: these are bandwidth / latency kernels. sm_100a specifics:
// 128-bit L1-bypassing vector accesses, .sys-scope release/acquire flags on
// NVLink-mapped signal pads, multimem.* through the NVSwitch, %globaltimer
// stamps, grids sized to leave SMs to the concurrently running backward pass.
//
// Synchronisation protocol of one optimizer step (all on the comm stream):
//   * every bucket kernel has ONE peer barrier, at its start ("this rank's
//     gradients of the bucket are final"); its own remote stores are fenced
//     (fence.sys) before the kernel ends but nobody waits for them there;
//   * the finalize (stand-alone kernel, or the last CTA of the last bucket
//     kernel) has the step's only other barrier: a rank sends its flag after
//     ALL its bucket kernels have completed, so passing it means every peer's
//     stores have landed in this rank's arenas AND every peer has finished
//     reading this rank's arenas -- the optimizer may read the gradients and
//     the next backward may overwrite them.
#include "adl_common.cuh"

#include <stdio.h>
#include <string.h>

// error word bits (device -> host, sticky)
#define ADL_ERR_TIMEOUT 1u
#define ADL_MAX_STAT_SMEM (96 * 1024)

// preconditioner modes (template parameter PINV)
//   0 none
//   1 `pinv` is a flat element-wise divisor in the gradient dtype (host path)
//   2 `pinv` are Adam's second moments (exp_avg_sq) in the arena layout, fp32
//     when `pinv_wide` (16-bit gradients with fp32 optimizer state) else in
//     the gradient dtype; divisor = sqrt(v) * coef[g][0] + coef[g][1], and
//     coef[g][0] == 0 switches preconditioning off (Adam warm-up).
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
  const void* pinv;                // see the PINV modes above (or nullptr)
  double* L;                       // [n_groups] partial: sum_r |G_r/P|^2
  double* T;                       // [n_groups] partial: |G/P|^2
  uint32_t* err;                   // sticky error word (device)
  unsigned long long timeout_ns;
  void* mc_buf;                    // NVLS: multicast address of the bucket (or nullptr)
  int pinv_mode;
  int pinv_wide;
  const float* pinv_coef;          // mode 2: [n_groups][2]
  void* stage[ADL_MAX_RANKS];      // one-shot: every rank's staging area of this bucket
                                   // ([world][n_vec] vectors, lane r written by rank r)
  int fuse_fin;                    // the last CTA to finish runs the finalize
  uint32_t* ticket;                // device: CTA completion counter (fuse_fin)
};

__device__ __forceinline__ bool wait_flag(const uint32_t* p, uint32_t epoch,
                                          unsigned long long timeout_ns, uint32_t* err) {
  if ((int32_t)(ld_acquire_sys(p) - epoch) >= 0) return true;
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
// the finalize once per optimizer step) * ADL_SITES_PER_STEP + the launch's
// ordinal within the step. Nothing launch-specific is baked into kernel
// arguments, so a captured CUDA graph can be replayed step after step.
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
// element-wise reciprocal preconditioner of one 16-byte gradient vector
// ---------------------------------------------------------------------------
// `v` is the vector index relative to the bucket, `g` its statistics group.
template <typename T, int PINV>
struct Precond {
  static constexpr int N = VecTraits<T>::N;
  Vec16 raw[2];
  __device__ __forceinline__ void issue(const void* pinv, int wide, int v) {
    if (PINV == 1) {
      raw[0] = ld_vec(static_cast<const Vec16*>(pinv) + v);
    } else if (PINV == 2) {
      if (wide && N == 8) {
        raw[0] = ld_vec(static_cast<const Vec16*>(pinv) + 2 * (size_t)v);
        raw[1] = ld_vec(static_cast<const Vec16*>(pinv) + 2 * (size_t)v + 1);
      } else {
        raw[0] = ld_vec(static_cast<const Vec16*>(pinv) + v);
      }
    }
  }
  // out[e] = 1 / divisor
  __device__ __forceinline__ void finish(const float* coef, int wide, int g, float* out) const {
    if (PINV == 1) {
      unpack<T>(raw[0], out);
#pragma unroll
      for (int e = 0; e < N; ++e) out[e] = 1.0f / out[e];
    } else if (PINV == 2) {
      const float c = (g >= 0) ? __ldg(coef + 2 * g) : 0.f;
      if (c == 0.f) {
#pragma unroll
        for (int e = 0; e < N; ++e) out[e] = 1.0f;
        return;
      }
      const float eps = __ldg(coef + 2 * g + 1);
      if (wide && N == 8) {
        unpack<float>(raw[0], out);
        unpack<float>(raw[1], out + 4);
      } else {
        unpack<T>(raw[0], out);
      }
#pragma unroll
      for (int e = 0; e < N; ++e) out[e] = 1.0f / fmaf(sqrtf(out[e]), c, eps);
    }
  }
};

// The bucket's segment table, staged in shared memory once per CTA (one coalesced round trip)
// when it fits: the per-thread binary search and cursor advance then cost shared-memory
// latency instead of a chain of dependent L2 round trips at the start of every kernel.
#define ADL_SEG_SMEM 512
struct SegCache {
  const int* end;
  const int* group;
  int n_seg;
  __device__ __forceinline__ void load(const SegTable& t, int* s_end, int* s_group) {
    n_seg = t.n_seg;
    if (t.n_seg <= ADL_SEG_SMEM) {
      for (int i = threadIdx.x; i < t.n_seg; i += blockDim.x) {
        s_end[i] = __ldg(t.seg_end + i);
        s_group[i] = __ldg(t.seg_group + i);
      }
      end = s_end;
      group = s_group;
    } else {
      end = t.seg_end;
      group = t.seg_group;
    }
    __syncthreads();
  }
};

// statistics group of bucket-relative vector v (monotone cursor per lane)
__device__ __forceinline__ int group_of(const SegCache& segs, int v, int& cur) {
  if (cur < 0) {
    int lo = 0, hi = segs.n_seg - 1;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (segs.end[mid] > v) hi = mid; else lo = mid + 1;
    }
    cur = lo;
  }
  while (segs.end[cur] <= v) ++cur;
  return segs.group[cur];
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
//   8 step_ns  9 accum_ns  10 accum_count  11 amp_scale  12..15 reserved
//   host mode   : 16.. raw rows [n_rows][G]
//   device mode : 16.. sqr_avg[G], var_avg[G], lr_factor[G]
// step_ns  = max over ranks of the interval between this finalize and the
//            previous step mark (0 right after a clock reset),
// sync_ns  = max over ranks of (finalize entry - end of the local backward)
//            + this rank's time inside the finalize,
// accum_ns / accum_count = accumulation micro-steps since the previous
//            finalize (adl_step_mark); accum_ns is the max over ranks.
#define ADL_MBOX_HDR 16
// per-parity exchange record: [4][G] statistic rows + ADL_XCHG_TAIL timing doubles
#define ADL_XCHG_TAIL 4
// step clock (device doubles): 0 accum_ns  1 accum_count
#define ADL_CLOCK_DOUBLES 4

struct FinalizeArgs {
  double* xchg[ADL_MAX_RANKS];     // every rank's exchange buffer [2][4*n_groups + ADL_XCHG_TAIL]
  uint32_t* pad[ADL_MAX_RANKS];
  int rank, world;
  uint32_t* step_ctr;              // device; bumped at the end of the finalize
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
  double* result;                  // device copy of the summed rows
  unsigned long long* t_start;     // device: %globaltimer at end of local backward
  double* gns_state;               // device estimator state or nullptr (host mode)
  const double* gns_ctrl;          // host-written control block (device memory)
  float* lr_factor;                // [n_groups + 1] out (device mode); last = finite flag
  uint32_t* err;
  unsigned long long timeout_ns;
  unsigned long long* last_stamp;  // device: %globaltimer of the previous step mark (0 = none)
  double* clock;                   // device step clock (ADL_CLOCK_DOUBLES)
  const float* amp_scale;          // device: AMP loss scale the gradients carry (or nullptr)
};

// Sum of (x, y) over the CTA (blockDim.x a multiple of 32, <= ADL_THREADS): warp shuffles, then
// one shared-memory stage. All threads call it and get the totals.
__device__ __forceinline__ void block_sum2(double& x, double& y, double* scratch) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    x += __shfl_xor_sync(0xffffffffu, x, o);
    y += __shfl_xor_sync(0xffffffffu, y, o);
  }
  const int warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
  __syncthreads();                            // scratch may still be read from a previous call
  if ((threadIdx.x & 31) == 0) { scratch[2 * warp] = x; scratch[2 * warp + 1] = y; }
  __syncthreads();
  double sx = 0.0, sy = 0.0;
  for (int w = 0; w < nwarp; ++w) { sx += scratch[2 * w]; sy += scratch[2 * w + 1]; }
  x = sx; y = sy;
}

// Executed by ONE CTA (all of its threads) once every bucket kernel of the step has completed on
// this rank. It sits on the step's critical path (the optimizer waits for it), so it is written
// against latency: every scalar and every per-group state word it will need is requested up
// front, in parallel and BEFORE the peer barrier; after the barrier there is one round of
// (remote) loads for the sums, then arithmetic on registers only, two block reductions, and one
// system fence before the mailbox sequence number.
__device__ __noinline__ void finalize_body(const FinalizeArgs& a) {
  __shared__ double scratch[2 * (ADL_THREADS / 32)];
  __shared__ double sh_c[24];                // preloaded scalars (see the enum below)
  __shared__ double sh_t[3];                 // step_ns, sync_entry_ns, accum_ns (max over ranks)
  __shared__ unsigned long long sh_entry;
  enum { C_ACCUM_SCALE = 0, C_SMOOTHING, C_RULE, C_RULE_ARG, C_ENABLED, C_AMP, C_SQR_UNBIAS, C_VAR_UNBIAS,
         C_PROGRESS, C_BIASED, C_PAIR_STATE, C_ERR, C_N };
  const int G = a.n_groups;
  const int n = a.n_rows * G;
  const int XS = 4 * G + ADL_XCHG_TAIL;
  const bool has_state = a.gns_state != nullptr;
  // ---- requests that do not depend on anything (one round trip, all in flight together) ----
  if (threadIdx.x < C_N) {
    double v = 0.0;
    const int c = threadIdx.x;
    if (c <= C_ENABLED) v = (has_state && a.gns_ctrl) ? a.gns_ctrl[c] : 0.0;
    else if (c == C_AMP) v = a.amp_scale ? (double)(*a.amp_scale) : 1.0;
    else if (c >= C_SQR_UNBIAS && c <= C_BIASED) v = has_state ? a.gns_state[4 * G + (c - C_SQR_UNBIAS)] : 0.0;
    else if (c == C_PAIR_STATE) v = a.pair_state ? (double)(*a.pair_state) : 0.0;
    else if (c == C_ERR) v = (double)(*a.err);
    sh_c[c] = v;
  }
  const uint32_t step = *reinterpret_cast<volatile uint32_t*>(a.step_ctr);
  const int parity = step & 1;
  double* mine = a.xchg[a.rank] + (size_t)parity * XS;
  // per-group estimator state of "my" groups (g = threadIdx.x, + blockDim.x, ...): at most
  // GPT groups per thread live in registers, the rest (huge group counts) is re-read later
  constexpr int GPT = 4;
  double st_sb[GPT], st_vb[GPT], st_sa[GPT], st_va[GPT];
#pragma unroll
  for (int k = 0; k < GPT; ++k) {
    const int g = threadIdx.x + k * blockDim.x;
    if (has_state && g < G) {
      st_sb[k] = a.gns_state[g]; st_vb[k] = a.gns_state[G + g];
      st_sa[k] = a.gns_state[2 * G + g]; st_va[k] = a.gns_state[3 * G + g];
    } else { st_sb[k] = st_vb[k] = st_sa[k] = st_va[k] = 0.0; }
  }
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int r = i / G;
    mine[i] = __ldcg(a.rows[r] + (i - r * G));
  }
  if (threadIdx.x == 0) {
    const unsigned long long now = globaltimer_ns();
    const unsigned long long last = a.last_stamp ? *a.last_stamp : 0ull;
    const unsigned long long t0 = a.t_start ? *a.t_start : now;
    double accum_ns = 0.0, accum_cnt = 0.0;
    if (a.clock) { accum_ns = a.clock[0]; accum_cnt = a.clock[1]; a.clock[0] = 0.0; a.clock[1] = 0.0; }
    if (a.last_stamp) *a.last_stamp = now;
    mine[4 * G + 0] = (last != 0ull && now > last) ? (double)(now - last) : 0.0;
    mine[4 * G + 1] = now > t0 ? (double)(now - t0) : 0.0;
    mine[4 * G + 2] = accum_ns;
    mine[4 * G + 3] = accum_cnt;
    sh_entry = now;
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
  }
  __syncthreads();
  double* slot = a.mailbox + (size_t)(step % a.ring) * a.slot_doubles;
  double* summed = a.result;                 // [4][G] device copy of the summed rows
  // ---- the sums: one round of loads over the peers' exchange records ----
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int r = i / G;
    double x;
    if (a.world > 1 && ((a.sum_mask >> r) & 1)) {
      double part[ADL_MAX_RANKS];
#pragma unroll
      for (int p = 0; p < ADL_MAX_RANKS; ++p)
        part[p] = (p < a.world) ? *reinterpret_cast<volatile double*>(a.xchg[p] + (size_t)parity * XS + i) : 0.0;
      x = 0.0;
#pragma unroll
      for (int p = 0; p < ADL_MAX_RANKS; ++p) x += part[p];      // fixed order: identical on all ranks
    } else {
      x = mine[i];
    }
    summed[i] = x;
    a.rows[r][i - r * G] = 0.0;               // partials restart from zero
  }
  if (threadIdx.x < 3) {                      // timings: max over ranks
    double x = mine[4 * G + threadIdx.x];
    for (int p = 0; p < a.world; ++p) {
      if (p == a.rank) continue;
      const double y = *reinterpret_cast<volatile double*>(
          a.xchg[p] + (size_t)parity * XS + 4 * G + threadIdx.x);
      x = fmax(x, y);
    }
    sh_t[threadIdx.x] = x;
  }
  __syncthreads();

  const bool device_mode = has_state && sh_c[C_ENABLED] != 0.0;
  double finite_flag = 1.0, gain = 1.0, progress = 0.0, scale_out = 0.0;
  const double amp = sh_c[C_AMP];
  if (!device_mode) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) slot[ADL_MBOX_HDR + i] = summed[i];
  } else {
    double* st = a.gns_state;
    double* tail = st + 4 * G;
    const double accum_scale = sh_c[C_ACCUM_SCALE];
    const double smoothing = sh_c[C_SMOOTHING];
    const int rule = (int)sh_c[C_RULE];
    const double inv_amp2 = 1.0 / (amp * amp);  // the statistics are of amp-scaled gradients
    const double* L = summed;
    const double* T = summed + G;
    const int count = a.world * a.micro_steps;
    const double scale = accum_scale * a.micro_steps;
    const double lr_scale = scale;            // ScalingRuleBase.step uses accum_scale * k
    const bool had_stash = a.pair_mode && sh_c[C_PAIR_STATE] != 0.0;
    const bool was_biased = sh_c[C_BIASED] != 0.0;
    // candidate update of my groups, computed unconditionally (committed once `finite` is known)
    const bool restart = count > 1 && was_biased;     // biased -> unbiased: the averages restart
    const double theta = pow(smoothing, (count > 1) ? scale : 2.0 * accum_scale);
    const double su = theta * (restart ? 0.0 : sh_c[C_SQR_UNBIAS]) + (1.0 - theta);
    const double vu = theta * (restart ? 0.0 : sh_c[C_VAR_UNBIAS]) + (1.0 - theta);
    double bad = 0.0, unused = 0.0;
    double nsb[GPT], nvb[GPT];
#pragma unroll
    for (int k = 0; k < GPT; ++k) { nsb[k] = 0.0; nvb[k] = 0.0; }
    for (int g = threadIdx.x, k = 0; g < G; g += blockDim.x, ++k) {
      const double Lg = L[g], Tg = T[g];
      if (!isfinite(Tg) || !isfinite(Lg)) bad += 1.0;
      double local, total, cnt, sc;
      if (count > 1) {
        local = Lg * inv_amp2 / count; total = Tg * inv_amp2; cnt = count; sc = scale;
      } else {
        local = 0.5 * (summed[2 * G + g] + Tg) * inv_amp2; total = summed[3 * G + g] * inv_amp2;
        cnt = 2.0; sc = 2.0 * accum_scale;
      }
      const double grad_sqr = (cnt * total - local) / (cnt - 1.0);
      const double grad_var = (local - total) * sc / (cnt - 1.0);
      const double sb0 = (k < GPT) ? st_sb[k] : st[g];
      const double vb0 = (k < GPT) ? st_vb[k] : st[G + g];
      const double sb = theta * (restart ? 0.0 : sb0) + (1.0 - theta) * grad_sqr;
      const double vb = theta * (restart ? 0.0 : vb0) + (1.0 - theta) * grad_var;
      if (k < GPT) { nsb[k] = sb; nvb[k] = vb; }
      else { summed[2 * G + g] = sb; summed[3 * G + g] = vb; }   // spill (only with > GPT groups per thread)
    }
    block_sum2(bad, unused, scratch);
    const bool finite = bad == 0.0;
    finite_flag = finite ? 1.0 : 0.0;
    // non-finite gradients: skip the statistics update (and the progress); a single sample
    // without a stash has nothing to difference yet
    const bool update = finite && (count > 1 || had_stash);
    double sqr_sum = 0.0, var_sum = 0.0;
    for (int g = threadIdx.x, k = 0; g < G; g += blockDim.x, ++k) {
      double sa, va;
      if (update) {
        const double sb = (k < GPT) ? nsb[k] : summed[2 * G + g];
        const double vb = (k < GPT) ? nvb[k] : summed[3 * G + g];
        st[g] = sb; st[G + g] = vb;
        sa = sb / su; va = vb / vu;
        st[2 * G + g] = sa; st[3 * G + g] = va;
      } else {
        if (finite && restart) { st[g] = 0.0; st[G + g] = 0.0; }
        sa = (k < GPT) ? st_sa[k] : st[2 * G + g];
        va = (k < GPT) ? st_va[k] : st[3 * G + g];
      }
      if (k < GPT) { st_sa[k] = sa; st_va[k] = va; }
      sqr_sum += fmax(sa, 0.0);
      var_sum += fmax(va, 1e-6);
    }
    block_sum2(sqr_sum, var_sum, scratch);
    gain = (var_sum + sqr_sum) / (var_sum / lr_scale + sqr_sum);
    progress = sh_c[C_PROGRESS];
    const double rule_arg = sh_c[C_RULE_ARG];
    for (int g = threadIdx.x, k = 0; g < G; g += blockDim.x, ++k) {
      const double sa = (k < GPT) ? st_sa[k] : st[2 * G + g];
      const double va = (k < GPT) ? st_va[k] : st[3 * G + g];
      const double var = fmax(va, 1e-6);
      const double sqr = fmax(sa, 0.0);
      const double ada = (var + sqr) / (var / lr_scale + sqr);
      double fct;
      if (rule == RULE_ADASCALE) fct = ada;
      else if (rule == RULE_ADAMSCALE) fct = sqrt(ada);
      else if (rule == RULE_LINEAR) fct = lr_scale;
      else if (rule == RULE_SQRT) fct = sqrt(lr_scale);
      else {                                  // LEGW: sqrt(scale) with progress warm-up
        const double total_steps = rule_arg * lr_scale;
        fct = sqrt(lr_scale) * ((progress < total_steps) ? progress / total_steps : 1.0);
      }
      a.lr_factor[g] = (float)fct;
      slot[ADL_MBOX_HDR + g] = sa;
      slot[ADL_MBOX_HDR + G + g] = va;
      slot[ADL_MBOX_HDR + 2 * G + g] = fct;
    }
    if (threadIdx.x == 0) {
      a.lr_factor[G] = finite ? 1.f : 0.f;    // the fused optimizer skips non-finite steps
      if (finite) {
        if (count > 1) {
          tail[GNS_BIASED] = 0.0;
        } else {
          tail[GNS_BIASED] = 1.0;
        }
        if (update) { tail[GNS_SQR_UNBIAS] = su; tail[GNS_VAR_UNBIAS] = vu; }
        else if (restart) { tail[GNS_SQR_UNBIAS] = 0.0; tail[GNS_VAR_UNBIAS] = 0.0; }
        progress += gain;                     // progress advances with every finite step
        tail[GNS_PROGRESS] = progress;
      }
      if (a.pair_state) *a.pair_state = (a.pair_mode && finite) ? 1 : 0;
    }
    scale_out = lr_scale;
  }
  __syncthreads();                            // every thread's mailbox payload has been issued
  if (threadIdx.x == 0) {
    const unsigned long long now = globaltimer_ns();
    slot[1] = finite_flag;
    slot[2] = gain;
    slot[3] = progress;
    slot[4] = sh_t[1] + (double)(now > sh_entry ? now - sh_entry : 0ull);
    slot[5] = sh_c[C_ERR];
    slot[6] = scale_out;
    slot[7] = (double)a.n_rows;
    slot[8] = sh_t[0];
    slot[9] = sh_t[2];
    slot[10] = mine[4 * G + 3];
    slot[11] = amp;
    __threadfence_system();                   // payload (all threads') before the sequence number
    *reinterpret_cast<volatile double*>(slot) = (double)(step + 1);   // publish last
    *a.step_ctr = step + 1;                                           // next optimizer step
  }
}

// Tail of every statistics-producing kernel. All threads of all CTAs call it
// after their last global access. With `fuse`, the last CTA of the grid to
// get here runs the step's finalize.
__device__ __forceinline__ void kernel_tail(bool fuse, bool peers, uint32_t* ticket, const FinalizeArgs& f) {
  __shared__ int s_last;
  __syncthreads();                            // every thread of the CTA has issued its stores / atomics
  if (threadIdx.x == 0) {
    // one cumulative fence per CTA (the flag-barrier idiom): the CTA's (remote) stores and
    // statistics atomics are performed before anything this thread does next
    if (peers) __threadfence_system(); else __threadfence();
    int last = 0;
    if (fuse) {
      const unsigned t = atomicAdd(ticket, 1u);
      last = (t == gridDim.x - 1);
      if (last) *ticket = 0u;                 // everybody else has drawn: reset for the next step
    }
    s_last = last;
  }
  if (!fuse) return;
  __syncthreads();
  if (s_last) {
    __threadfence();
    finalize_body(f);
  }
}

// ---------------------------------------------------------------------------
// two-shot P2P flavour
// ---------------------------------------------------------------------------
// W > 0: world size known at compile time (2, 4, 8): the per-peer loads are a
// fully unrolled register array and each thread keeps U = 16/W vectors in
// flight (16 independent 16-byte requests per thread; with 32 CTAs x 512
// threads that is ~4 MB outstanding, enough to cover the ~2-3 us NVLink
// round trip at full link bandwidth). W == 0: generic fallback, runtime world.
template <int W> struct ReduceUnroll { static constexpr int U = 16 / W; };
template <> struct ReduceUnroll<0> { static constexpr int U = 1; };
template <> struct ReduceUnroll<1> { static constexpr int U = 8; };

template <typename T, int W, int PINV>
__global__ void __launch_bounds__(ADL_THREADS, 1)
allreduce_gns_kernel(const ReduceArgs a, const FinalizeArgs f) {
  extern __shared__ double s_stats[];                 // [2][n_groups]
  constexpr int N = VecTraits<T>::N;
  constexpr int U = ReduceUnroll<W>::U;
  constexpr int WMAX = (W > 0) ? W : ADL_MAX_RANKS;
  smem_stats_zero(s_stats, 2 * a.n_groups);
  GroupAccum<2> accum;
  accum.init(s_stats, a.n_groups);
  __shared__ int s_seg_end[ADL_SEG_SMEM], s_seg_group[ADL_SEG_SMEM];
  SegCache segs;
  segs.load(a.segs, s_seg_end, s_seg_group);

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
    Precond<T, PINV> pc[U];
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
        pc[u].issue(a.pinv, a.pinv_wide, base + idx[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      float sq[2] = {0.f, 0.f};
      int g = -1;
      if (active[u]) {
        const int v = base + idx[u];
        g = group_of(segs, v, cur[u]);
        float sum[N], pinv[N];
        if (PINV) pc[u].finish(a.pinv_coef, a.pinv_wide, g, pinv);
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
              const float y = PINV ? x[e] * pinv[e] : x[e];
              sq[0] = fmaf(y, y, sq[0]);
            }
          }
        }
#pragma unroll
        for (int e = 0; e < N; ++e) {
          sum[e] *= a.scale;
          const float y = PINV ? sum[e] * pinv[e] : sum[e];
          sq[1] = fmaf(y, y, sq[1]);
        }
        const Vec16 out = pack<T>(sum);
#pragma unroll
        for (int p = 0; p < WMAX; ++p)
          if (p < world) st_vec(dst[p] + v, out);
        if (!a.want_local) sq[0] = 0.f;
      }
      accum.add(g, sq);
    }
  }
  accum.flush_warp();
  double* outs[2] = {a.want_local ? a.L : nullptr, a.T};
  smem_stats_flush<2>(s_stats, a.n_groups, outs);
  kernel_tail(a.fuse_fin != 0, world > 1, a.ticket, f);
}

// ---------------------------------------------------------------------------
// one-shot push flavour (latency-bound buckets)
// ---------------------------------------------------------------------------
// Every rank copies its bucket into lane `rank` of every peer's staging area,
// fences, raises ONE flag per peer; once all flags are in, the whole
// reduction is local (HBM / L2 reads of the W lanes). Compared with the
// two-shot flavour that is one NVLink round trip instead of two, at (W-1)
// times the bytes -- the host picks it for buckets where latency dominates.
// Statistics: rank r accumulates the vectors of "its" slice only (the same
// partition as the two-shot flavour), so the cross-rank sum of the partials
// is unchanged.
template <typename T, int W, int PINV>
__global__ void __launch_bounds__(ADL_THREADS, 1)
allreduce_oneshot_kernel(const ReduceArgs a, const FinalizeArgs f) {
  extern __shared__ double s_stats[];                 // [2][n_groups]
  constexpr int N = VecTraits<T>::N;
  constexpr int WMAX = (W > 0) ? W : ADL_MAX_RANKS;
  smem_stats_zero(s_stats, 2 * a.n_groups);
  GroupAccum<2> accum;
  accum.init(s_stats, a.n_groups);
  __shared__ int s_seg_end[ADL_SEG_SMEM], s_seg_group[ADL_SEG_SMEM];
  SegCache segs;
  segs.load(a.segs, s_seg_end, s_seg_group);
  const int world = (W > 0) ? W : a.world;
  const int stride = gridDim.x * blockDim.x;
  const int first = blockIdx.x * blockDim.x + threadIdx.x;
  Vec16* mine = static_cast<Vec16*>(a.buf[a.rank]);
  const int iters = (a.n_vec + stride - 1) / stride;

  // push: lane `rank` of every peer's staging area (the own lane is read in place)
  for (int it = 0; it < iters; ++it) {
    const int v = first + it * stride;
    if (v < a.n_vec) {
      const Vec16 x = ld_vec(mine + v);
#pragma unroll
      for (int p = 1; p < WMAX; ++p) {
        if (p < world) {
          const int q = (a.rank + p) % world;
          st_vec(static_cast<Vec16*>(a.stage[q]) + (size_t)a.rank * a.n_vec + v, x);
        }
      }
    }
  }
  cta_barrier_peers(a, 0);             // every peer's push of this CTA's vectors has landed

  const Vec16* lanes = static_cast<const Vec16*>(a.stage[a.rank]);
  const int slice = a.n_vec / world;
  int cur = -1;
  for (int it = 0; it < iters; ++it) {
    const int v = first + it * stride;
    float sq[2] = {0.f, 0.f};
    int g = -1;
    if (v < a.n_vec) {
      Vec16 in[WMAX];                  // indexed by RANK: the sum order is the same on every rank
      Precond<T, PINV> pc;
#pragma unroll
      for (int r = 0; r < WMAX; ++r) {
        if (r < world)
          in[r] = (r == a.rank) ? ld_vec(mine + v) : ld_vec(lanes + (size_t)r * a.n_vec + v);
      }
      pc.issue(a.pinv, a.pinv_wide, v);
      g = group_of(segs, v, cur);
      float sum[N], pinv[N];
      if (PINV) pc.finish(a.pinv_coef, a.pinv_wide, g, pinv);
#pragma unroll
      for (int e = 0; e < N; ++e) sum[e] = 0.f;
#pragma unroll
      for (int r = 0; r < WMAX; ++r) {
        if (r < world) {
          float x[N];
          unpack<T>(in[r], x);
#pragma unroll
          for (int e = 0; e < N; ++e) {
            sum[e] += x[e];
            const float y = PINV ? x[e] * pinv[e] : x[e];
            sq[0] = fmaf(y, y, sq[0]);
          }
        }
      }
#pragma unroll
      for (int e = 0; e < N; ++e) {
        sum[e] *= a.scale;
        const float y = PINV ? sum[e] * pinv[e] : sum[e];
        sq[1] = fmaf(y, y, sq[1]);
      }
      st_vec(mine + v, pack<T>(sum));
      const bool owned = (v / slice) == a.rank;
      if (!owned) { sq[0] = 0.f; sq[1] = 0.f; g = -1; }
      if (!a.want_local) sq[0] = 0.f;
    }
    accum.add(g, sq);
  }
  accum.flush_warp();
  double* outs[2] = {a.want_local ? a.L : nullptr, a.T};
  smem_stats_flush<2>(s_stats, a.n_groups, outs);
  kernel_tail(a.fuse_fin != 0, true, a.ticket, f);
}

// ---------------------------------------------------------------------------
// NVLS flavour: the NVSwitch reduces. `multimem.ld_reduce` on the multicast
// address returns sum_r g_r of a vector in ONE load (the switch pulls every
// GPU's copy and adds in flight), `multimem.st` writes the mean into every
// GPU's arena with ONE store. Per GPU that is ~B(1+1/W) each way instead of
// 2(W-1)/W*B, and W times fewer load instructions.
// The switch hides the per-replica values, so sum_r |g_r|^2 comes from this
// rank's own copy: the own slice is read next to the multimem load in the
// main loop; a foreign slice q is read BEFORE this rank tells rank q that its
// gradients are ready (per-peer start flags, released slice by slice), because
// rank q's multimem.st will overwrite it. CTA c reads, in every slice, exactly
// the vectors CTA c of the owning rank will write.
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

template <typename T, int PINV>
__global__ void __launch_bounds__(ADL_THREADS, 1)
allreduce_nvls_kernel(const ReduceArgs a, const FinalizeArgs f) {
  extern __shared__ double s_stats[];                 // [2][n_groups]
  constexpr int N = VecTraits<T>::N;
  constexpr int U = 8;
  smem_stats_zero(s_stats, 2 * a.n_groups);
  GroupAccum<2> accum;
  accum.init(s_stats, a.n_groups);
  __shared__ int s_seg_end[ADL_SEG_SMEM], s_seg_group[ADL_SEG_SMEM];
  SegCache segs;
  segs.load(a.segs, s_seg_end, s_seg_group);
  const int stride = gridDim.x * blockDim.x;
  const int first = blockIdx.x * blockDim.x + threadIdx.x;
  const Vec16* mine = static_cast<const Vec16*>(a.buf[a.rank]);
  const uint32_t epoch = launch_epoch(a.step_ctr, a.site);

  const int slice = a.n_vec / a.world;
  const int iters = (slice + stride * U - 1) / (stride * U);
  // foreign slices: local statistic first, then this rank's "ready" flag to the owner
  for (int k = 1; k < a.world; ++k) {
    const int q = (a.rank + k) % a.world;
    if (a.want_local) {
      int cur[U];
#pragma unroll
      for (int u = 0; u < U; ++u) cur[u] = -1;
      for (int it = 0; it < iters; ++it) {
        Vec16 in[U];
        Precond<T, PINV> pc[U];
        int idx[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          idx[u] = first + (it * U + u) * stride;
          if (idx[u] < slice) {
            in[u] = ld_vec(mine + q * slice + idx[u]);
            pc[u].issue(a.pinv, a.pinv_wide, q * slice + idx[u]);
          }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          float sq[2] = {0.f, 0.f};
          int g = -1;
          if (idx[u] < slice) {
            const int v = q * slice + idx[u];
            g = group_of(segs, v, cur[u]);
            float x[N], pinv[N];
            unpack<T>(in[u], x);
            if (PINV) pc[u].finish(a.pinv_coef, a.pinv_wide, g, pinv);
#pragma unroll
            for (int e = 0; e < N; ++e) {
              const float y = PINV ? x[e] * pinv[e] : x[e];
              sq[0] = fmaf(y, y, sq[0]);
            }
          }
          accum.add(g, sq);
        }
      }
    }
    __syncthreads();                                  // every lane has consumed its loads of slice q
    if (threadIdx.x == 0) {
      __threadfence_system();
      st_release_sys(pad_slot(a.pad[q], 0, blockIdx.x, a.rank), epoch);
    }
  }
  if (threadIdx.x > 0 && (int)threadIdx.x < a.world) {
    const int peer = (a.rank + threadIdx.x) % a.world;
    wait_flag(pad_slot(a.pad[a.rank], 0, blockIdx.x, peer), epoch, a.timeout_ns, a.err);
  }
  __syncthreads();                                    // every rank's grads are ready

  const int base = a.rank * slice;
  Vec16* mc = static_cast<Vec16*>(a.mc_buf);
  int cur[U];
#pragma unroll
  for (int u = 0; u < U; ++u) cur[u] = -1;
  for (int it = 0; it < iters; ++it) {
    Vec16 in[U], own[U];
    Precond<T, PINV> pc[U];
    int idx[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      idx[u] = first + (it * U + u) * stride;
      if (idx[u] < slice) {
        in[u] = Multimem<T>::ld_reduce(mc + base + idx[u]);
        if (a.want_local) own[u] = ld_vec(mine + base + idx[u]);
        pc[u].issue(a.pinv, a.pinv_wide, base + idx[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      float sq[2] = {0.f, 0.f};
      int g = -1;
      if (idx[u] < slice) {
        const int v = base + idx[u];
        g = group_of(segs, v, cur[u]);
        float x[N], pinv[N];
        if (PINV) pc[u].finish(a.pinv_coef, a.pinv_wide, g, pinv);
        if (a.want_local) {
          unpack<T>(own[u], x);
#pragma unroll
          for (int e = 0; e < N; ++e) {
            const float y = PINV ? x[e] * pinv[e] : x[e];
            sq[0] = fmaf(y, y, sq[0]);
          }
        }
        unpack<T>(in[u], x);
#pragma unroll
        for (int e = 0; e < N; ++e) {
          x[e] *= a.scale;
          const float y = PINV ? x[e] * pinv[e] : x[e];
          sq[1] = fmaf(y, y, sq[1]);
        }
        multimem_st(mc + v, pack<T>(x));
      }
      accum.add(g, sq);
    }
  }
  accum.flush_warp();
  double* outs[2] = {a.want_local ? a.L : nullptr, a.T};
  smem_stats_flush<2>(s_stats, a.n_groups, outs);
  kernel_tail(a.fuse_fin != 0, true, a.ticket, f);
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
  int pinv_mode;
  int pinv_wide;
  const float* pinv_coef;
  int fuse_fin;
  uint32_t* ticket;
};

// MODE 0: fold_acc   (a += g ; s0 += |g|^2 ; g = 0)
// MODE 1: fold_final (s0 += |g|^2 ; g += a ; a = 0)
// MODE 2: pair       (s0 += |g|^2 ; if flag: s1 += |pv|^2, s2 += |(g+pv)/2|^2 ; pv = g)
template <typename T, int MODE, int PINV>
__global__ void __launch_bounds__(ADL_THREADS, 2)
local_kernel(const LocalArgs a, const FinalizeArgs f) {
  extern __shared__ double s_stats[];                 // [3][n_groups]
  constexpr int N = VecTraits<T>::N;
  constexpr int K = 3;
  constexpr int U = 4;                                // vectors in flight per thread and tensor
  smem_stats_zero(s_stats, K * a.n_groups);
  GroupAccum<K> accum;
  accum.init(s_stats, a.n_groups);
  __shared__ int s_seg_end[ADL_SEG_SMEM], s_seg_group[ADL_SEG_SMEM];
  SegCache segs;
  segs.load(a.segs, s_seg_end, s_seg_group);
  const int stride = gridDim.x * blockDim.x;
  const int first = blockIdx.x * blockDim.x + threadIdx.x;
  const int iters = (a.n_vec + stride * U - 1) / (stride * U);
  const bool have_prev = a.flag_ptr ? (*a.flag_ptr != 0) : (a.flag != 0);
  int cur[U];
#pragma unroll
  for (int u = 0; u < U; ++u) cur[u] = -1;
  for (int it = 0; it < iters; ++it) {
    Vec16 gv[U], ov[U];
    Precond<T, PINV> pc[U];
    int idx[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      idx[u] = first + (it * U + u) * stride;
      if (idx[u] < a.n_vec) {
        gv[u] = ld_vec(static_cast<const Vec16*>(a.g) + idx[u]);
        if (MODE == 2) ov[u] = ld_vec(static_cast<const Vec16*>(a.pv) + idx[u]);
        else ov[u] = ld_vec(static_cast<const Vec16*>(a.a) + idx[u]);
        pc[u].issue(a.pinv, a.pinv_wide, idx[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int v = idx[u];
      float sq[K] = {0.f, 0.f, 0.f};
      int grp = -1;
      if (v < a.n_vec) {
        float g[N], o[N], pinv[N];
        grp = group_of(segs, v, cur[u]);
        if (PINV) pc[u].finish(a.pinv_coef, a.pinv_wide, grp, pinv);
        unpack<T>(gv[u], g);
        unpack<T>(ov[u], o);
#pragma unroll
        for (int e = 0; e < N; ++e) {
          const float y = PINV ? g[e] * pinv[e] : g[e];
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
              const float p = PINV ? o[e] * pinv[e] : o[e];
              const float m = 0.5f * ((PINV ? g[e] * pinv[e] : g[e]) + p);
              sq[1] = fmaf(p, p, sq[1]);
              sq[2] = fmaf(m, m, sq[2]);
            }
          }
          st_vec(static_cast<Vec16*>(a.pv) + v, gv[u]);
        }
      }
      accum.add(grp, sq);
    }
  }
  accum.flush_warp();
  double* outs[K] = {a.s0, a.s1, a.s2};
  smem_stats_flush<K>(s_stats, a.n_groups, outs);
  if (MODE == 2 && a.fuse_fin) kernel_tail(true, false, a.ticket, f);
}

__global__ void __launch_bounds__(256, 1) finalize_stats_kernel(const FinalizeArgs a) {
  finalize_body(a);
}

__global__ void stamp_kernel(unsigned long long* dst) { *dst = globaltimer_ns(); }

// accumulation micro-step: add the interval since the previous mark to the step clock
__global__ void step_mark_kernel(unsigned long long* last_stamp, double* clock) {
  const unsigned long long now = globaltimer_ns();
  const unsigned long long last = *last_stamp;
  if (last != 0ull && now > last) { clock[0] += (double)(now - last); clock[1] += 1.0; }
  *last_stamp = now;
}

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
#define ADL_SMEM_ATTR(...) \
  ADL_CHECK(cudaFuncSetAttribute(__VA_ARGS__, cudaFuncAttributeMaxDynamicSharedMemorySize, ADL_MAX_STAT_SMEM))
template <typename T, int PINV>
static int set_attrs_for_p() {
  ADL_SMEM_ATTR(allreduce_gns_kernel<T, 0, PINV>);
  ADL_SMEM_ATTR(allreduce_gns_kernel<T, 1, PINV>);
  ADL_SMEM_ATTR(allreduce_gns_kernel<T, 2, PINV>);
  ADL_SMEM_ATTR(allreduce_gns_kernel<T, 4, PINV>);
  ADL_SMEM_ATTR(allreduce_gns_kernel<T, 8, PINV>);
  ADL_SMEM_ATTR(allreduce_oneshot_kernel<T, 0, PINV>);
  ADL_SMEM_ATTR(allreduce_oneshot_kernel<T, 2, PINV>);
  ADL_SMEM_ATTR(allreduce_oneshot_kernel<T, 4, PINV>);
  ADL_SMEM_ATTR(allreduce_oneshot_kernel<T, 8, PINV>);
  ADL_SMEM_ATTR(allreduce_nvls_kernel<T, PINV>);
  ADL_SMEM_ATTR(local_kernel<T, 0, PINV>);
  ADL_SMEM_ATTR(local_kernel<T, 1, PINV>);
  ADL_SMEM_ATTR(local_kernel<T, 2, PINV>);
  return 0;
}
template <typename T>
static int set_attrs_for() {
  if (int rc = set_attrs_for_p<T, 0>()) return rc;
  if (int rc = set_attrs_for_p<T, 1>()) return rc;
  if (int rc = set_attrs_for_p<T, 2>()) return rc;
  return 0;
}

// dispatch helper: dtype code x preconditioner mode -> template instance
#define ADL_DISPATCH_TP(dtype, pinv, CALL)                                         \
  do {                                                                             \
    if ((pinv) < 0 || (pinv) > 2) return -5;                                       \
    switch ((dtype) * 3 + (pinv)) {                                                \
      case 0: CALL(float, 0); break;                                               \
      case 1: CALL(float, 1); break;                                               \
      case 2: CALL(float, 2); break;                                               \
      case 3: CALL(__nv_bfloat16, 0); break;                                       \
      case 4: CALL(__nv_bfloat16, 1); break;                                       \
      case 5: CALL(__nv_bfloat16, 2); break;                                       \
      case 6: CALL(__half, 0); break;                                              \
      case 7: CALL(__half, 1); break;                                              \
      case 8: CALL(__half, 2); break;                                              \
      default: return -2;                                                          \
    }                                                                              \
  } while (0)

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
// flavour: 0 = two-shot P2P, 1 = one-shot push (args->stage), 2 = NVLS (args->mc_buf)
// fin: finalize arguments when args->fuse_fin (else ignored; may be nullptr)
// threads: CTA size (a power of two, 64..ADL_THREADS); small CTAs leave registers and warp slots
//          of their SM to the backward kernels running next to them
int adl_allreduce_gns(const ReduceArgs* args, const FinalizeArgs* fin, int dtype, int flavour,
                      int grid, int threads, void* stream) {
  if (g_device >= 0) ADL_CHECK(cudaSetDevice(g_device));
  if (threads < 64 || threads > ADL_THREADS || (threads & (threads - 1))) return -8;
  const size_t smem = sizeof(double) * 2 * args->n_groups;
  if (smem > ADL_MAX_STAT_SMEM) return -4;
  if (dtype < 0 || dtype > 2) return -2;
  if (args->fuse_fin && fin == nullptr) return -6;
  FinalizeArgs none;
  memset(&none, 0, sizeof(none));
  const FinalizeArgs& f = args->fuse_fin ? *fin : none;
  const int pinv = args->pinv != nullptr ? args->pinv_mode : 0;
  cudaStream_t s = (cudaStream_t)stream;
  if (args->world > 1 && flavour == 2) {
    if (args->mc_buf == nullptr) return -7;
#define CALL_NVLS(T, P) allreduce_nvls_kernel<T, P><<<grid, threads, smem, s>>>(*args, f)
    ADL_DISPATCH_TP(dtype, pinv, CALL_NVLS);
#undef CALL_NVLS
    return (int)cudaGetLastError();
  }
  if (args->world > 1 && flavour == 1) {
#define CALL_OS(T, P)                                                                                  \
  do {                                                                                                 \
    switch (args->world) {                                                                             \
      case 2: allreduce_oneshot_kernel<T, 2, P><<<grid, threads, smem, s>>>(*args, f); break;      \
      case 4: allreduce_oneshot_kernel<T, 4, P><<<grid, threads, smem, s>>>(*args, f); break;      \
      case 8: allreduce_oneshot_kernel<T, 8, P><<<grid, threads, smem, s>>>(*args, f); break;      \
      default: allreduce_oneshot_kernel<T, 0, P><<<grid, threads, smem, s>>>(*args, f); break;     \
    }                                                                                                  \
  } while (0)
    ADL_DISPATCH_TP(dtype, pinv, CALL_OS);
#undef CALL_OS
    return (int)cudaGetLastError();
  }
#define CALL_AR(T, P)                                                                              \
  do {                                                                                             \
    switch (args->world) {                                                                         \
      case 1: allreduce_gns_kernel<T, 1, P><<<grid, threads, smem, s>>>(*args, f); break;      \
      case 2: allreduce_gns_kernel<T, 2, P><<<grid, threads, smem, s>>>(*args, f); break;      \
      case 4: allreduce_gns_kernel<T, 4, P><<<grid, threads, smem, s>>>(*args, f); break;      \
      case 8: allreduce_gns_kernel<T, 8, P><<<grid, threads, smem, s>>>(*args, f); break;      \
      default: allreduce_gns_kernel<T, 0, P><<<grid, threads, smem, s>>>(*args, f); break;     \
    }                                                                                              \
  } while (0)
  ADL_DISPATCH_TP(dtype, pinv, CALL_AR);
#undef CALL_AR
  return (int)cudaGetLastError();
}

// mode: 0 fold_acc, 1 fold_final, 2 pair
int adl_local(const LocalArgs* args, const FinalizeArgs* fin, int mode, int dtype, int grid,
              int threads, void* stream) {
  if (g_device >= 0) ADL_CHECK(cudaSetDevice(g_device));
  if (threads < 64 || threads > ADL_THREADS || (threads & (threads - 1))) return -8;
  const size_t smem = sizeof(double) * 3 * args->n_groups;
  if (smem > ADL_MAX_STAT_SMEM) return -4;
  if (dtype < 0 || dtype > 2) return -2;
  if (args->fuse_fin && (fin == nullptr || mode != 2)) return -6;
  FinalizeArgs none;
  memset(&none, 0, sizeof(none));
  const FinalizeArgs& f = args->fuse_fin ? *fin : none;
  const int pinv = args->pinv != nullptr ? args->pinv_mode : 0;
  cudaStream_t s = (cudaStream_t)stream;
#define CALL_L(T, P)                                                                         \
  do {                                                                                       \
    if (mode == 0) local_kernel<T, 0, P><<<grid, threads, smem, s>>>(*args, f);          \
    else if (mode == 1) local_kernel<T, 1, P><<<grid, threads, smem, s>>>(*args, f);     \
    else if (mode == 2) local_kernel<T, 2, P><<<grid, threads, smem, s>>>(*args, f);     \
    else return -3;                                                                          \
  } while (0)
  ADL_DISPATCH_TP(dtype, pinv, CALL_L);
#undef CALL_L
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

int adl_step_mark(unsigned long long* last_stamp, double* clock, void* stream) {
  if (g_device >= 0) ADL_CHECK(cudaSetDevice(g_device));
  step_mark_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(last_stamp, clock);
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
int adl_mbox_hdr() { return ADL_MBOX_HDR; }
int adl_xchg_tail() { return ADL_XCHG_TAIL; }

}  // extern "C"
