// adaptdl_b200 -- fused BatchNorm (+ residual add) (+ ReLU) for channels-last
// activations, training mode, forward and backward (sm_100a).
//
// The reference's headline workload (examples/pytorch-cifar ResNet-18) spends
// two thirds of its GPU time in BatchNorm / ReLU / residual-add kernels, each a
// separate full pass over the activation in stock PyTorch:
//   forward : bn (read x twice, write y) + add (read 2, write 1) + relu (read 1, write 1)
//   backward: relu_bwd (read 2, write 1) + bn_bwd (read dy/x several times, write dx)
// Here an activation is [M = N*H*W, C] (channels-last); every kernel moves 16-byte
// vectors of 8 bf16 / 4 fp32 channels:
//   forward : stats (read x) -> finalize (C values) -> apply (re-read x out of the
//             126 MB L2, read residual, write y)
//   backward: bstats (read dy, y, x) -> bfinalize -> bapply (re-read from L2, write
//             dx and the residual gradient)
// Reductions are two-level and deterministic (per-CTA partials in a fixed order, no
// float atomics).
#include "adl_common.cuh"

namespace {

constexpr int BN_THREADS = 256;
// row vectors in flight per thread and tensor: the backward kernels stream three tensors
template <bool BWD> struct BnUnroll { static constexpr int U = BWD ? 2 : 4; };

struct BnArgs {
  const void* x;        // [M, C]
  const void* res;      // [M, C] residual added before the activation (or null)
  void* y;              // [M, C]
  const void* dy;       // backward: [M, C]
  void* dx;             // backward: [M, C]
  void* dres;           // backward: [M, C] gradient of the residual (or null)
  const float* gamma;   // [C]
  const float* beta;    // [C]
  float* mean;          // [C]  saved batch mean
  float* rstd;          // [C]  saved 1/sqrt(var + eps)
  float* running_mean;  // [C] (or null)
  float* running_var;   // [C] (or null)
  long long* num_batches_tracked;  // scalar counter of nn.BatchNorm (or null): +1 per forward
  float* dgamma;        // [C]
  float* dbeta;         // [C]
  float* partial;       // [C / cb, grid.y, 2, cb] scratch
  int* counters;        // [C / cb] zero-initialised tickets (left zero again)
  uint8_t* mask;        // unused (kept for ABI stability; a 1-bit ReLU mask replacing the y read of
                        // the backward pass was measured and was SLOWER: profiles/r2_validate)
  int cb;               // channels per reduce CTA (<= 64)
  float* coef;          // [2, C] scratch: forward (scale, shift); backward (s1/M, s2/M)
  int M, C;
  int n_partial;
  int relu;
  float eps, momentum;
};

// Per-channel sums over the rows of two quantities (forward: x, x^2; backward: g, g*xhat),
// then the per-channel terms, in ONE launch:
//   grid.x = channel blocks of `cb` channels, grid.y = row chunks. Every CTA writes its
//   partial sums; the LAST CTA to finish in a channel block (ticket counter, self-resetting)
//   folds that block's grid.y partials in a fixed order (deterministic) and derives
//   mean / rstd / scale / shift (forward) or dgamma / dbeta / s1/M / s2/M (backward).
template <typename T, bool BWD>
__global__ void __launch_bounds__(BN_THREADS)
bn_reduce_kernel(const BnArgs a) {
  constexpr int V = VecTraits<T>::N;
  constexpr int BN_UNROLL = BnUnroll<BWD>::U;
  __shared__ float red[2 * 32 * 64];             // [2][rows_per_iter][cb]
  __shared__ int is_last;
  // programmatic dependent launch: the apply kernel of this layer may be scheduled while
  // this grid is still running (it parks in griddepcontrol.wait until we are done)
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const int cb = a.cb;
  const int tpr = cb / V;                        // threads per row
  const int rpi = BN_THREADS / tpr;              // rows per iteration
  const int c0 = blockIdx.x * cb;
  const int lc = (threadIdx.x % tpr) * V;        // channel within the block
  const int my_c = c0 + lc;
  const int my_r = threadIdx.x / tpr;
  float s0[V], s1[V];
#pragma unroll
  for (int e = 0; e < V; ++e) { s0[e] = 0.f; s1[e] = 0.f; }
  float mu[V], rs[V];
  if (BWD) {
#pragma unroll
    for (int e = 0; e < V; ++e) { mu[e] = a.mean[my_c + e]; rs[e] = a.rstd[my_c + e]; }
  }
  const long long stride = (long long)gridDim.y * rpi;
  for (long long r0 = (long long)blockIdx.y * rpi + my_r; r0 < a.M; r0 += stride * BN_UNROLL) {
    Vec16 vx[BN_UNROLL], vg[BN_UNROLL], vy[BN_UNROLL];
#pragma unroll
    for (int u = 0; u < BN_UNROLL; ++u) {
      const long long r = r0 + u * stride;
      if (r < a.M) {
        const size_t off = ((size_t)r * a.C + my_c) * sizeof(T);
        vx[u] = ld_vec(static_cast<const char*>(a.x) + off);
        if (BWD) {
          vg[u] = ld_vec(static_cast<const char*>(a.dy) + off);
          if (a.relu) vy[u] = ld_vec(static_cast<const char*>(a.y) + off);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < BN_UNROLL; ++u) {
      const long long r = r0 + u * stride;
      if (r < a.M) {
        float fx[V];
        unpack<T>(vx[u], fx);
        if (!BWD) {
#pragma unroll
          for (int e = 0; e < V; ++e) { s0[e] += fx[e]; s1[e] = fmaf(fx[e], fx[e], s1[e]); }
        } else {
          float fg[V], fy[V];
          unpack<T>(vg[u], fg);
          if (a.relu) unpack<T>(vy[u], fy);
#pragma unroll
          for (int e = 0; e < V; ++e) {
            const float g = (a.relu && !(fy[e] > 0.f)) ? 0.f : fg[e];
            s0[e] += g;
            s1[e] = fmaf(g, (fx[e] - mu[e]) * rs[e], s1[e]);
          }
        }
      }
    }
  }
  float* r0s = red;
  float* r1s = red + rpi * cb;
#pragma unroll
  for (int e = 0; e < V; ++e) {
    r0s[my_r * cb + lc + e] = s0[e];
    r1s[my_r * cb + lc + e] = s1[e];
  }
  __syncthreads();
  float* mine = a.partial + ((size_t)blockIdx.x * gridDim.y + blockIdx.y) * 2 * cb;
  for (int c = threadIdx.x; c < 2 * cb; c += BN_THREADS) {
    const float* src = (c < cb) ? (r0s + c) : (r1s + (c - cb));
    float t = 0.f;
    for (int r = 0; r < rpi; ++r) t += src[r * cb];
    mine[c] = t;
  }
  // ---- last CTA of this channel block finalizes ----
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) is_last = (atomicAdd(a.counters + blockIdx.x, 1) == (int)gridDim.y - 1);
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  // fold the grid.y partials of this channel block: warp w takes partials w, w+8, ...,
  // lane l the l-th float4 of the 2*cb values; then the 8 warp sums are added in order
  const int nv4 = 2 * cb / 4;                    // float4s per partial (<= 32)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (lane < nv4) {
    const float4* base = reinterpret_cast<const float4*>(a.partial + (size_t)blockIdx.x * gridDim.y * 2 * cb) + lane;
    const int gy = (int)gridDim.y;
    int p = warp;
    for (; p + 24 < gy; p += 32) {
      const float4 t0 = __ldcg(base + (size_t)p * nv4), t1 = __ldcg(base + (size_t)(p + 8) * nv4);
      const float4 t2 = __ldcg(base + (size_t)(p + 16) * nv4), t3 = __ldcg(base + (size_t)(p + 24) * nv4);
      acc.x += (t0.x + t1.x) + (t2.x + t3.x); acc.y += (t0.y + t1.y) + (t2.y + t3.y);
      acc.z += (t0.z + t1.z) + (t2.z + t3.z); acc.w += (t0.w + t1.w) + (t2.w + t3.w);
    }
    for (; p < gy; p += 8) {
      const float4 t = __ldcg(base + (size_t)p * nv4);
      acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
    }
  }
  __syncthreads();                               // everyone is done with `red`
  if (lane < nv4) reinterpret_cast<float4*>(red)[warp * nv4 + lane] = acc;
  __syncthreads();
  float* tot = red + 8 * 2 * cb;                 // [2][cb]
  if (threadIdx.x < 2 * cb) {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < BN_THREADS / 32; ++w) t += (double)red[w * 2 * cb + threadIdx.x];
    tot[threadIdx.x] = (float)(t / (double)a.M);  // mean-like quantities from here on
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    a.counters[blockIdx.x] = 0;
    if (!BWD && blockIdx.x == 0 && a.num_batches_tracked != nullptr) *a.num_batches_tracked += 1;
  }
  if (threadIdx.x >= cb) return;
  const int c = c0 + threadIdx.x;
  const float m0 = tot[threadIdx.x], m1 = tot[cb + threadIdx.x];
  if (!BWD) {
    const float var = fmaxf(m1 - m0 * m0, 0.f);
    const float rstd = rsqrtf(var + a.eps);
    a.mean[c] = m0;
    a.rstd[c] = rstd;
    const float scale = a.gamma[c] * rstd;
    a.coef[c] = scale;
    a.coef[a.C + c] = a.beta[c] - m0 * scale;
    if (a.running_mean != nullptr) {
      const float unbiased = a.M > 1 ? var * ((float)a.M / (float)(a.M - 1)) : var;
      a.running_mean[c] = (1.f - a.momentum) * a.running_mean[c] + a.momentum * m0;
      a.running_var[c] = (1.f - a.momentum) * a.running_var[c] + a.momentum * unbiased;
    }
  } else {
    a.dbeta[c] = m0 * (float)a.M;
    a.dgamma[c] = m1 * (float)a.M;
    a.coef[c] = m0;
    a.coef[a.C + c] = m1;
  }
}

// forward:  y  = act(x * scale + shift + res)
// backward: dx = gamma * rstd * (g - s1/M - xhat * s2/M),  dres = g,  g = dy * [y > 0]
template <typename T, bool BWD>
__global__ void __launch_bounds__(BN_THREADS)
bn_apply_kernel(const BnArgs a) {
  constexpr int V = VecTraits<T>::N;
  constexpr int BN_UNROLL = BnUnroll<BWD>::U;
  const int tpr = a.C / V;
  const int rpi = BN_THREADS / tpr;
  const int my_c = (threadIdx.x % tpr) * V;
  const int my_r = threadIdx.x / tpr;
  // launched with programmatic stream serialization right behind bn_reduce: everything
  // above overlapped its tail; its coefficients are visible after this wait
  asm volatile("griddepcontrol.wait;" ::: "memory");
  float k0[V], k1[V], mu[V], rs[V];
#pragma unroll
  for (int e = 0; e < V; ++e) {
    k0[e] = a.coef[my_c + e];
    k1[e] = a.coef[a.C + my_c + e];
    if (BWD) {
      mu[e] = a.mean[my_c + e];
      rs[e] = a.rstd[my_c + e];
    }
  }
  float gs[V];
  if (BWD) {
#pragma unroll
    for (int e = 0; e < V; ++e) gs[e] = a.gamma[my_c + e] * rs[e];
  }
  const long long stride = (long long)gridDim.x * rpi;
  for (long long r0 = (long long)blockIdx.x * rpi + my_r; r0 < a.M; r0 += stride * BN_UNROLL) {
    Vec16 vx[BN_UNROLL], vb[BN_UNROLL], vy[BN_UNROLL];
#pragma unroll
    for (int u = 0; u < BN_UNROLL; ++u) {
      const long long r = r0 + u * stride;
      if (r < a.M) {
        const size_t off = ((size_t)r * a.C + my_c) * sizeof(T);
        vx[u] = ld_vec(static_cast<const char*>(a.x) + off);
        if (!BWD) {
          if (a.res) vb[u] = ld_vec(static_cast<const char*>(a.res) + off);
        } else {
          vb[u] = ld_vec(static_cast<const char*>(a.dy) + off);
          if (a.relu) vy[u] = ld_vec(static_cast<const char*>(a.y) + off);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < BN_UNROLL; ++u) {
      const long long r = r0 + u * stride;
      if (r < a.M) {
        const size_t off = ((size_t)r * a.C + my_c) * sizeof(T);
        float fx[V], fb[V], out[V];
        unpack<T>(vx[u], fx);
        if (!BWD) {
          if (a.res) unpack<T>(vb[u], fb);
#pragma unroll
          for (int e = 0; e < V; ++e) {
            float v = fmaf(fx[e], k0[e], k1[e]);
            if (a.res) v += fb[e];
            out[e] = a.relu ? fmaxf(v, 0.f) : v;
          }
          st_vec(static_cast<char*>(a.y) + off, pack<T>(out));
        } else {
          float fy[V], g[V];
          unpack<T>(vb[u], fb);
          if (a.relu) unpack<T>(vy[u], fy);
#pragma unroll
          for (int e = 0; e < V; ++e) {
            g[e] = (a.relu && !(fy[e] > 0.f)) ? 0.f : fb[e];
            const float xhat = (fx[e] - mu[e]) * rs[e];
            out[e] = gs[e] * (g[e] - k0[e] - xhat * k1[e]);
          }
          st_vec(static_cast<char*>(a.dx) + off, pack<T>(out));
          if (a.dres) st_vec(static_cast<char*>(a.dres) + off, pack<T>(g));
        }
      }
    }
  }
}

// The apply kernel is a programmatic dependent of the reduce kernel (PDL): its launch latency
// (2-3 us, 40 kernel pairs per ResNet-18 step) hides behind the reduction; under stream capture
// the attribute becomes a programmatic edge of the CUDA graph.
static int g_bn_pdl = 0;

template <typename T, bool BWD>
int launch_pair(const BnArgs& a, dim3 grid, int grid_apply, cudaStream_t s) {
  bn_reduce_kernel<T, BWD><<<grid, BN_THREADS, 0, s>>>(a);
  if (!g_bn_pdl) {
    bn_apply_kernel<T, BWD><<<grid_apply, BN_THREADS, 0, s>>>(a);
    return (int)cudaGetLastError();
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid_apply);
  cfg.blockDim = dim3(BN_THREADS);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, bn_apply_kernel<T, BWD>, a);
  if (e != cudaSuccess) return (int)e;
  return (int)cudaGetLastError();
}

template <typename T>
int run(const BnArgs& a, int backward, int grid_y, int grid_apply, cudaStream_t s) {
  const dim3 grid(a.C / a.cb, grid_y);
  return backward ? launch_pair<T, true>(a, grid, grid_apply, s)
                  : launch_pair<T, false>(a, grid, grid_apply, s);
}


// ---------------------------------------------------------------------------
// Single-launch variant: reduce -> grid barrier -> apply in ONE cooperative kernel.
//
// The two-kernel path above is latency-bound on ResNet-sized activations (2-16 MB, L2 resident):
// launch (2-3 us) -> loads (1 us) -> partial store + ticket -> the LAST CTA folds up to 592
// partial rows alone (2-3 us) -> second launch -> apply. Here every CTA owns a set of rows with
// ALL channels; after the grid barrier every CTA folds the (few) partial rows itself, in a fixed
// order (deterministic, identical in all CTAs), derives the coefficients into shared memory and
// streams its rows again (out of L2) for the elementwise pass. One launch, no serial fold, no
// coefficient round trip through global memory.
//
// Launched with cudaLaunchCooperativeKernel (co-residency of the grid is guaranteed by the
// runtime, also under stream capture); the barrier is two self-resetting counters.
// ---------------------------------------------------------------------------
constexpr int BN_FUSED_MAX_C = 1024;
constexpr int BN_FUSED_THREADS = 512;               // one CTA per SM: 512 threads of loads in flight

__device__ __forceinline__ void bn_grid_barrier(int* ctr, int n_cta) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(ctr, 1);
    while (*reinterpret_cast<volatile int*>(ctr) < n_cta) { }
    __threadfence();
    if (atomicAdd(ctr + 1, 1) == n_cta - 1) {       // last one out resets both counters
      ctr[0] = 0;
      ctr[1] = 0;
    }
  }
  __syncthreads();
}

template <typename T, bool BWD>
__global__ void __launch_bounds__(BN_FUSED_THREADS, 1)
bn_fused_kernel(const BnArgs a) {
  constexpr int V = VecTraits<T>::N;
  constexpr int U = BnUnroll<BWD>::U;
  constexpr int NT = BN_FUSED_THREADS;
  __shared__ float red[2 * NT * 8];               // [2][rows_per_iter][C] (rpi * C == NT * V); fold scratch
  __shared__ float ks[2 * BN_FUSED_MAX_C];        // forward (scale, shift); backward (s1/M, s2/M)
  const int C = a.C;
  const int tpr = C / V;                          // threads per row
  const int rpi = NT / tpr;                       // rows per iteration
  const int my_c = (threadIdx.x % tpr) * V;
  const int my_r = threadIdx.x / tpr;
  const int G = gridDim.x;

  float mu[V], rs[V];
  if (BWD) {
#pragma unroll
    for (int e = 0; e < V; ++e) { mu[e] = a.mean[my_c + e]; rs[e] = a.rstd[my_c + e]; }
  }
  // ---- phase 1: this CTA's partial sums (forward: x, x^2; backward: g, g * xhat) ----
  float s0[V], s1[V];
#pragma unroll
  for (int e = 0; e < V; ++e) { s0[e] = 0.f; s1[e] = 0.f; }
  const long long stride = (long long)G * rpi;
  for (long long r0 = (long long)blockIdx.x * rpi + my_r; r0 < a.M; r0 += stride * U) {
    Vec16 vx[U], vg[U], vy[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long r = r0 + u * stride;
      if (r < a.M) {
        const size_t off = ((size_t)r * C + my_c) * sizeof(T);
        vx[u] = ld_vec(static_cast<const char*>(a.x) + off);
        if (BWD) {
          vg[u] = ld_vec(static_cast<const char*>(a.dy) + off);
          if (a.relu) vy[u] = ld_vec(static_cast<const char*>(a.y) + off);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long r = r0 + u * stride;
      if (r < a.M) {
        float fx[V];
        unpack<T>(vx[u], fx);
        if (!BWD) {
#pragma unroll
          for (int e = 0; e < V; ++e) { s0[e] += fx[e]; s1[e] = fmaf(fx[e], fx[e], s1[e]); }
        } else {
          float fg[V], fy[V];
          unpack<T>(vg[u], fg);
          if (a.relu) unpack<T>(vy[u], fy);
#pragma unroll
          for (int e = 0; e < V; ++e) {
            const float g = (a.relu && !(fy[e] > 0.f)) ? 0.f : fg[e];
            s0[e] += g;
            s1[e] = fmaf(g, (fx[e] - mu[e]) * rs[e], s1[e]);
          }
        }
      }
    }
  }
  float* r0s = red;
  float* r1s = red + rpi * C;
#pragma unroll
  for (int e = 0; e < V; ++e) {
    r0s[my_r * C + my_c + e] = s0[e];
    r1s[my_r * C + my_c + e] = s1[e];
  }
  __syncthreads();
  float* mine = a.partial + (size_t)blockIdx.x * 2 * C;
  for (int c = threadIdx.x; c < 2 * C; c += NT) {
    const float* src = (c < C) ? (r0s + c) : (r1s + (c - C));
    float t = 0.f;
    for (int r = 0; r < rpi; ++r) t += src[r * C];
    mine[c] = t;
  }
  bn_grid_barrier(a.counters, G);

  // ---- fold all CTAs' partials: identical (fixed order) in every CTA, deterministic.
  // The [G][2C] matrix is read as float4 columns by all 512 threads, eight rows in flight per
  // thread (one L2 round trip for G * C <= 8192), then combined through shared memory.
  {
    const int q = 2 * C / 4;                      // float4 columns
    const int L = NT / q > 0 ? NT / q : 1;        // row lanes
    float* fold = red;                            // [L][2C]
    for (int qi0 = 0; qi0 < q; qi0 += NT) {       // q > NT only for C > 1024 (never: C <= 1024)
      const int qi = qi0 + (int)(threadIdx.x % q);
      const int li = threadIdx.x / q;
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      if (li < L && qi < q) {
        const float4* base = reinterpret_cast<const float4*>(a.partial) + qi;
        for (int g0 = li; g0 < G; g0 += 8 * L) {
          float4 t[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int g = g0 + j * L;
            t[j] = (g < G) ? __ldcg(base + (size_t)g * q) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) { acc.x += t[j].x; acc.y += t[j].y; acc.z += t[j].z; acc.w += t[j].w; }
        }
        reinterpret_cast<float4*>(fold)[li * q + qi] = acc;
      }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < 2 * C; c += NT) {
      double t = 0.0;
      for (int l = 0; l < L; ++l) t += (double)fold[l * 2 * C + c];
      ks[c] = (float)(t / (double)a.M);
    }
    __syncthreads();
  }
  for (int c = threadIdx.x; c < C; c += NT) {
    const float m0 = ks[c], m1 = ks[C + c];
    if (!BWD) {
      const float var = fmaxf(m1 - m0 * m0, 0.f);
      const float rstd = rsqrtf(var + a.eps);
      const float scale = a.gamma[c] * rstd;
      if (blockIdx.x == 0) {
        a.mean[c] = m0;
        a.rstd[c] = rstd;
        if (a.running_mean != nullptr) {
          const float unbiased = a.M > 1 ? var * ((float)a.M / (float)(a.M - 1)) : var;
          a.running_mean[c] = (1.f - a.momentum) * a.running_mean[c] + a.momentum * m0;
          a.running_var[c] = (1.f - a.momentum) * a.running_var[c] + a.momentum * unbiased;
        }
      }
      // (scale, shift) replace (mean, mean-of-squares): same thread, same slots
      ks[c] = scale;
      ks[C + c] = a.beta[c] - m0 * scale;
    } else if (blockIdx.x == 0) {
      a.dbeta[c] = m0 * (float)a.M;
      a.dgamma[c] = m1 * (float)a.M;
    }
  }
  if (!BWD && blockIdx.x == 0 && threadIdx.x == 0 && a.num_batches_tracked != nullptr)
    *a.num_batches_tracked += 1;
  __syncthreads();

  // ---- phase 2: elementwise pass over the same rows (second read comes out of L2) ----
  float k0[V], k1[V], gs[V];
#pragma unroll
  for (int e = 0; e < V; ++e) {
    k0[e] = ks[my_c + e];
    k1[e] = ks[C + my_c + e];
    if (BWD) gs[e] = a.gamma[my_c + e] * rs[e];
  }
  for (long long r0 = (long long)blockIdx.x * rpi + my_r; r0 < a.M; r0 += stride * U) {
    Vec16 vx[U], vb[U], vy[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long r = r0 + u * stride;
      if (r < a.M) {
        const size_t off = ((size_t)r * C + my_c) * sizeof(T);
        vx[u] = ld_vec(static_cast<const char*>(a.x) + off);
        if (!BWD) {
          if (a.res) vb[u] = ld_vec(static_cast<const char*>(a.res) + off);
        } else {
          vb[u] = ld_vec(static_cast<const char*>(a.dy) + off);
          if (a.relu) vy[u] = ld_vec(static_cast<const char*>(a.y) + off);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long r = r0 + u * stride;
      if (r < a.M) {
        const size_t off = ((size_t)r * C + my_c) * sizeof(T);
        float fx[V], fb[V], out[V];
        unpack<T>(vx[u], fx);
        if (!BWD) {
          if (a.res) unpack<T>(vb[u], fb);
#pragma unroll
          for (int e = 0; e < V; ++e) {
            float v = fmaf(fx[e], k0[e], k1[e]);
            if (a.res) v += fb[e];
            out[e] = a.relu ? fmaxf(v, 0.f) : v;
          }
          st_vec(static_cast<char*>(a.y) + off, pack<T>(out));
        } else {
          float fy[V], g[V];
          unpack<T>(vb[u], fb);
          if (a.relu) unpack<T>(vy[u], fy);
#pragma unroll
          for (int e = 0; e < V; ++e) {
            g[e] = (a.relu && !(fy[e] > 0.f)) ? 0.f : fb[e];
            const float xhat = (fx[e] - mu[e]) * rs[e];
            out[e] = gs[e] * (g[e] - k0[e] - xhat * k1[e]);
          }
          st_vec(static_cast<char*>(a.dx) + off, pack<T>(out));
          if (a.dres) st_vec(static_cast<char*>(a.dres) + off, pack<T>(g));
        }
      }
    }
  }
}

template <typename T, bool BWD>
int launch_fused(const BnArgs& a, int grid, cudaStream_t s) {
  BnArgs copy = a;
  void* params[1] = {&copy};
  cudaError_t e = cudaLaunchCooperativeKernel(
      reinterpret_cast<const void*>(&bn_fused_kernel<T, BWD>), dim3(grid), dim3(BN_FUSED_THREADS),
      params, 0, s);
  if (e != cudaSuccess) return (int)e;
  return (int)cudaGetLastError();
}

template <typename T>
int run_fused(const BnArgs& a, int backward, int grid, cudaStream_t s) {
  return backward ? launch_fused<T, true>(a, grid, s) : launch_fused<T, false>(a, grid, s);
}

}  // namespace

extern "C" int adl_bind_thread();

extern "C" {

int adl_sizeof_bn_args() { return (int)sizeof(BnArgs); }

// pdl: launch bn_apply as a programmatic dependent of bn_reduce (two-kernel path)
void adl_bn_config(int pdl) { g_bn_pdl = pdl; }

// Largest co-resident grid of the single-launch kernels on this device (CTAs), 0 if unsupported.
int adl_bn_fused_max_grid(int dev) {
  int coop = 0, sms = 0, per_sm = 0;
  if (cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev) != cudaSuccess || !coop) return 0;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, bn_fused_kernel<__nv_bfloat16, true>,
                                                    BN_FUSED_THREADS, 0) != cudaSuccess) return 0;
  return per_sm > 0 ? sms : 0;      // one CTA per SM
}

// Single cooperative launch per direction. `grid` CTAs (<= adl_bn_fused_max_grid), a.partial holds
// grid * 2 * C floats, a.counters two zeroed ints (left zero), C <= 1024.
int adl_bn_act_fused(const void* args, int dtype, int backward, int grid, void* stream) {
  const BnArgs* a = static_cast<const BnArgs*>(args);
  if (int rc = adl_bind_thread()) return rc;
  const int v = dtype == 0 ? 4 : 8;
  if (a->C % v != 0 || a->C > BN_FUSED_MAX_C) return -20;
  const int tpr = a->C / v;
  if (tpr > BN_FUSED_THREADS || BN_FUSED_THREADS % tpr != 0) return -21;
  if (grid <= 0 || a->n_partial != grid) return -22;
  if ((2 * a->C / 4) > BN_FUSED_THREADS || BN_FUSED_THREADS % (2 * a->C / 4) != 0) return -25;
  cudaStream_t s = (cudaStream_t)stream;
  switch (dtype) {
    case 0: return run_fused<float>(*a, backward, grid, s);
    case 1: return run_fused<__nv_bfloat16>(*a, backward, grid, s);
    case 2: return run_fused<__half>(*a, backward, grid, s);
  }
  return -23;
}

// dtype: 0 fp32, 1 bf16, 2 fp16. Requirements (checked by the caller too): C a multiple of
// the vector width (4 / 8) and of cb, cb <= 64, C / width and cb / width divisors of 256,
// a.n_partial == grid (row chunks of the reduction).
int adl_bn_act(const void* args, int dtype, int backward, int grid, int grid_apply, void* stream) {
  const BnArgs* a = static_cast<const BnArgs*>(args);
  if (int rc = adl_bind_thread()) return rc;
  const int v = dtype == 0 ? 4 : 8;
  if (a->C % v != 0) return -20;
  const int tpr = a->C / v;
  if (tpr > BN_THREADS || BN_THREADS % tpr != 0) return -21;
  if (grid <= 0 || grid_apply <= 0 || a->n_partial != grid) return -22;
  if (a->cb <= 0 || a->cb > 64 || a->C % a->cb != 0 || a->cb % v != 0 || BN_THREADS % (a->cb / v) != 0) return -24;
  cudaStream_t s = (cudaStream_t)stream;
  switch (dtype) {
    case 0: return run<float>(*a, backward, grid, grid_apply, s);
    case 1: return run<__nv_bfloat16>(*a, backward, grid, grid_apply, s);
    case 2: return run<__half>(*a, backward, grid, grid_apply, s);
  }
  return -23;
}

}  // extern "C"
