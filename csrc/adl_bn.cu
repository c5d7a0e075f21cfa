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
template <bool BWD> struct BnUnroll { static constexpr int U = BWD ? 3 : 6; };

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
  // fold the grid.y partials of this channel block (fixed order: deterministic). The
  // [grid.y][2*cb] matrix is read as float4 columns: thread = (row lane, column), EIGHT rows in
  // flight per thread, so the whole fold is one or two L2 round trips instead of a chain of
  // grid.y / 32 dependent ones; the row lanes are then combined through shared memory.
  const int nv4 = 2 * cb / 4;                    // float4 columns per partial (<= 32)
  const int lanes = BN_THREADS / nv4;            // row lanes (>= 8)
  const int qi = threadIdx.x % nv4, li = threadIdx.x / nv4;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  {
    const float4* base = reinterpret_cast<const float4*>(a.partial + (size_t)blockIdx.x * gridDim.y * 2 * cb) + qi;
    const int gy = (int)gridDim.y;
    for (int p0 = li; p0 < gy; p0 += 8 * lanes) {
      float4 t[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int p = p0 + j * lanes;
        t[j] = (p < gy) ? __ldcg(base + (size_t)p * nv4) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) { acc.x += t[j].x; acc.y += t[j].y; acc.z += t[j].z; acc.w += t[j].w; }
    }
  }
  __syncthreads();                               // everyone is done with `red`
  reinterpret_cast<float4*>(red)[li * nv4 + qi] = acc;          // [lanes][2*cb] floats (<= 16 KB)
  __syncthreads();
  __shared__ float tot[2 * 64];                  // [2][cb]
  if (threadIdx.x < 2 * cb) {
    double t = 0.0;
    for (int l = 0; l < lanes; ++l) t += (double)red[l * 2 * cb + threadIdx.x];
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

template <typename T>
int run(const BnArgs& a, int backward, int grid_y, int grid_apply, cudaStream_t s) {
  const dim3 grid(a.C / a.cb, grid_y);
  if (!backward) {
    bn_reduce_kernel<T, false><<<grid, BN_THREADS, 0, s>>>(a);
    bn_apply_kernel<T, false><<<grid_apply, BN_THREADS, 0, s>>>(a);
  } else {
    bn_reduce_kernel<T, true><<<grid, BN_THREADS, 0, s>>>(a);
    bn_apply_kernel<T, true><<<grid_apply, BN_THREADS, 0, s>>>(a);
  }
  return (int)cudaGetLastError();
}


}  // namespace

extern "C" int adl_bind_thread();

extern "C" {

int adl_sizeof_bn_args() { return (int)sizeof(BnArgs); }

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
