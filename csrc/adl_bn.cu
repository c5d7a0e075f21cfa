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
  float* dgamma;        // [C]
  float* dbeta;         // [C]
  float* partial;       // [grid, 2, C] scratch
  float* coef;          // [2, C] scratch: forward (scale, shift); backward (s1/M, s2/M)
  int M, C;
  int n_partial;
  int relu;
  float eps, momentum;
};

// sum over rows of two per-channel quantities, K in {fwd: x, x^2 ; bwd: g, g*xhat}
template <typename T, bool BWD>
__global__ void __launch_bounds__(BN_THREADS)
bn_reduce_kernel(const BnArgs a) {
  constexpr int V = VecTraits<T>::N;
  constexpr int BN_UNROLL = BnUnroll<BWD>::U;
  extern __shared__ float red[];                 // [2][rows_per_iter][C]
  const int tpr = a.C / V;                       // threads per row
  const int rpi = BN_THREADS / tpr;              // rows per iteration
  const int my_c = (threadIdx.x % tpr) * V;
  const int my_r = threadIdx.x / tpr;
  float s0[V], s1[V];
#pragma unroll
  for (int e = 0; e < V; ++e) { s0[e] = 0.f; s1[e] = 0.f; }
  float mu[V], rs[V];
  if (BWD) {
#pragma unroll
    for (int e = 0; e < V; ++e) { mu[e] = a.mean[my_c + e]; rs[e] = a.rstd[my_c + e]; }
  }
  const long long stride = (long long)gridDim.x * rpi;
  for (long long r0 = (long long)blockIdx.x * rpi + my_r; r0 < a.M; r0 += stride * BN_UNROLL) {
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
  float* r1s = red + rpi * a.C;
#pragma unroll
  for (int e = 0; e < V; ++e) {
    r0s[my_r * a.C + my_c + e] = s0[e];
    r1s[my_r * a.C + my_c + e] = s1[e];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < 2 * a.C; c += BN_THREADS) {
    const float* src = (c < a.C) ? (r0s + c) : (r1s + (c - a.C));
    float t = 0.f;
    for (int r = 0; r < rpi; ++r) t += src[r * a.C];
    a.partial[(size_t)blockIdx.x * 2 * a.C + c] = t;
  }
}

// one thread per channel: fold the partials (fixed order) and derive the per-channel terms
template <bool BWD>
__global__ void bn_finalize_kernel(const BnArgs a) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= a.C) return;
  double t0 = 0.0, t1 = 0.0;
  for (int p = 0; p < a.n_partial; ++p) {
    t0 += (double)a.partial[(size_t)p * 2 * a.C + c];
    t1 += (double)a.partial[(size_t)p * 2 * a.C + a.C + c];
  }
  const double inv_m = 1.0 / (double)a.M;
  if (!BWD) {
    const double mean = t0 * inv_m;
    double var = t1 * inv_m - mean * mean;
    if (var < 0.0) var = 0.0;
    const float rstd = (float)(1.0 / sqrt(var + (double)a.eps));
    a.mean[c] = (float)mean;
    a.rstd[c] = rstd;
    const float scale = a.gamma[c] * rstd;
    a.coef[c] = scale;
    a.coef[a.C + c] = a.beta[c] - (float)mean * scale;
    if (a.running_mean != nullptr) {
      const double unbiased = a.M > 1 ? var * (double)a.M / (double)(a.M - 1) : var;
      a.running_mean[c] = (1.f - a.momentum) * a.running_mean[c] + a.momentum * (float)mean;
      a.running_var[c] = (1.f - a.momentum) * a.running_var[c] + a.momentum * (float)unbiased;
    }
  } else {
    a.dbeta[c] = (float)t0;
    a.dgamma[c] = (float)t1;
    a.coef[c] = (float)(t0 * inv_m);
    a.coef[a.C + c] = (float)(t1 * inv_m);
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
int run(const BnArgs& a, int backward, int grid, cudaStream_t s) {
  constexpr int V = VecTraits<T>::N;
  const int tpr = a.C / V;
  const int rpi = BN_THREADS / tpr;
  const size_t smem = (size_t)2 * rpi * a.C * sizeof(float);
  const int fin_blocks = (a.C + 127) / 128;
  if (!backward) {
    bn_reduce_kernel<T, false><<<grid, BN_THREADS, smem, s>>>(a);
    bn_finalize_kernel<false><<<fin_blocks, 128, 0, s>>>(a);
    bn_apply_kernel<T, false><<<grid, BN_THREADS, 0, s>>>(a);
  } else {
    bn_reduce_kernel<T, true><<<grid, BN_THREADS, smem, s>>>(a);
    bn_finalize_kernel<true><<<fin_blocks, 128, 0, s>>>(a);
    bn_apply_kernel<T, true><<<grid, BN_THREADS, 0, s>>>(a);
  }
  return (int)cudaGetLastError();
}

}  // namespace

extern "C" int adl_bind_thread();

extern "C" {

int adl_sizeof_bn_args() { return (int)sizeof(BnArgs); }

// dtype: 0 fp32, 1 bf16, 2 fp16. Requirements (checked by the caller too): C a multiple of
// the vector width (4 / 8), C / width a divisor of 256, a.n_partial == grid.
int adl_bn_act(const void* args, int dtype, int backward, int grid, void* stream) {
  const BnArgs* a = static_cast<const BnArgs*>(args);
  if (int rc = adl_bind_thread()) return rc;
  const int v = dtype == 0 ? 4 : 8;
  if (a->C % v != 0) return -20;
  const int tpr = a->C / v;
  if (tpr > BN_THREADS || BN_THREADS % tpr != 0) return -21;
  if (grid <= 0 || a->n_partial != grid) return -22;
  cudaStream_t s = (cudaStream_t)stream;
  switch (dtype) {
    case 0: return run<float>(*a, backward, grid, s);
    case 1: return run<__nv_bfloat16>(*a, backward, grid, s);
    case 2: return run<__half>(*a, backward, grid, s);
  }
  return -23;
}

}  // extern "C"
