// adaptdl_b200 -- fused dropout + residual add + LayerNorm, forward and backward (sm_100a).
//
// The transformer workloads (examples/BERT, examples/transformer) close every sub-layer with
//     y = LayerNorm(x + dropout(h))
// Stock PyTorch runs it as dropout, add, layer_norm forward and layer_norm_grad_input,
// GammaBetaBackward, masked_scale, add backward: seven launches and ~36 bytes per element
// of HBM traffic per sub-layer. Here: ONE forward kernel (reads x, h and the keep-mask, writes
// y and the pre-norm sum z), ONE backward kernel (reads dy, z, mask; writes dx and dh and the
// per-CTA partial sums of dgamma / dbeta) and a small column reduction.
//
// One warp owns a row at a time: the row lives in registers between the statistics and the
// normalisation (two-pass variance), 16-byte vectors, lanes keep the same columns for every
// row they visit so the dgamma / dbeta partials accumulate in registers.
#include "adl_common.cuh"

namespace {

constexpr int LN_THREADS = 256;
constexpr int LN_WARPS = LN_THREADS / 32;
constexpr int LN_MAXV = 4;            // 16-byte vectors per lane (template NV <= 4): D <= 32 * 4 * (4 | 8)

struct LnArgs {
  const void* x;        // [M, D] residual input
  const void* h;        // [M, D] sub-layer output (dropout applies to it)        | backward: dy
  const uint8_t* mask;  // [M, D] 1 = keep (or null: no dropout)
  void* y;              // [M, D] LayerNorm output                               | backward: dx
  void* z;              // [M, D] x + dropout(h), saved for the backward pass     | backward: z (input)
  void* dh;             // backward: [M, D] gradient of h
  const float* gamma;   // [D]
  const float* beta;    // [D]
  float* mean;          // [M]
  float* rstd;          // [M]
  float* partial;       // backward: [grid, 2, D] (dgamma, dbeta) partial sums
  float* dgamma;        // [D]
  float* dbeta;         // [D]
  int M, D;
  int n_partial;
  float scale;          // 1 / (1 - p)
  float eps;
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// keep-mask bytes of one vector (V = 4 or 8 elements)
template <int V>
__device__ __forceinline__ void ld_mask(const uint8_t* p, float* keep) {
  if (V == 8) {
    const uint2 m = *reinterpret_cast<const uint2*>(p);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      keep[e] = (float)((m.x >> (8 * e)) & 0xffu);
      keep[4 + e] = (float)((m.y >> (8 * e)) & 0xffu);
    }
  } else {
    const uint32_t m = *reinterpret_cast<const uint32_t*>(p);
#pragma unroll
    for (int e = 0; e < 4; ++e) keep[e] = (float)((m >> (8 * e)) & 0xffu);
  }
}

template <typename T, int NV>
__global__ void __launch_bounds__(LN_THREADS)
ln_fwd_kernel(const LnArgs a) {
  constexpr int V = VecTraits<T>::N;
  // gamma / beta once per CTA into shared memory: a lane's columns are the same for every row it
  // visits, and 48 scalar global loads per row (stride 32 B across the warp: one 4-byte word used
  // per 32-byte sector) were 4/5 of this kernel's L1 traffic and its actual limiter (ncu:
  // profiles/r2_bert/ncu_ln_before.txt)
  __shared__ __align__(16) float s_gamma[32 * LN_MAXV * 8];
  __shared__ __align__(16) float s_beta[32 * LN_MAXV * 8];
  for (int c = threadIdx.x; c < a.D; c += LN_THREADS) { s_gamma[c] = a.gamma[c]; s_beta[c] = a.beta[c]; }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int nvec = a.D / V;
  const float inv_d = 1.f / (float)a.D;
  for (int row = blockIdx.x * LN_WARPS + warp; row < a.M; row += gridDim.x * LN_WARPS) {
    const size_t base = (size_t)row * a.D;
    float zr[NV][V];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int v = lane + 32 * i;
      if (v < nvec) {
        const size_t off = base + (size_t)v * V;
        float fx[V], fh[V];
        unpack<T>(ld_vec(static_cast<const T*>(a.x) + off), fx);
        if (a.h != nullptr) {
          unpack<T>(ld_vec(static_cast<const T*>(a.h) + off), fh);
          if (a.mask != nullptr) {
            float keep[V];
            ld_mask<V>(a.mask + off, keep);
#pragma unroll
            for (int e = 0; e < V; ++e) fh[e] *= keep[e] * a.scale;
          }
#pragma unroll
          for (int e = 0; e < V; ++e) fx[e] += fh[e];
        }
        // the sum is what the backward pass sees: round it like the stored copy
        const Vec16 packed = pack<T>(fx);
        if (a.z != nullptr) st_vec(static_cast<T*>(a.z) + off, packed);
        unpack<T>(packed, zr[i]);
#pragma unroll
        for (int e = 0; e < V; ++e) sum += zr[i][e];
      }
    }
    const float mean = warp_sum(sum) * inv_d;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      if (lane + 32 * i < nvec) {
#pragma unroll
        for (int e = 0; e < V; ++e) { const float d = zr[i][e] - mean; sq = fmaf(d, d, sq); }
      }
    }
    const float rstd = rsqrtf(warp_sum(sq) * inv_d + a.eps);
    if (lane == 0) { a.mean[row] = mean; a.rstd[row] = rstd; }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int v = lane + 32 * i;
      if (v < nvec) {
        float out[V], gm[V], bt[V];
#pragma unroll
        for (int e = 0; e < V; e += 4) {
          *reinterpret_cast<float4*>(gm + e) = *reinterpret_cast<const float4*>(s_gamma + v * V + e);
          *reinterpret_cast<float4*>(bt + e) = *reinterpret_cast<const float4*>(s_beta + v * V + e);
        }
#pragma unroll
        for (int e = 0; e < V; ++e) out[e] = fmaf((zr[i][e] - mean) * rstd, gm[e], bt[e]);
        st_vec(static_cast<T*>(a.y) + base + (size_t)v * V, pack<T>(out));
      }
    }
  }
}

// dz = rstd * (g - mean(g) - xhat * mean(g * xhat)),  g = dy * gamma
// dx = dz,  dh = dz * keep * scale;  partial dgamma += dy * xhat, dbeta += dy
template <typename T, int NV>
__global__ void __launch_bounds__(LN_THREADS)
ln_bwd_kernel(const LnArgs a) {
  constexpr int V = VecTraits<T>::N;
  // gamma staged in shared memory (see ln_fwd_kernel); `slab` is the fold scratch, one row per warp
  __shared__ __align__(16) float s_gamma[32 * LN_MAXV * 8];
  __shared__ __align__(16) float slab[LN_WARPS][32 * LN_MAXV * 8];   // 32 KB
  for (int c = threadIdx.x; c < a.D; c += LN_THREADS) s_gamma[c] = a.gamma[c];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int nvec = a.D / V;
  const float inv_d = 1.f / (float)a.D;
  float dg[NV][V], db[NV][V];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
#pragma unroll
    for (int e = 0; e < V; ++e) { dg[i][e] = 0.f; db[i][e] = 0.f; }
  }
  for (int row = blockIdx.x * LN_WARPS + warp; row < a.M; row += gridDim.x * LN_WARPS) {
    const size_t base = (size_t)row * a.D;
    const float mean = a.mean[row], rstd = a.rstd[row];
    float xh[NV][V], g[NV][V];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int v = lane + 32 * i;
      if (v < nvec) {
        const size_t off = base + (size_t)v * V;
        float fz[V], fdy[V], gm[V];
        unpack<T>(ld_vec(static_cast<const T*>(a.z) + off), fz);
        unpack<T>(ld_vec(static_cast<const T*>(a.h) + off), fdy);   // a.h carries dy
#pragma unroll
        for (int e = 0; e < V; e += 4)
          *reinterpret_cast<float4*>(gm + e) = *reinterpret_cast<const float4*>(s_gamma + v * V + e);
#pragma unroll
        for (int e = 0; e < V; ++e) {
          xh[i][e] = (fz[e] - mean) * rstd;
          g[i][e] = fdy[e] * gm[e];
          s1 += g[i][e];
          s2 = fmaf(g[i][e], xh[i][e], s2);
          dg[i][e] = fmaf(fdy[e], xh[i][e], dg[i][e]);
          db[i][e] += fdy[e];
        }
      }
    }
    const float c1 = warp_sum(s1) * inv_d, c2 = warp_sum(s2) * inv_d;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int v = lane + 32 * i;
      if (v < nvec) {
        const size_t off = base + (size_t)v * V;
        float dz[V];
#pragma unroll
        for (int e = 0; e < V; ++e) dz[e] = rstd * (g[i][e] - c1 - xh[i][e] * c2);
        st_vec(static_cast<T*>(a.y) + off, pack<T>(dz));              // a.y carries dx
        if (a.dh != nullptr) {
          if (a.mask != nullptr) {
            float keep[V];
            ld_mask<V>(a.mask + off, keep);
#pragma unroll
            for (int e = 0; e < V; ++e) dz[e] *= keep[e] * a.scale;
          }
          st_vec(static_cast<T*>(a.dh) + off, pack<T>(dz));
        }
      }
    }
  }
  // fold the eight warps' partials in a fixed order and store this CTA's row of partial sums:
  // every warp drops its registers into its own shared-memory row (plain stores), then all
  // threads add the eight rows column by column. (The first version let the warps take turns
  // doing read-modify-writes on one row: 15 of the 26 stall cycles per instruction were that
  // barrier, profiles/r2_bert/ncu_ln_before.txt.)
  float* out = a.partial + (size_t)blockIdx.x * 2 * a.D;
#pragma unroll
  for (int pass = 0; pass < 2; ++pass) {
    __syncthreads();                                   // previous pass has been read out
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int v = lane + 32 * i;
      if (v < nvec) {
#pragma unroll
        for (int e = 0; e < V; e += 4) {
          const float* src = pass == 0 ? &dg[i][e] : &db[i][e];
          *reinterpret_cast<float4*>(&slab[warp][v * V + e]) = make_float4(src[0], src[1], src[2], src[3]);
        }
      }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < a.D; c += LN_THREADS) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < LN_WARPS; ++w) t += slab[w][c];
      out[pass * a.D + c] = t;
    }
  }
}

// column sums of the per-CTA partials [n_partial][2 * D]: a CTA owns 64 consecutive floats of the
// 2*D-wide row (16 float4 columns x 16 row lanes), every thread keeps EIGHT rows in flight (the
// first version walked its rows one dependent load at a time: 23 us for 1.8 MB, pure latency),
// the row lanes are combined through shared memory in a fixed order (deterministic).
__global__ void __launch_bounds__(LN_THREADS)
ln_param_grad_kernel(const LnArgs a) {
  constexpr int QC = 16;                               // float4 columns per CTA
  constexpr int LANES = LN_THREADS / QC;               // 16 row lanes
  __shared__ float4 red[LANES][QC];
  const int qi = threadIdx.x % QC, li = threadIdx.x / QC;
  const int q = 2 * a.D / 4;                           // float4 columns of a partial row
  const int col4 = blockIdx.x * QC + qi;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (col4 < q) {
    const float4* base = reinterpret_cast<const float4*>(a.partial) + col4;
    for (int p0 = li; p0 < a.n_partial; p0 += 8 * LANES) {
      float4 t[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int p = p0 + j * LANES;
        t[j] = (p < a.n_partial) ? __ldcg(base + (size_t)p * q) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) { acc.x += t[j].x; acc.y += t[j].y; acc.z += t[j].z; acc.w += t[j].w; }
    }
  }
  red[li][qi] = acc;
  __syncthreads();
  if (threadIdx.x < QC && col4 < q) {
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int l = 0; l < LANES; ++l) { const float4 r = red[l][qi]; t.x += r.x; t.y += r.y; t.z += r.z; t.w += r.w; }
    const int c = col4 * 4;                            // position in the [dgamma | dbeta] row
    float* dst = (c < a.D) ? (a.dgamma + c) : (a.dbeta + (c - a.D));
    *reinterpret_cast<float4*>(dst) = t;
  }
}

template <typename T, int NV>
void launch(const LnArgs& a, int backward, int grid, cudaStream_t s) {
  if (!backward) {
    ln_fwd_kernel<T, NV><<<grid, LN_THREADS, 0, s>>>(a);
  } else {
    ln_bwd_kernel<T, NV><<<grid, LN_THREADS, 0, s>>>(a);
    ln_param_grad_kernel<<<(2 * a.D / 4 + 15) / 16, LN_THREADS, 0, s>>>(a);
  }
}

template <typename T>
int run(const LnArgs& a, int backward, int grid, cudaStream_t s) {
  const int nv = (a.D / VecTraits<T>::N + 31) / 32;    // vectors per lane
  switch (nv) {
    case 1: launch<T, 1>(a, backward, grid, s); break;
    case 2: launch<T, 2>(a, backward, grid, s); break;
    case 3: launch<T, 3>(a, backward, grid, s); break;
    case 4: launch<T, 4>(a, backward, grid, s); break;
    default: return -33;
  }
  return (int)cudaGetLastError();
}

}  // namespace

extern "C" int adl_bind_thread();

extern "C" {

int adl_sizeof_ln_args() { return (int)sizeof(LnArgs); }

// dtype: 0 fp32, 1 bf16, 2 fp16. D must be a multiple of the vector width (4 / 8) and at most
// 32 * 4 vectors (D <= 1024 for bf16 / fp16, 512 for fp32); backward: n_partial == grid.
int adl_dropout_add_ln(const void* args, int dtype, int backward, int grid, void* stream) {
  const LnArgs* a = static_cast<const LnArgs*>(args);
  if (int rc = adl_bind_thread()) return rc;
  const int v = dtype == 0 ? 4 : 8;
  if (a->D % v != 0 || a->D / v > 32 * LN_MAXV) return -30;
  if (grid <= 0 || (backward && a->n_partial != grid)) return -31;
  cudaStream_t s = (cudaStream_t)stream;
  switch (dtype) {
    case 0: return run<float>(*a, backward, grid, s);
    case 1: return run<__nv_bfloat16>(*a, backward, grid, s);
    case 2: return run<__half>(*a, backward, grid, s);
  }
  return -32;
}

}  // extern "C"
