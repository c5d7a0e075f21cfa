// adaptdl_b200 -- fused optimizer step over a flat gradient arena (sm_100a).
//
// One launch updates EVERY parameter of an arena (reference call site K12:
// torch.optim runs ~4-5 launches per param group, 62 groups for the CIFAR
// workload). Gradients are read from the flat arena the fused all-reduce just
// wrote; parameters are reached through a per-segment pointer table (they
// stay ordinary, separately allocated nn.Parameters); optimizer state
// (momentum / Adam moments) lives in flat arenas whose views are exposed to
// torch as optimizer.state[...] for checkpoint compatibility.
//
// 16-bit parameters (bf16 / fp16 arenas) are updated in MIXED PRECISION: an fp32 master copy
// of the weights and fp32 optimizer state live in flat arenas of the same layout (`master`,
// state0/state1 then hold floats); the kernel reads the 16-bit gradient, updates master and
// state in fp32 and writes the rounded 16-bit parameter the forward pass uses -- no per-tensor
// cast kernels anywhere in the step.
//
// The learning rate of group g is  hyper[g].lr * lr_factor[g]  where
// lr_factor is written ON THE DEVICE by the gradient-noise-scale estimator
// (adl_finalize_stats) -- the AdaScale gain never round-trips through Python,
// so the whole training step is host-sync free and CUDA-graph capturable.
#include "adl_common.cuh"

#define ADL_HYPER_STRIDE 8
// hyper[g*8 + k]: 0 lr, 1 momentum|beta1, 2 weight_decay, 3 nesterov|adamw flag,
//                 4 beta2, 5 eps, 6 (unused), 7 (unused)

struct OptimArgs {
  const void* grad;                      // flat gradient arena
  void* state0;                          // momentum buffer | exp_avg   (flat, same layout)
  void* state1;                          // exp_avg_sq                  (Adam only)
  const unsigned long long* param_ptr;   // [n_seg] base address of each parameter's storage
  const int* seg_start;                  // [n_seg] first arena element of the segment
  const int* seg_numel;                  // [n_seg] elements in the parameter
  SegTable segs;                         // seg_end (vectors), seg_group
  int n_vec;                             // vectors in the arena
  const float* hyper;                    // [n_groups][ADL_HYPER_STRIDE]
  const float* lr_factor;                // [n_groups + 1] (last = finite flag) or nullptr
  const uint32_t* step_ctr;              // optimizer steps finalized (Adam bias correction)
  const int* step_offset;                // device int: adam_step = *step_ctr + *step_offset
  int n_groups;
  float* master;                         // fp32 master weights, arena layout (16-bit arenas only)
  const float* grad_scale;               // device: AMP loss scale carried by the gradients (or nullptr)
  float* pinv_coef;                      // Adam: [n_groups][2] preconditioner coefficients for the NEXT
                                         // backward's statistics (adl_kernels.cu PINV mode 2), or nullptr
};

template <typename T> __device__ __forceinline__ float to_f(T x);
template <> __device__ __forceinline__ float to_f<float>(float x) { return x; }
template <> __device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 x) { return __bfloat162float(x); }
template <> __device__ __forceinline__ float to_f<__half>(__half x) { return __half2float(x); }
template <typename T> __device__ __forceinline__ T from_f(float x);
template <> __device__ __forceinline__ float from_f<float>(float x) { return x; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float x) { return __float2bfloat16_rn(x); }
template <> __device__ __forceinline__ __half from_f<__half>(float x) { return __float2half_rn(x); }

// N consecutive floats of a flat fp32 array starting at element e0 (e0 % N == 0)
template <int N>
__device__ __forceinline__ void ld_f32(const float* base, int e0, float* out) {
#pragma unroll
  for (int i = 0; i < N; i += 4) unpack<float>(ld_vec(base + e0 + i), out + i);
}
template <int N>
__device__ __forceinline__ void st_f32(float* base, int e0, const float* in) {
#pragma unroll
  for (int i = 0; i < N; i += 4) st_vec(base + e0 + i, pack<float>(in + i));
}

// ADAM: 0 = SGD(momentum, nesterov), 1 = Adam / AdamW. WIDE: fp32 master weights + fp32 state
// next to 16-bit parameters / gradients.
template <typename T, int ADAM, bool WIDE>
__global__ void __launch_bounds__(ADL_THREADS, 2) fused_optim_kernel(const OptimArgs a) {
  constexpr int N = VecTraits<T>::N;
  if (a.lr_factor && a.lr_factor[a.n_groups] == 0.f) return;   // non-finite gradients: skip
  const int stride = gridDim.x * blockDim.x;
  float bc1 = 1.f, bc2_rsqrt = 1.f;
  float step = 0.f;
  if (ADAM) step = (float)((long long)(*a.step_ctr) + (long long)(*a.step_offset));
  const float inv_gs = a.grad_scale ? 1.f / *a.grad_scale : 1.f;
  if (ADAM && a.pinv_coef && blockIdx.x == 0) {
    // Adam-preconditioned gradient statistics (reference gradient_noise_scale.py:289-311):
    // after `step` updates the divisor is sqrt(v / (1 - beta2^step)) + eps, and none for
    // the first five steps
    for (int g = threadIdx.x; g < a.n_groups; g += blockDim.x) {
      const float* h = a.hyper + g * ADL_HYPER_STRIDE;
      a.pinv_coef[2 * g] = step >= 5.f ? rsqrtf(1.f - powf(h[4], step)) : 0.f;
      a.pinv_coef[2 * g + 1] = h[5];
    }
  }
  int cur = -1;
  for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < a.n_vec; v += stride) {
    if (cur < 0) cur = seg_find(a.segs, v);
    while (__ldg(a.segs.seg_end + cur) <= v) ++cur;
    const int e0 = v * N;
    const int s_start = __ldg(a.seg_start + cur);
    const int s_numel = __ldg(a.seg_numel + cur);
    const int rel = e0 - s_start;
    if (rel < 0 || rel >= s_numel) continue;             // padding between segments
    const int valid = min(N, s_numel - rel);
    const int g = __ldg(a.segs.seg_group + cur);
    const float* h = a.hyper + g * ADL_HYPER_STRIDE;
    const float lr = h[0] * (a.lr_factor ? a.lr_factor[g] : 1.f);
    const float wd = h[2];
    T* pbase = reinterpret_cast<T*>(__ldg(a.param_ptr + cur)) + rel;
    const bool vec_ok = valid == N && ((reinterpret_cast<uintptr_t>(pbase) & 15) == 0);

    float gr[N], p[N], s0[N], s1[N];
    unpack<T>(ld_vec(static_cast<const Vec16*>(a.grad) + v), gr);
    if (a.grad_scale) {
#pragma unroll
      for (int e = 0; e < N; ++e) gr[e] *= inv_gs;
    }
    if (a.state0) {
      if (WIDE) ld_f32<N>(static_cast<const float*>(a.state0), e0, s0);
      else unpack<T>(ld_vec(static_cast<const Vec16*>(a.state0) + v), s0);
    } else {
#pragma unroll
      for (int e = 0; e < N; ++e) s0[e] = 0.f;
    }
    if (ADAM) {
      if (WIDE) ld_f32<N>(static_cast<const float*>(a.state1), e0, s1);
      else unpack<T>(ld_vec(static_cast<const Vec16*>(a.state1) + v), s1);
    }
    if (WIDE) {
      ld_f32<N>(a.master, e0, p);
    } else if (vec_ok) {
      unpack<T>(*reinterpret_cast<const Vec16*>(pbase), p);
    } else {
#pragma unroll
      for (int e = 0; e < N; ++e) p[e] = e < valid ? to_f<T>(pbase[e]) : 0.f;
    }
    if (!ADAM) {
      const float mom = h[1];
      const bool nesterov = h[3] != 0.f;
#pragma unroll
      for (int e = 0; e < N; ++e) {
        float d = gr[e];
        if (wd != 0.f) d = fmaf(wd, p[e], d);
        if (mom != 0.f) {
          s0[e] = fmaf(mom, s0[e], d);
          d = nesterov ? fmaf(mom, s0[e], d) : s0[e];
        }
        p[e] = fmaf(-lr, d, p[e]);
      }
    } else {
      const float b1 = h[1], b2 = h[4], eps = h[5];
      const bool adamw = h[3] != 0.f;
      bc1 = 1.f - powf(b1, step);
      bc2_rsqrt = rsqrtf(1.f - powf(b2, step));
      const float step_size = lr / bc1;
#pragma unroll
      for (int e = 0; e < N; ++e) {
        float d = gr[e];
        if (wd != 0.f) {
          if (adamw) p[e] *= (1.f - lr * wd); else d = fmaf(wd, p[e], d);
        }
        s0[e] = fmaf(b1, s0[e], (1.f - b1) * d);
        s1[e] = fmaf(b2, s1[e], (1.f - b2) * d * d);
        const float denom = fmaf(sqrtf(s1[e]), bc2_rsqrt, eps);
        p[e] = fmaf(-step_size, s0[e] / denom, p[e]);
      }
    }
    if (WIDE) st_f32<N>(a.master, e0, p);
    if (vec_ok) {
      *reinterpret_cast<Vec16*>(pbase) = pack<T>(p);
    } else {
#pragma unroll
      for (int e = 0; e < N; ++e) if (e < valid) pbase[e] = from_f<T>(p[e]);
    }
    // state arenas share the gradient arena's layout (padding included)
    if (!ADAM) {
      if (h[1] != 0.f && a.state0) {
        if (WIDE) st_f32<N>(static_cast<float*>(a.state0), e0, s0);
        else st_vec(static_cast<Vec16*>(a.state0) + v, pack<T>(s0));
      }
    } else if (WIDE) {
      st_f32<N>(static_cast<float*>(a.state0), e0, s0);
      st_f32<N>(static_cast<float*>(a.state1), e0, s1);
    } else {
      st_vec(static_cast<Vec16*>(a.state0) + v, pack<T>(s0));
      st_vec(static_cast<Vec16*>(a.state1) + v, pack<T>(s1));
    }
  }
}

extern "C" int adl_bind_thread();

extern "C" {

int adl_fused_optim(const OptimArgs* args, int adam, int dtype, int grid, void* stream) {
  if (int rc = adl_bind_thread()) return rc;
  cudaStream_t s = (cudaStream_t)stream;
#define LAUNCH_O(T, WIDE)                                                              \
  do {                                                                                 \
    if (adam) fused_optim_kernel<T, 1, WIDE><<<grid, ADL_THREADS, 0, s>>>(*args);      \
    else fused_optim_kernel<T, 0, WIDE><<<grid, ADL_THREADS, 0, s>>>(*args);           \
  } while (0)
  const bool wide = args->master != nullptr;
  if (wide && dtype == 0) return -3;
  if (dtype == 0) LAUNCH_O(float, false);
  else if (dtype == 1) { if (wide) LAUNCH_O(__nv_bfloat16, true); else LAUNCH_O(__nv_bfloat16, false); }
  else if (dtype == 2) { if (wide) LAUNCH_O(__half, true); else LAUNCH_O(__half, false); }
  else return -2;
#undef LAUNCH_O
  return (int)cudaGetLastError();
}

// Adam's step counter advances only on finite (applied) updates
__global__ void optim_advance_kernel(int* steps, const float* finite) {
  if (finite == nullptr || *finite != 0.f) *steps += 1;
}

int adl_optim_advance(int* steps, const float* finite, void* stream) {
  if (int rc = adl_bind_thread()) return rc;
  optim_advance_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(steps, finite);
  return (int)cudaGetLastError();
}

int adl_sizeof_optim_args() { return (int)sizeof(OptimArgs); }

}  // extern "C"
