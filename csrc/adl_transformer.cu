// adaptdl_b200 -- data-movement kernels of the transformer workloads (sm_100a).
//
// A BERT-base step (examples/BERT, seq 128, batch 32) spends ~2.5 ms of its 10.7 ms in
// PyTorch's generic strided-copy / cat / reduce kernels at layout boundaries, each running at
// 5-25 % of the copy bandwidth (profiles/r2_bert/op_profile_bert.log):
//
//   qkv [N,S,3,H,D] <-> q,k,v [N,H,S,D]      aten::cat 47 us + aten::copy_ 26 us per layer (backward)
//   attention out [N,H,S,D] -> [N,S,H*D]      aten::copy_ 60 us per layer
//   bias gradients  sum_rows dY[M,N]          aten::sum 15-19 us, four per layer
//   MLM logits [M, 29056] bf16 (padded) -> [M, 28996] (user-visible) and back: 4 x 211-263 us
//
// These are pure bandwidth problems with friendly geometry (128-byte head rows, 16-byte aligned
// matrix rows); the kernels below move 16-byte vectors with every load and store coalesced:
//
//   adl_heads_split   qkv[N,S,W*H,D]  -> W tensors [N,H,S,D]      (W = 3; W = 1: a plain transpose)
//   adl_heads_merge   W tensors [N,H,S,D] -> [N,S,W*H,D]           (the inverse; attention output with W = 1)
//   adl_colsum        out[N] = sum_m x[m, n]  (fp32 accumulation, deterministic two-level fold)
//   adl_slice_cast    bf16 [M, ld] (first n columns) <-> fp32 [M, n] contiguous
#include "adl_common.cuh"

namespace {

constexpr int TR_THREADS = 256;

// src[a][b][w*H + h][v] -> dst[w][a][h][b][v]     (v = 16-byte vectors of one head row)
__global__ void __launch_bounds__(TR_THREADS)
heads_split_kernel(const Vec16* __restrict__ src, Vec16* __restrict__ dst, int A, int B, int W, int H, int vecs) {
  const long long total = (long long)A * B * W * H * vecs;
  const long long plane = (long long)A * H * B * vecs;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < total; i0 += 4 * stride) {
    Vec16 val[4];
    long long idx[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      idx[u] = i0 + u * stride;
      if (idx[u] < total) val[u] = ld_vec(src + idx[u]);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (idx[u] < total) {
        long long r = idx[u];
        const int v = (int)(r % vecs); r /= vecs;
        const int c = (int)(r % (W * H)); r /= (W * H);
        const int b = (int)(r % B);
        const int a = (int)(r / B);
        const int w = c / H, h = c - w * H;
        st_vec(dst + w * plane + (((long long)a * H + h) * B + b) * vecs + v, val[u]);
      }
    }
  }
}

// W source planes (each [a][h][b][v], separately allocated) -> dst[a][b][w*H + h][v]
struct Planes { const Vec16* p[4]; };

__global__ void __launch_bounds__(TR_THREADS)
heads_merge_kernel(const Planes src, Vec16* __restrict__ dst, int A, int B, int W, int H, int vecs) {
  const long long total = (long long)A * B * W * H * vecs;
  const long long stride = (long long)gridDim.x * blockDim.x;
  // iterate in DESTINATION order (coalesced stores; the loads are whole 128-byte head rows)
  for (long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < total; i0 += 4 * stride) {
    Vec16 val[4];
    long long idx[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      idx[u] = i0 + u * stride;
      if (idx[u] < total) {
        long long r = idx[u];
        const int v = (int)(r % vecs); r /= vecs;
        const int c = (int)(r % (W * H)); r /= (W * H);
        const int b = (int)(r % B);
        const int a = (int)(r / B);
        const int w = c / H, h = c - w * H;
        val[u] = ld_vec(src.p[w] + (((long long)a * H + h) * B + b) * vecs + v);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (idx[u] < total) st_vec(dst + idx[u], val[u]);
  }
}

// ---------------------------------------------------------------------------
// column sums of a row-major [M, N] matrix (bias gradients), fp32 accumulation.
// grid = (column blocks of 64 columns, row chunks); every CTA writes a partial row; the last
// CTA of a column block (ticket) folds the chunks in a fixed order (deterministic).
// ---------------------------------------------------------------------------
constexpr int CS_COLS = 64;                       // columns per CTA (8 vectors of 8 bf16)

template <typename T>
__global__ void __launch_bounds__(TR_THREADS)
colsum_kernel(const T* __restrict__ x, float* __restrict__ out, float* __restrict__ partial,
              int* __restrict__ tickets, int M, int N) {
  constexpr int V = VecTraits<T>::N;
  constexpr int TPR = CS_COLS / V;                // threads per row
  constexpr int RPI = TR_THREADS / TPR;           // rows per iteration
  __shared__ float red[RPI][CS_COLS + 1];
  __shared__ int is_last;
  const int c0 = blockIdx.x * CS_COLS;
  const int lc = (threadIdx.x % TPR) * V;
  const int my_r = threadIdx.x / TPR;
  const bool col_ok = c0 + lc < N;                // N is a multiple of V
  float s[V];
#pragma unroll
  for (int e = 0; e < V; ++e) s[e] = 0.f;
  const int stride = gridDim.y * RPI;
  for (int r0 = blockIdx.y * RPI + my_r; r0 < M; r0 += 4 * stride) {
    Vec16 val[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int r = r0 + u * stride;
      if (r < M && col_ok) val[u] = ld_vec(x + (size_t)r * N + c0 + lc);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int r = r0 + u * stride;
      if (r < M && col_ok) {
        float f[V];
        unpack<T>(val[u], f);
#pragma unroll
        for (int e = 0; e < V; ++e) s[e] += f[e];
      }
    }
  }
#pragma unroll
  for (int e = 0; e < V; ++e) red[my_r][lc + e] = s[e];
  __syncthreads();
  float* mine = partial + ((size_t)blockIdx.x * gridDim.y + blockIdx.y) * CS_COLS;
  if (threadIdx.x < CS_COLS) {
    float t = 0.f;
    for (int r = 0; r < RPI; ++r) t += red[r][threadIdx.x];
    mine[threadIdx.x] = t;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) is_last = (atomicAdd(tickets + blockIdx.x, 1) == (int)gridDim.y - 1);
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  // fold gridDim.y partial rows of 64 floats: thread = (row lane, column), 8 in flight
  const int col = threadIdx.x % CS_COLS, lane = threadIdx.x / CS_COLS;   // 4 row lanes
  constexpr int LANES = TR_THREADS / CS_COLS;
  const float* base = partial + (size_t)blockIdx.x * gridDim.y * CS_COLS + col;
  float acc = 0.f;
  for (int p0 = lane; p0 < (int)gridDim.y; p0 += 8 * LANES) {
    float t[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int p = p0 + j * LANES;
      t[j] = (p < (int)gridDim.y) ? __ldcg(base + (size_t)p * CS_COLS) : 0.f;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) acc += t[j];
  }
  __syncthreads();
  red[lane][col] = acc;
  __syncthreads();
  if (threadIdx.x < CS_COLS && c0 + threadIdx.x < N) {
    float t = 0.f;
    for (int l = 0; l < LANES; ++l) t += red[l][threadIdx.x];
    out[c0 + threadIdx.x] = t;
  }
  if (threadIdx.x == 0) tickets[blockIdx.x] = 0;
}

// ---------------------------------------------------------------------------
// padded bf16 logits <-> contiguous fp32 logits
//   to_f32 : dst[m][j] = float(src[m * ld + j])                j < n
//   to_bf16: dst[m * ld + j] = bf16(src[m][j]) (j < n), 0 (n <= j < ld)
// n * 4 bytes and ld * 2 bytes are multiples of 16 (the caller checks), so a thread moves eight
// elements with one 16-byte access on the bf16 side and two on the fp32 side; the last group
// of a row may be partial (n % 8 == 4).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(TR_THREADS)
slice_to_f32_kernel(const __nv_bfloat16* __restrict__ src, float* __restrict__ dst, int M, int n, int ld) {
  const int groups = (n + 7) / 8;
  const long long total = (long long)M * groups;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int m = (int)(i / groups), g = (int)(i - (long long)m * groups);
    const int j = g * 8;
    float f[8];
    unpack<__nv_bfloat16>(ld_vec(src + (size_t)m * ld + j), f);      // ld >= n rounded up to 8
    float* d = dst + (size_t)m * n + j;
    st_vec(d, pack<float>(f));
    if (j + 8 <= n) st_vec(d + 4, pack<float>(f + 4));
  }
}

__global__ void __launch_bounds__(TR_THREADS)
f32_to_slice_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int M, int n, int ld) {
  const int groups = ld / 8;
  const long long total = (long long)M * groups;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int m = (int)(i / groups), g = (int)(i - (long long)m * groups);
    const int j = g * 8;
    float f[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) f[e] = 0.f;
    const float* s = src + (size_t)m * n + j;
    if (j + 4 <= n) unpack<float>(ld_vec(s), f);
    if (j + 8 <= n) unpack<float>(ld_vec(s + 4), f + 4);
    st_vec(dst + (size_t)m * ld + j, pack<__nv_bfloat16>(f));
  }
}

// ---------------------------------------------------------------------------
// dZ = dD * keep * scale * gelu'(Z)      (backward of  D = dropout(gelu(Z)) in one pass)
// gelu'(z) = Phi(z) + z * phi(z)  (erf form, matching aten::gelu_backward "none")
// ---------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(TR_THREADS)
gelu_dropout_bwd_kernel(const T* __restrict__ dd, const T* __restrict__ z, const uint8_t* __restrict__ keep,
                        T* __restrict__ dz, long long n_vec, float scale) {
  constexpr int V = VecTraits<T>::N;                  // 8
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < n_vec; i0 += 2 * stride) {
    Vec16 vd[2], vz[2];
    uint2 vm[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const long long i = i0 + u * stride;
      if (i < n_vec) {
        vd[u] = ld_vec(dd + i * V);
        vz[u] = ld_vec(z + i * V);
        if (keep) vm[u] = *reinterpret_cast<const uint2*>(keep + i * V);
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const long long i = i0 + u * stride;
      if (i < n_vec) {
        float fd[V], fz[V], out[V];
        unpack<T>(vd[u], fd);
        unpack<T>(vz[u], fz);
#pragma unroll
        for (int e = 0; e < V; ++e) {
          float k = scale;
          if (keep) {
            const uint32_t word = e < 4 ? vm[u].x : vm[u].y;
            k = ((word >> (8 * (e & 3))) & 0xffu) ? scale : 0.f;
          }
          const float x = fz[e];
          const float cdf = 0.5f * (1.f + erff(x * 0.70710678118654752f));
          const float pdf = 0.39894228040143268f * __expf(-0.5f * x * x);
          out[e] = fd[e] * k * fmaf(x, pdf, cdf);
        }
        st_vec(dz + i * V, pack<T>(out));
      }
    }
  }
}

int grid_for(long long items, int per_thread) {
  long long need = (items + (long long)TR_THREADS * per_thread - 1) / ((long long)TR_THREADS * per_thread);
  if (need < 1) need = 1;
  if (need > 148 * 8) need = 148 * 8;
  return (int)need;
}

}  // namespace

extern "C" int adl_bind_thread();

extern "C" {

// 2-byte elements (bf16 / fp16): D % 8 == 0. split: src [A,B,W*H,D] -> dst [W][A,H,B,D]
// (one buffer); merge: W separately allocated planes [A,H,B,D] (src0..src2; W <= 3) ->
// dst [A,B,W*H,D].
int adl_heads_permute(const void* src0, const void* src1, const void* src2, void* dst, int A, int B, int W,
                      int H, int D, int merge, void* stream) {
  if (int rc = adl_bind_thread()) return rc;
  if (D % 8 != 0 || A <= 0 || B <= 0 || W <= 0 || W > 3 || H <= 0) return -40;
  const int vecs = D / 8;
  const long long total = (long long)A * B * W * H * vecs;
  const int grid = grid_for(total, 4);
  cudaStream_t s = (cudaStream_t)stream;
  if (merge) {
    Planes planes;
    planes.p[0] = static_cast<const Vec16*>(src0);
    planes.p[1] = static_cast<const Vec16*>(src1);
    planes.p[2] = static_cast<const Vec16*>(src2);
    planes.p[3] = nullptr;
    heads_merge_kernel<<<grid, TR_THREADS, 0, s>>>(planes, static_cast<Vec16*>(dst), A, B, W, H, vecs);
  } else {
    heads_split_kernel<<<grid, TR_THREADS, 0, s>>>(static_cast<const Vec16*>(src0), static_cast<Vec16*>(dst), A, B, W, H, vecs);
  }
  return (int)cudaGetLastError();
}

// out[N] (fp32) = column sums of x[M, N]; dtype 0 fp32, 1 bf16, 2 fp16; N a multiple of the
// vector width; partial: chunks * ceil(N/64) * 64 floats; tickets: ceil(N/64) zeroed ints.
int adl_colsum(const void* x, float* out, float* partial, int* tickets, int M, int N, int dtype, int chunks,
               void* stream) {
  if (int rc = adl_bind_thread()) return rc;
  const int v = dtype == 0 ? 4 : 8;
  if (N % v != 0 || M <= 0 || chunks <= 0) return -41;
  const dim3 grid((N + CS_COLS - 1) / CS_COLS, chunks);
  cudaStream_t s = (cudaStream_t)stream;
  switch (dtype) {
    case 0: colsum_kernel<float><<<grid, TR_THREADS, 0, s>>>(static_cast<const float*>(x), out, partial, tickets, M, N); break;
    case 1: colsum_kernel<__nv_bfloat16><<<grid, TR_THREADS, 0, s>>>(static_cast<const __nv_bfloat16*>(x), out, partial, tickets, M, N); break;
    case 2: colsum_kernel<__half><<<grid, TR_THREADS, 0, s>>>(static_cast<const __half*>(x), out, partial, tickets, M, N); break;
    default: return -42;
  }
  return (int)cudaGetLastError();
}

// dz = dd * keep * scale * gelu'(z); 16-bit tensors of n elements (n % 8 == 0), keep: bytes (1 = kept)
// or nullptr. dtype 1 bf16, 2 fp16.
int adl_gelu_dropout_bwd(const void* dd, const void* z, const void* keep, void* dz, long long n, float scale,
                         int dtype, void* stream) {
  if (int rc = adl_bind_thread()) return rc;
  if (n <= 0 || n % 8 != 0) return -44;
  const long long n_vec = n / 8;
  const int grid = grid_for(n_vec, 2);
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == 1)
    gelu_dropout_bwd_kernel<__nv_bfloat16><<<grid, TR_THREADS, 0, s>>>(
        static_cast<const __nv_bfloat16*>(dd), static_cast<const __nv_bfloat16*>(z),
        static_cast<const uint8_t*>(keep), static_cast<__nv_bfloat16*>(dz), n_vec, scale);
  else if (dtype == 2)
    gelu_dropout_bwd_kernel<__half><<<grid, TR_THREADS, 0, s>>>(
        static_cast<const __half*>(dd), static_cast<const __half*>(z), static_cast<const uint8_t*>(keep),
        static_cast<__half*>(dz), n_vec, scale);
  else
    return -45;
  return (int)cudaGetLastError();
}

// to_bf16 == 0: dst fp32 [M, n] <- src bf16 [M, ld];  to_bf16 == 1: dst bf16 [M, ld] <- src fp32 [M, n]
int adl_slice_cast(const void* src, void* dst, int M, int n, int ld, int to_bf16, void* stream) {
  if (int rc = adl_bind_thread()) return rc;
  if (ld % 8 != 0 || n % 4 != 0 || n > ld || M <= 0) return -43;
  cudaStream_t s = (cudaStream_t)stream;
  if (to_bf16) {
    const int grid = grid_for((long long)M * (ld / 8), 2);
    f32_to_slice_kernel<<<grid, TR_THREADS, 0, s>>>(static_cast<const float*>(src), static_cast<__nv_bfloat16*>(dst), M, n, ld);
  } else {
    const int grid = grid_for((long long)M * ((n + 7) / 8), 2);
    slice_to_f32_kernel<<<grid, TR_THREADS, 0, s>>>(static_cast<const __nv_bfloat16*>(src), static_cast<float*>(dst), M, n, ld);
  }
  return (int)cudaGetLastError();
}

}  // extern "C"
