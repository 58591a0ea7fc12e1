// K3b: Y = X * W^T + bias for 80-99 % unstructured-sparse W stored as CSR over output rows
// (2-byte column index, the layout SURVEY.md section 2.3 K3b names; the reference only *stores* COO:
// sparse_caption/pruning/prune.py:200-221, utils/model_utils.py:110-118).
//
// CTA = (tile of M_T = 32*MPL activation rows) x (chunk of output columns).  The X tile is staged in shared
// memory TRANSPOSED ([K][M_T]) so that "lane = row" reads of X[:, col] are bank-conflict free; each warp walks
// whole CSR rows: the (col,val) pairs are fetched 32 at a time (coalesced) and broadcast with shuffles; every
// lane keeps MPL fp32 accumulators.  One CSR row's result is written by one warp: deterministic, no atomics.
#include "sc_common.cuh"

namespace {

struct SpmmArgs {
  const void* x;            // [M,K]
  const int* row_ptr;       // [N+1]
  const unsigned short* col;  // [nnz]
  const void* val;          // [nnz] (same dtype as x)
  const float* bias; const float* residual; void* y; int y_bf16; int relu;
  int M, N, K, n_per_cta;
};

template <typename T, int MPL>
__global__ void __launch_bounds__(256) csr_spmm_kernel(const SpmmArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* xs = (T*)smem_raw;  // [K][M_T]
  constexpr int M_T = 32 * MPL;
  const int m0 = blockIdx.x * M_T;
  const int nb = blockIdx.y * a.n_per_cta;
  const int ne = min(a.N, nb + a.n_per_cta);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const T* x = (const T*)a.x;
  // stage X^T: consecutive threads read consecutive k of one row (coalesced), write column m of xs
  for (int e = tid; e < M_T * a.K; e += 256) {
    const int m = e / a.K, k = e - m * a.K;
    xs[(size_t)k * M_T + m] = (m0 + m < a.M) ? x[(size_t)(m0 + m) * a.K + k] : sc::from_f32<T>(0.f);
  }
  __syncthreads();
  const T* vals = (const T*)a.val;
  for (int n = nb + warp; n < ne; n += 8) {
    float acc[MPL];
#pragma unroll
    for (int i = 0; i < MPL; ++i) acc[i] = 0.f;
    const int pe = a.row_ptr[n + 1];
    for (int p0 = a.row_ptr[n]; p0 < pe; p0 += 32) {
      const int p = p0 + lane;
      int c = 0; float v = 0.f;
      if (p < pe) { c = a.col[p]; v = sc::to_f32<T>(vals[p]); }
      const int cnt = min(32, pe - p0);
      for (int l = 0; l < cnt; ++l) {
        const int cc = __shfl_sync(0xffffffffu, c, l);
        const float vv = __shfl_sync(0xffffffffu, v, l);
        const T* xr = xs + (size_t)cc * M_T + lane * MPL;
#pragma unroll
        for (int i = 0; i < MPL; ++i) acc[i] = fmaf(vv, sc::to_f32<T>(xr[i]), acc[i]);
      }
    }
    const float bz = a.bias ? a.bias[n] : 0.f;
#pragma unroll
    for (int i = 0; i < MPL; ++i) {
      const int m = m0 + lane * MPL + i;
      if (m >= a.M) continue;
      float r = acc[i] + bz;
      if (a.relu) r = fmaxf(r, 0.f);
      if (a.residual) r += a.residual[(size_t)m * a.N + n];
      if (a.y_bf16) ((__nv_bfloat16*)a.y)[(size_t)m * a.N + n] = __float2bfloat16_rn(r);
      else ((float*)a.y)[(size_t)m * a.N + n] = r;
    }
  }
}

// K larger than one shared-memory tile (fp32 verification mode with K = dim_feedforward): the CTA walks K in
// chunks of KC columns; each warp keeps a cursor and fp32 accumulators for its <= 8 CSR rows across chunks
// (columns are sorted inside a CSR row, so the entries of one chunk form a prefix of what is left).
template <typename T, int MPL>
__global__ void __launch_bounds__(256) csr_spmm_chunked_kernel(const SpmmArgs a, int KC) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* xs = (T*)smem_raw;  // [KC][M_T]
  constexpr int M_T = 32 * MPL;
  constexpr int RPW = 8;  // CSR rows per warp (n_per_cta == 64)
  const int m0 = blockIdx.x * M_T;
  const int nb = blockIdx.y * 64;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const T* x = (const T*)a.x;
  const T* vals = (const T*)a.val;
  float acc[RPW][MPL];
  int cur[RPW], end[RPW];
#pragma unroll
  for (int i = 0; i < RPW; ++i) {
    const int n = nb + warp + 8 * i;
    cur[i] = n < a.N ? a.row_ptr[n] : 0;
    end[i] = n < a.N ? a.row_ptr[n + 1] : 0;
#pragma unroll
    for (int j = 0; j < MPL; ++j) acc[i][j] = 0.f;
  }
  for (int k0 = 0; k0 < a.K; k0 += KC) {
    const int kw = min(KC, a.K - k0);
    __syncthreads();
    for (int e = tid; e < M_T * kw; e += 256) {
      const int m = e / kw, k = e - m * kw;
      xs[(size_t)k * M_T + m] = (m0 + m < a.M) ? x[(size_t)(m0 + m) * a.K + k0 + k] : sc::from_f32<T>(0.f);
    }
    __syncthreads();
    const int kend = k0 + kw;
#pragma unroll
    for (int i = 0; i < RPW; ++i) {
      while (cur[i] < end[i]) {
        const int p = cur[i] + lane;
        int c = 0x7fffffff; float v = 0.f;
        if (p < end[i]) { c = a.col[p]; v = sc::to_f32<T>(vals[p]); }
        const int cnt = __popc(__ballot_sync(0xffffffffu, c < kend));
        for (int l = 0; l < cnt; ++l) {
          const int cc = __shfl_sync(0xffffffffu, c, l) - k0;
          const float vv = __shfl_sync(0xffffffffu, v, l);
          const T* xr = xs + (size_t)cc * M_T + lane * MPL;
#pragma unroll
          for (int j = 0; j < MPL; ++j) acc[i][j] = fmaf(vv, sc::to_f32<T>(xr[j]), acc[i][j]);
        }
        cur[i] += cnt;
        if (cnt < 32) break;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < RPW; ++i) {
    const int n = nb + warp + 8 * i;
    if (n >= a.N) continue;
    const float bz = a.bias ? a.bias[n] : 0.f;
#pragma unroll
    for (int j = 0; j < MPL; ++j) {
      const int m = m0 + lane * MPL + j;
      if (m >= a.M) continue;
      float r = acc[i][j] + bz;
      if (a.relu) r = fmaxf(r, 0.f);
      if (a.residual) r += a.residual[(size_t)m * a.N + n];
      if (a.y_bf16) ((__nv_bfloat16*)a.y)[(size_t)m * a.N + n] = __float2bfloat16_rn(r);
      else ((float*)a.y)[(size_t)m * a.N + n] = r;
    }
  }
}

template <typename T, int MPL>
int launch_chunked(const SpmmArgs& a, cudaStream_t stream) {
  constexpr int M_T = 32 * MPL;
  const int KC = (int)((160 * 1024) / (M_T * sizeof(T)));
  const size_t smem = (size_t)KC * M_T * sizeof(T);
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(csr_spmm_chunked_kernel<T, MPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    SC_CHECK(e == cudaSuccess, (int)e, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr = true;
  }
  dim3 grid((a.M + M_T - 1) / M_T, (a.N + 63) / 64);
  csr_spmm_chunked_kernel<T, MPL><<<grid, 256, smem, stream>>>(a, KC);
  SC_LAUNCH_CHECK("sc_csr_spmm(chunked)");
  return SC_OK;
}

template <typename T, int MPL>
int launch(const SpmmArgs& a0, cudaStream_t stream) {
  SpmmArgs a = a0;
  constexpr int M_T = 32 * MPL;
  const size_t smem = (size_t)a.K * M_T * sizeof(T);
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(csr_spmm_kernel<T, MPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    SC_CHECK(e == cudaSuccess, (int)e, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr = true;
  }
  const int m_tiles = (a.M + M_T - 1) / M_T;
  // enough column chunks to give ~2 CTAs per SM, but at least 8 CSR rows per warp pass
  int chunks = (2 * 148 + m_tiles - 1) / m_tiles;
  if (chunks < 1) chunks = 1;
  int npc = (a.N + chunks - 1) / chunks;
  npc = ((npc + 7) / 8) * 8;
  if (npc < 8) npc = 8;
  a.n_per_cta = npc;
  dim3 grid(m_tiles, (a.N + npc - 1) / npc);
  csr_spmm_kernel<T, MPL><<<grid, 256, smem, stream>>>(a);
  SC_LAUNCH_CHECK("sc_csr_spmm");
  return SC_OK;
}

}  // namespace

extern "C" int sc_csr_spmm(const void* x, int dtype, const int* row_ptr, const unsigned short* col_idx, const void* vals,
                           const float* bias, const float* residual, void* y, int y_dtype, int M, int N, int K, int relu,
                           cudaStream_t stream) {
  SC_CHECK(M > 0 && N > 0 && K > 0, SC_ERR_SHAPE, "sc_csr_spmm: M=%d N=%d K=%d", M, N, K);
  SC_CHECK(K <= 65536, SC_ERR_UNSUPPORTED, "sc_csr_spmm: K=%d does not fit 16-bit column indices", K);
  SC_CHECK(y_dtype == SC_F32 || y_dtype == SC_BF16, SC_ERR_DTYPE, "sc_csr_spmm: bad y dtype");
  SpmmArgs a;
  a.x = x; a.row_ptr = row_ptr; a.col = col_idx; a.val = vals; a.bias = bias; a.residual = residual; a.y = y;
  a.y_bf16 = (y_dtype == SC_BF16); a.relu = relu; a.M = M; a.N = N; a.K = K; a.n_per_cta = 0;
  const size_t budget = 200 * 1024;
  if (dtype == SC_BF16) {
    if ((size_t)K * 128 * 2 <= budget && M > 64) return launch<__nv_bfloat16, 4>(a, stream);
    if ((size_t)K * 64 * 2 <= budget && M > 32) return launch<__nv_bfloat16, 2>(a, stream);
    if ((size_t)K * 32 * 2 <= budget) return launch<__nv_bfloat16, 1>(a, stream);
    return launch_chunked<__nv_bfloat16, 2>(a, stream);
  } else if (dtype == SC_F32) {
    if ((size_t)K * 64 * 4 <= budget && M > 32) return launch<float, 2>(a, stream);
    if ((size_t)K * 32 * 4 <= budget) return launch<float, 1>(a, stream);
    return launch_chunked<float, 2>(a, stream);
  }
  SC_CHECK(false, SC_ERR_DTYPE, "sc_csr_spmm: bad dtype %d", dtype);
  return SC_OK;
}
