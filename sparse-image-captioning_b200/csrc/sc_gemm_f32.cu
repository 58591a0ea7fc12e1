// fp32 verification path of K1:  Y = X * (W (.) mask)^T + bias  with fp32 FMA accumulation.
// north_star asks for 1e-5 relative parity in fp32 next to the bf16 tensor-core path; this kernel is that
// mode (same mask functions, same epilogue), used by the parity tests and by precision="fp32" models.
// Reference: sparse_caption/pruning/masked_layer.py:84-110,134-135.
#include "sc_common.cuh"

namespace {

constexpr int TM = 64, TN = 64, TK = 16;

struct Args {
  const float* x; const float* w; const float* mask; const float* uniforms;
  int mask_mode; unsigned long long seed, stream_id;
  const float* bias; const float* residual; void* y; int y_bf16; int relu;
  int M, N, K;
  float dropout_p; unsigned long long drop_seed, drop_stream;
  int wgrad, wg_mode, bypass; float sp_coeff; int accumulate;
  const float* wg_w; const float* wg_s; const float* wg_u; float* dw; float* ds;
};

__device__ __forceinline__ float masked_w(const Args& a, const sc::Philox& ph, size_t e) {
  float w = __ldg(a.w + e);
  switch (a.mask_mode) {
    case SC_MASK_NONE: return w;
    case SC_MASK_ROUND: return w * sc::mask_round(__ldg(a.mask + e));
    case SC_MASK_RAW: return w * __ldg(a.mask + e);
    case SC_MASK_UNIFORM: return (__ldg(a.uniforms + e) < sc::sigmoidf_(__ldg(a.mask + e))) ? w : 0.f;
    default: {  // SC_MASK_BERNOULLI: same Philox call layout as the tensor path (4 elements per call)
      uint4 r = ph(e >> 2, a.stream_id);
      uint32_t bits = (e & 3) == 0 ? r.x : (e & 3) == 1 ? r.y : (e & 3) == 2 ? r.z : r.w;
      return (sc::u24(bits) < sc::sigmoidf_(__ldg(a.mask + e))) ? w : 0.f;
    }
  }
}

__global__ void __launch_bounds__(256) sc_gemm_f32_kernel(const Args a) {
  __shared__ float sx[TK][TM + 4];
  __shared__ float sw[TK][TN + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
  const int tx = tid & 15, ty = tid >> 4;  // 16 x 16 threads, 4x4 outputs each
  const sc::Philox ph(a.seed);
  const sc::Philox dph(a.drop_seed);
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < a.K; k0 += TK) {
    // 64 rows x 16 k per operand = 1024 elements, 4 per thread; consecutive threads walk k (coalesced)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = tid + i * 256;
      const int r = idx >> 4, kk = idx & 15;
      const int k = k0 + kk;
      float xv = 0.f, wv = 0.f;
      if (k < a.K) {
        if (m0 + r < a.M) xv = __ldg(a.x + (size_t)(m0 + r) * a.K + k);
        if (n0 + r < a.N) wv = masked_w(a, ph, (size_t)(n0 + r) * a.K + k);
      }
      sx[kk][r] = xv;
      sw[kk][r] = wv;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < TK; ++kk) {
      float xr[4], wr[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) xr[i] = sx[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) wr[j] = sw[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(xr[i], wr[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int row = m0 + ty * 4 + i;
    if (row >= a.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int col = n0 + tx * 4 + j;
      if (col >= a.N) continue;
      float v = acc[i][j];
      const size_t e = (size_t)row * a.N + col;
      if (a.wgrad) {
        const float sv = a.wg_s ? __ldg(a.wg_s + e) : 0.f;
        // in wgrad mode a.mask_mode describes the mask of the weight being differentiated; operands are unmasked
        const float m = sc::mask_value(a.wg_mode, sv, a.wg_u ? __ldg(a.wg_u + e) : 0.f, ph, e, a.stream_id);
        float gw, gs;
        sc::mask_grad_elem(a.wg_mode, v, __ldg(a.wg_w + e), sv, m, a.bypass, a.sp_coeff, gw, gs);
        if (a.dw) a.dw[e] = (a.accumulate ? a.dw[e] : 0.f) + gw;
        if (a.ds) a.ds[e] = (a.accumulate ? a.ds[e] : 0.f) + gs;
        continue;
      }
      if (a.bias) v += __ldg(a.bias + col);
      if (a.relu) v = fmaxf(v, 0.f);
      if (a.dropout_p > 0.f) v *= sc::keep_scale(dph, e, a.drop_stream, a.dropout_p);
      if (a.residual) v += __ldg(a.residual + e);
      if (a.y_bf16) ((__nv_bfloat16*)a.y)[(size_t)row * a.N + col] = __float2bfloat16_rn(v);
      else ((float*)a.y)[(size_t)row * a.N + col] = v;
    }
  }
}

}  // namespace

int sc_gemm_f32_launch(const float* x, const float* w, const float* mask, int mask_mode, const float* uniforms,
                       unsigned long long seed, unsigned long long stream_id, const float* bias, const float* residual,
                       void* y, int y_dtype, int M, int N, int K, int relu, const ScGemmExtra* ex, cudaStream_t stream) {
  SC_CHECK(M > 0 && N > 0 && K > 0, SC_ERR_SHAPE, "sc_linear: empty problem M=%d N=%d K=%d", M, N, K);
  SC_CHECK((ex && ex->wgrad) || mask_mode == SC_MASK_NONE || mask != nullptr, SC_ERR_SHAPE, "sc_linear: mask_mode %d needs a mask", mask_mode);
  SC_CHECK((ex && ex->wgrad) || mask_mode != SC_MASK_UNIFORM || uniforms != nullptr, SC_ERR_SHAPE, "sc_linear: uniforms missing");
  const bool wgrad = ex && ex->wgrad;
  Args a{};
  a.x = x; a.w = w; a.mask = wgrad ? nullptr : mask; a.uniforms = wgrad ? nullptr : uniforms;
  a.mask_mode = wgrad ? SC_MASK_NONE : mask_mode; a.seed = seed; a.stream_id = stream_id;
  a.bias = bias; a.residual = residual; a.y = y; a.y_bf16 = (y_dtype == SC_BF16); a.relu = relu; a.M = M; a.N = N; a.K = K;
  if (ex) {
    a.dropout_p = ex->dropout_p; a.drop_seed = ex->drop_seed; a.drop_stream = ex->drop_stream;
    a.wgrad = ex->wgrad; a.wg_mode = mask_mode; a.bypass = ex->bypass; a.sp_coeff = ex->sp_coeff; a.accumulate = ex->accumulate;
    a.wg_w = ex->wg_w; a.wg_s = ex->wg_s; a.wg_u = ex->wg_u; a.dw = ex->dw; a.ds = ex->ds;
  }
  dim3 grid((N + TN - 1) / TN, (M + TM - 1) / TM);
  sc_gemm_f32_kernel<<<grid, 256, 0, stream>>>(a);
  SC_LAUNCH_CHECK("sc_gemm_f32_kernel");
  return SC_OK;
}
