// C-ABI glue: error string, version, and the sc_linear dispatcher (tensor-core bf16 path vs fp32 verification path).
#include "sc_common.cuh"
#include <cstring>

static thread_local char g_err[512] = "";
int g_sc_pdl = 7;  // programmatic dependent launch between consecutive kernels of a stream (sc_set_pdl): bit 0 GEMM /
                   // inference kernels, bit 1 training row / attention kernels, bit 2 early trigger inside the GEMM (after its loads are issued)

void sc_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int sc_gemm_bf16_launch(const void* x, const void* w, int w_dtype, const float* mask, int mask_mode, const float* uniforms,
                        unsigned long long seed, unsigned long long stream_id, const float* bias, const float* residual,
                        void* y, int y_dtype, int M, int N, int K, int relu, int block_n, const ScGemmExtra* ex,
                        cudaStream_t stream);
int sc_gemm_f32_launch(const float* x, const float* w, const float* mask, int mask_mode, const float* uniforms,
                       unsigned long long seed, unsigned long long stream_id, const float* bias, const float* residual,
                       void* y, int y_dtype, int M, int N, int K, int relu, const ScGemmExtra* ex, cudaStream_t stream);

int sc_gemm_wgrad_splits(int N, int K, int M, int max_splits);
extern "C" int sc_mask_grad_reduce_launch(const float* part, int splits, size_t stride, const float* w, const float* mask,
                                          int mask_mode, const float* uniforms, unsigned long long seed,
                                          unsigned long long stream_id, int bypass, float sp_coeff, float* dw, float* ds,
                                          int accumulate, size_t n, cudaStream_t stream);

static int linear_dispatch(const void* x, int x_dtype, const void* w, int w_dtype, const float* mask, int mask_mode,
                           const float* uniforms, unsigned long long seed, unsigned long long stream_id, const float* bias,
                           const float* residual, void* y, int y_dtype, int M, int N, int K, int relu, int tile_n,
                           const ScGemmExtra* ex, cudaStream_t stream) {
  SC_CHECK(mask_mode >= SC_MASK_NONE && mask_mode <= SC_MASK_UNIFORM, SC_ERR_UNSUPPORTED, "sc_linear: mask_mode %d", mask_mode);
  if (x_dtype == SC_BF16)
    return sc_gemm_bf16_launch(x, w, w_dtype, mask, mask_mode, uniforms, seed, stream_id, bias, residual, y, y_dtype, M, N, K,
                               relu, tile_n, ex, stream);
  SC_CHECK(x_dtype == SC_F32, SC_ERR_DTYPE, "sc_linear: bad x dtype %d", x_dtype);
  SC_CHECK(w_dtype == SC_F32, SC_ERR_DTYPE, "sc_linear: fp32 activations need fp32 weights");
  return sc_gemm_f32_launch((const float*)x, (const float*)w, mask, mask_mode, uniforms, seed, stream_id, bias, residual, y,
                            y_dtype, M, N, K, relu, ex, stream);
}

extern "C" {

const char* sc_last_error(void) { return g_err; }
int sc_version(void) { return 101; }
int sc_set_pdl(int enabled) { g_sc_pdl = enabled == 1 ? 7 : (enabled & 7); return SC_OK; }

int sc_linear(const void* x, int x_dtype, const void* w, int w_dtype, const float* mask, int mask_mode,
              const float* uniforms, unsigned long long seed, unsigned long long stream_id, const float* bias,
              const float* residual, void* y, int y_dtype, int M, int N, int K, int relu, int tile_n,
              cudaStream_t stream) {
  return linear_dispatch(x, x_dtype, w, w_dtype, mask, mask_mode, uniforms, seed, stream_id, bias, residual, y, y_dtype, M, N,
                         K, relu, tile_n, nullptr, stream);
}

// Inference GEMM with a LayerNorm folded around it (models/transformer.py:329-358: x + sublayer(norm(x))).
//   consumer (ln_stats != NULL): w holds W (.) a, ln_c[n] = sum_k w[n,k], bias = W b + bias; x is the bf16 copy of the
//     un-normalised residual stream; y = rstd * (x w^T) - rstd * mean * ln_c + bias with the row statistics of ln_stats.
//   producer (stats_out / y_bf16_copy != NULL): after bias / ReLU / residual the fp32 result is stored to y, its bf16
//     copy to y_bf16_copy and per 32-column chunk (sum, M2) to stats_out [M][N/32][2].
int sc_linear_ln(const void* x, const void* w, const float* bias, const float* residual, void* y, int y_dtype, int M, int N,
                 int K, int relu, int tile_n, const float* ln_stats, const float* ln_c, float ln_eps, void* y_bf16_copy,
                 float* stats_out, cudaStream_t stream) {
  ScGemmExtra ex = {};
  ex.ln_stats = ln_stats; ex.ln_c = ln_c; ex.ln_eps = ln_eps; ex.y2 = y_bf16_copy; ex.stats_out = stats_out;
  return sc_gemm_bf16_launch(x, w, SC_BF16, nullptr, SC_MASK_NONE, nullptr, 0, 0, bias, residual, y, y_dtype, M, N, K, relu,
                             tile_n, &ex, stream);
}

// Generator fused with the row pass of the beam step (models/transformer.py:405-413 OutputEmbedding +
// models/caption_model.py:56-111 beam_step): the [M, N] logits are never written.  For every row and every (256-column
// tile, epilogue-warp half) the GEMM epilogue leaves one record of 12 floats {max, sum exp(x - max), 5 largest logits,
// their column indices (int bits)} in partials [M][sc_linear_topk_parts(N)][12]; sc_beam_step_partials reduces them.
int sc_linear_topk_parts(int N) { return 2 * ((N + 255) / 256); }
int sc_linear_topk(const void* x, const void* w, const float* bias, int M, int N, int K, float* partials, int candidates,
                   cudaStream_t stream) {
  SC_CHECK(partials != nullptr && ((uintptr_t)partials & 3) == 0, SC_ERR_ALIGN, "sc_linear_topk: partials missing");
  SC_CHECK(candidates >= 1 && candidates <= 5, SC_ERR_UNSUPPORTED, "sc_linear_topk: candidates=%d not in [1,5]", candidates);
  ScGemmExtra ex = {};
  ex.topk_part = partials; ex.topk_n = candidates;
  return sc_gemm_bf16_launch(x, w, SC_BF16, nullptr, SC_MASK_NONE, nullptr, 0, 0, bias, nullptr, nullptr, SC_F32, M, N, K, 0, 0, &ex,
                             stream);
}

// Backward chain, dX GEMM of a linear whose INPUT was h = dropout(relu(.)) (feed-forward w_2): y [M,N] bf16 =
// (x w^T) * scale where h != 0, else 0 - i.e. the gradient with respect to the pre-activation of the previous linear, ready
// as its weight-gradient / dX operand - and colsum += column sums of y (that linear's bias gradient).  Replaces the fp32
// store of dX plus the separate sc_prep_grad pass.
int sc_linear_hmask(const void* x, const void* w, const void* h, float scale, float* colsum, void* y, int M, int N, int K,
                    cudaStream_t stream) {
  SC_CHECK(h != nullptr && ((uintptr_t)h & 15) == 0 && (colsum == nullptr || ((uintptr_t)colsum & 3) == 0), SC_ERR_ALIGN,
           "sc_linear_hmask: h must be 16-byte aligned");
  ScGemmExtra ex = {};
  ex.hmask = h; ex.hscale = scale; ex.colsum = colsum;
  return sc_gemm_bf16_launch(x, w, SC_BF16, nullptr, SC_MASK_NONE, nullptr, 0, 0, nullptr, nullptr, y, SC_BF16, M, N, K, 0, 0, &ex,
                             stream);
}

// training forward: y = dropout(act(x (W.m)^T + b), p) + residual, dropout mask = Philox(drop_seed, drop_stream, element)
int sc_linear_dropout(const void* x, int x_dtype, const void* w, int w_dtype, const float* mask, int mask_mode,
                      const float* uniforms, unsigned long long seed, unsigned long long stream_id, const float* bias,
                      const float* residual, void* y, int y_dtype, int M, int N, int K, int relu, int tile_n,
                      float dropout_p, unsigned long long drop_seed, unsigned long long drop_stream, cudaStream_t stream) {
  SC_CHECK(dropout_p >= 0.f && dropout_p < 1.f, SC_ERR_SHAPE, "sc_linear_dropout: p=%f", dropout_p);
  ScGemmExtra ex = {};
  ex.dropout_p = dropout_p; ex.drop_seed = drop_seed; ex.drop_stream = drop_stream;
  return linear_dispatch(x, x_dtype, w, w_dtype, mask, mask_mode, uniforms, seed, stream_id, bias, residual, y, y_dtype, M, N,
                         K, relu, tile_n, &ex, stream);
}

// K2 weight gradient: dWm[N,K] = dyT[N,M] * xT[K,M]^T with the straight-through epilogue
//   dW (+)= dWm (.) m ;  dS (+)= dWm (.) W (.) sigmoid'(S) [(.) 1 when bypass / raw] + sparsity_coeff * sigmoid'(S)
// dyT, xT: transposed activations in `dtype`; w, mask: the fp32 weight and its logits (mask regenerated from
// (mask_mode, seed, stream_id, element) exactly as the forward drew it); dw / ds may be NULL.
// workspace != NULL (bf16 operands): two kernels - the GEMM stores split-K partial products dWm_s (fp32, plain coalesced
// stores) into the workspace, sc_mask_grad_reduce sums them and applies the straight-through epilogue once per element at
// full-GPU parallelism.  workspace == NULL: one kernel with the epilogue fused into the GEMM (split-K through atomics).
int sc_linear_wgrad(const void* dyT, const void* xT, int dtype, const float* w, const float* mask, int mask_mode,
                    const float* uniforms, unsigned long long seed, unsigned long long stream_id, int bypass_sigmoid_grad,
                    float sparsity_coeff, float* dw, float* ds, int accumulate, int N, int K, int M, int tile_n,
                    void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  SC_CHECK(w != nullptr && (dw != nullptr || ds != nullptr), SC_ERR_SHAPE, "sc_linear_wgrad: w and one of dw/ds are required");
  SC_CHECK(mask_mode == SC_MASK_NONE || mask != nullptr, SC_ERR_SHAPE, "sc_linear_wgrad: mask missing");
  const size_t nk = (size_t)N * K;
  if (workspace != nullptr && dtype == SC_BF16 && nk % 4 == 0 && workspace_bytes >= nk * sizeof(float)) {
    SC_CHECK(((uintptr_t)workspace & 15) == 0, SC_ERR_ALIGN, "sc_linear_wgrad: workspace must be 16-byte aligned");
    const int max_splits = (int)(workspace_bytes / (nk * sizeof(float)) > 64 ? 64 : workspace_bytes / (nk * sizeof(float)));
    const int splits = sc_gemm_wgrad_splits(N, K, M, max_splits);
    ScGemmExtra ex = {};
    ex.partial_splits = splits; ex.split_stride = nk;
    int rc = linear_dispatch(dyT, dtype, xT, dtype, nullptr, SC_MASK_NONE, nullptr, 0, 0, nullptr, nullptr, workspace, SC_F32, N, K,
                             M, 0, tile_n, &ex, stream);
    if (rc) return rc;
    return sc_mask_grad_reduce_launch((const float*)workspace, splits, nk, w, mask, mask_mode, uniforms, seed, stream_id,
                                      bypass_sigmoid_grad, sparsity_coeff, dw, ds, accumulate, nk, stream);
  }
  ScGemmExtra ex = {};
  ex.wgrad = 1; ex.bypass = bypass_sigmoid_grad; ex.sp_coeff = sparsity_coeff; ex.accumulate = accumulate;
  ex.wg_w = w; ex.wg_s = mask; ex.wg_u = uniforms; ex.dw = dw; ex.ds = ds;
  // GEMM view: rows = N (weight rows), cols = K (weight cols), contraction = M tokens; y unused (dw stands in for alignment checks)
  void* ydummy = dw ? (void*)dw : (void*)ds;
  return linear_dispatch(dyT, dtype, xT, dtype, nullptr, mask_mode, nullptr, seed, stream_id, nullptr, nullptr, ydummy, SC_F32, N,
                         K, M, 0, tile_n, &ex, stream);
}

// K2 without transposed copies: dy [M, N] and x [M, K] are the row-major bf16 activations as the forward / the
// gradient preparation left them; the GEMM reads them as MN-major UMMA tiles (TMA boxes of 64 tokens x 64 features),
// token counts need no padding (out-of-range rows are zero-filled by TMA).  Always the two-kernel form (workspace).
int sc_linear_wgrad_rowmajor(const void* dy, const void* x, const float* w, const float* mask, int mask_mode,
                             const float* uniforms, unsigned long long seed, unsigned long long stream_id,
                             int bypass_sigmoid_grad, float sparsity_coeff, float* dw, float* ds, int accumulate, int N, int K,
                             int M, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  SC_CHECK(w != nullptr && (dw != nullptr || ds != nullptr), SC_ERR_SHAPE, "sc_linear_wgrad_rowmajor: w and one of dw/ds are required");
  SC_CHECK(mask_mode == SC_MASK_NONE || mask != nullptr, SC_ERR_SHAPE, "sc_linear_wgrad_rowmajor: mask missing");
  const size_t nk = (size_t)N * K;
  SC_CHECK(workspace != nullptr && ((uintptr_t)workspace & 15) == 0 && workspace_bytes >= nk * sizeof(float) && nk % 4 == 0,
           SC_ERR_WORKSPACE, "sc_linear_wgrad_rowmajor: 16-byte aligned workspace of >= %zu bytes needed", nk * sizeof(float));
  const int max_splits = (int)(workspace_bytes / (nk * sizeof(float)) > 64 ? 64 : workspace_bytes / (nk * sizeof(float)));
  const int splits = sc_gemm_wgrad_splits(N, K, M, max_splits);
  ScGemmExtra ex = {};
  ex.partial_splits = splits; ex.split_stride = nk; ex.mn_major = 1;
  int rc = sc_gemm_bf16_launch(dy, x, SC_BF16, nullptr, SC_MASK_NONE, nullptr, 0, 0, nullptr, nullptr, workspace, SC_F32, N, K, M, 0,
                               0, &ex, stream);
  if (rc) return rc;
  return sc_mask_grad_reduce_launch((const float*)workspace, splits, nk, w, mask, mask_mode, uniforms, seed, stream_id,
                                    bypass_sigmoid_grad, sparsity_coeff, dw, ds, accumulate, nk, stream);
}

}  // extern "C"
