// C-ABI glue: error string, version, and the sc_linear dispatcher (tensor-core bf16 path vs fp32 verification path).
#include "sc_common.cuh"
#include <cstring>

static thread_local char g_err[512] = "";

void sc_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int sc_gemm_bf16_launch(const void* x, const void* w, int w_dtype, const float* mask, int mask_mode, const float* uniforms,
                        unsigned long long seed, unsigned long long stream_id, const float* bias, const float* residual,
                        void* y, int y_dtype, int M, int N, int K, int relu, int block_n, cudaStream_t stream);
int sc_gemm_f32_launch(const float* x, const float* w, const float* mask, int mask_mode, const float* uniforms,
                       unsigned long long seed, unsigned long long stream_id, const float* bias, const float* residual,
                       void* y, int y_dtype, int M, int N, int K, int relu, cudaStream_t stream);

extern "C" {

const char* sc_last_error(void) { return g_err; }
int sc_version(void) { return 100; }  // round 1

int sc_linear(const void* x, int x_dtype, const void* w, int w_dtype, const float* mask, int mask_mode,
              const float* uniforms, unsigned long long seed, unsigned long long stream_id, const float* bias,
              const float* residual, void* y, int y_dtype, int M, int N, int K, int relu, int tile_n,
              cudaStream_t stream) {
  SC_CHECK(mask_mode >= SC_MASK_NONE && mask_mode <= SC_MASK_UNIFORM, SC_ERR_UNSUPPORTED, "sc_linear: mask_mode %d", mask_mode);
  if (x_dtype == SC_BF16)
    return sc_gemm_bf16_launch(x, w, w_dtype, mask, mask_mode, uniforms, seed, stream_id, bias, residual, y, y_dtype, M, N, K,
                               relu, tile_n, stream);
  SC_CHECK(x_dtype == SC_F32, SC_ERR_DTYPE, "sc_linear: bad x dtype %d", x_dtype);
  SC_CHECK(w_dtype == SC_F32, SC_ERR_DTYPE, "sc_linear: fp32 activations need fp32 weights");
  return sc_gemm_f32_launch((const float*)x, (const float*)w, mask, mask_mode, uniforms, seed, stream_id, bias, residual, y,
                            y_dtype, M, N, K, relu, stream);
}

}  // extern "C"
