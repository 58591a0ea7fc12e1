// Shared helpers for the sm_100a kernels behind the C ABI declared in include/sc_b200.h.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdint>
#include <cstdio>
#include <cstdarg>

#define SC_F32 0
#define SC_BF16 1

// status codes (include/sc_b200.h)
#define SC_OK 0
#define SC_ERR_SHAPE (-1)
#define SC_ERR_ALIGN (-2)
#define SC_ERR_DTYPE (-3)
#define SC_ERR_WORKSPACE (-4)
#define SC_ERR_UNSUPPORTED (-5)
#define SC_ERR_DRIVER (-6)

// thread-local last-error string (sc_last_error)
void sc_set_error(const char* fmt, ...);

#define SC_CHECK(cond, code, ...)        \
  do {                                   \
    if (!(cond)) {                       \
      sc_set_error(__VA_ARGS__);         \
      return (code);                     \
    }                                    \
  } while (0)

// After a launch: positive return = cudaError_t of the launch.
#define SC_LAUNCH_CHECK(name)                                             \
  do {                                                                    \
    cudaError_t e__ = cudaGetLastError();                                 \
    if (e__ != cudaSuccess) {                                             \
      sc_set_error("%s: launch failed: %s", name, cudaGetErrorString(e__)); \
      return (int)e__;                                                    \
    }                                                                     \
  } while (0)

// optional GEMM epilogue extensions (training: dropout / weight-gradient epilogue; inference: folded LayerNorm)
struct ScGemmExtra {
  float dropout_p; unsigned long long drop_seed, drop_stream;
  int wgrad, bypass; float sp_coeff; int accumulate;
  const float* wg_w; const float* wg_s; const float* wg_u; float* dw; float* ds;
  // consumer of a LayerNorm folded into the weights (W' = W (.) a, ln_c[n] = sum_k W'[n,k], bias' = W b + bias):
  //   y = rstd[row] * acc - rstd[row] * mean[row] * ln_c[col] + bias'[col], row statistics from ln_stats
  //   ln_stats: fp32 [M][K/32][2] = per 32-column chunk (sum, M2 about the chunk mean) written by the producer
  const float* ln_stats; const float* ln_c; float ln_eps;
  // producer of the residual stream: also emit a bf16 copy of y (the next GEMM's TMA operand) and the row statistics
  void* y2; float* stats_out;
  // split-K partial products (two-kernel weight gradient): split s stores its fp32 tile at y + s * split_stride
  int partial_splits; size_t split_stride;
  // operands given as x [K, M] and w [K, N] row-major (y = x^T w): MN-major UMMA tiles, no transposed copies
  int mn_major;
  // generator fused with the beam step's row pass (kEpi == 3): [M][2 * ceil(N / 256)][12] records, no output tile
  float* topk_part; int topk_n;  // candidates actually needed per record (<= 3: cheaper epilogue)
  // dX GEMM preparing the next linear's gradient operand: y (bf16) = acc * hscale where hmask (bf16 [M,N]) != 0 else 0,
  // colsum (fp32 [N], accumulated) += column sums of y
  const void* hmask; float hscale; float* colsum;
};

// Programmatic dependent launch (griddepcontrol): kernels launched through sc::launch_pdl may start while their
// predecessor in the stream drains; they must call sc::pdl_wait() before touching global memory.
extern int g_sc_pdl;

namespace sc {

__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = (g_sc_pdl & 1) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
// same for the training-side row / attention kernels (bit 1 of the sc_set_pdl mask)
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl_aux(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = (g_sc_pdl & 2) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// ---------------------------------------------------------------------------------------------
// Supermask mask functions (reference: sparse_caption/pruning/sampler.py:43-66,
// masked_layer.py:84-110).
// ---------------------------------------------------------------------------------------------
#define SC_MASK_NONE 0       // plain weight
#define SC_MASK_ROUND 1      // eval: rint(sigmoid(S))
#define SC_MASK_BERNOULLI 2  // train: Bernoulli(sigmoid(S)) from Philox4x32-10(seed, offset, element)
#define SC_MASK_RAW 3        // snip / mag_* / lottery_* / mask_freeze: S used as the mask itself
#define SC_MASK_UNIFORM 4    // train with caller-provided uniforms: (u < sigmoid(S))  (parity tests)

// torch.round(torch.sigmoid(S)) in fp32 equals (S > 1.5 * 2^-24): sigmoid(S) only exceeds 0.5 by
// one ulp once exp(-S) rounds to 1-2^-23, and round-half-even sends exactly 0.5 to 0.
// Checked bit-for-bit against torch in tests/golden/binarize.npz.
#define SC_BINARIZE_THRESHOLD 8.940696716308594e-08f
__device__ __forceinline__ float mask_round(float s) { return s > SC_BINARIZE_THRESHOLD ? 1.f : 0.f; }

__device__ __forceinline__ float sigmoidf_(float s) { return 1.f / (1.f + __expf(-s)); }

// Philox4x32-10 (Salmon et al. 2011).  One call yields 4 x 32 random bits.
struct Philox {
  uint32_t k0, k1;
  // seed: an immediate value, or - bit 63 set - a DEVICE POINTER (low 63 bits) to the 64-bit seed.  The pointer form
  // lets a captured CUDA graph draw fresh masks / dropout on every replay: the host rewrites the seed word per step.
  __device__ __forceinline__ Philox(uint64_t seed) {
    if (seed >> 63) seed = *(const unsigned long long*)(seed & 0x7fffffffffffffffull);
    k0 = (uint32_t)seed; k1 = (uint32_t)(seed >> 32);
  }
  __device__ __forceinline__ uint4 operator()(uint64_t ctr, uint64_t stream) const {
    uint32_t c0 = (uint32_t)ctr, c1 = (uint32_t)(ctr >> 32), c2 = (uint32_t)stream, c3 = (uint32_t)(stream >> 32);
    uint32_t a = k0, b = k1;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
      uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
      uint32_t n0 = h1 ^ c1 ^ a, n1 = l1, n2 = h0 ^ c3 ^ b, n3 = l0;
      c0 = n0; c1 = n1; c2 = n2; c3 = n3;
      a += 0x9E3779B9u; b += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
  }
};
__device__ __forceinline__ float u24(uint32_t x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }

// Bernoulli masks for 4 consecutive elements whose linear index starts at e4*4.
__device__ __forceinline__ void bernoulli4(const Philox& ph, uint64_t e4, uint64_t stream, const float s[4], float m[4]) {
  uint4 r = ph(e4, stream);
  m[0] = u24(r.x) < sigmoidf_(s[0]) ? 1.f : 0.f;
  m[1] = u24(r.y) < sigmoidf_(s[1]) ? 1.f : 0.f;
  m[2] = u24(r.z) < sigmoidf_(s[2]) ? 1.f : 0.f;
  m[3] = u24(r.w) < sigmoidf_(s[3]) ? 1.f : 0.f;
}

// mask value of element e under `mode` (shared by forward prologues and backward epilogues so that both see the
// same sample): s = logit / raw mask, u = injected uniform.
__device__ __forceinline__ float mask_value(int mode, float s, float u, const Philox& ph, size_t e, unsigned long long stream) {
  switch (mode) {
    case SC_MASK_NONE: return 1.f;
    case SC_MASK_ROUND: return mask_round(s);
    case SC_MASK_RAW: return s;
    case SC_MASK_UNIFORM: return u < sigmoidf_(s) ? 1.f : 0.f;
    default: {
      uint4 rr = ph(e >> 2, stream);
      const uint32_t bits = (e & 3) == 0 ? rr.x : (e & 3) == 1 ? rr.y : (e & 3) == 2 ? rr.z : rr.w;
      return u24(bits) < sigmoidf_(s) ? 1.f : 0.f;
    }
  }
}

// inverted-dropout factor of element e: 0 (dropped, probability p) or 1/(1-p)
__device__ __forceinline__ float keep_scale(const Philox& ph, size_t e, unsigned long long stream, float p) {
  if (p <= 0.f) return 1.f;
  uint4 rr = ph(e >> 2, stream);
  const uint32_t bits = (e & 3) == 0 ? rr.x : (e & 3) == 1 ? rr.y : (e & 3) == 2 ? rr.z : rr.w;
  return u24(bits) >= p ? 1.f / (1.f - p) : 0.f;
}

// straight-through gradient of one masked weight element (sampler.py:15-17,32-34; prune.py:249-258):
//   dW = g*m ; dS = g*W*sigmoid'(S) [* 1 when bypass / raw] + sp_coeff*sigmoid'(S)
__device__ __forceinline__ void mask_grad_elem(int mode, float g, float w, float s, float m, int bypass, float sp_coeff,
                                               float& dw, float& ds) {
  dw = g * m;
  float d = g * w;
  float dsig = 1.f;
  if (mode != SC_MASK_RAW && mode != SC_MASK_NONE) { const float sig = sigmoidf_(s); dsig = sig * (1.f - sig); }
  if (!(bypass || mode == SC_MASK_RAW)) d *= dsig;
  if (sp_coeff != 0.f) d += sp_coeff * dsig;
  ds = d;
}

}  // namespace sc
