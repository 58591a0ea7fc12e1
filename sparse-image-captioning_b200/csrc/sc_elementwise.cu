// K9 (row/elementwise fusions) and the mask utilities of K1/K10.
//   sc_layernorm        : reference LayerNorm, a*(x-mean)/(std_unbiased+eps)+b  (models/transformer.py:329-341)
//   sc_embed_pe         : (W (.) m)[ids]*sqrt(d) + pe[pos]                      (transformer.py:383-401, masked_layer.py:160-169)
//   sc_apply_mask       : W (.) mask(S) -> bf16/fp32 ("densify")                (pruning/prune.py:165-174)
//   sc_mask_count       : sum rint(sigmoid(S))                                  (prune.py:124-144, 249-252)
//   sc_cast_f32_bf16    : activation cast
#include "sc_common.cuh"

namespace {

// one warp per row; D <= 2048, D % 32 == 0 keeps everything in registers
template <typename OutT, int kMaxPerLane>
__global__ void __launch_bounds__(256) layernorm_kernel(const float* __restrict__ x, const float* __restrict__ a,
                                                        const float* __restrict__ b, OutT* __restrict__ y, int rows, int D,
                                                        float eps) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  sc::pdl_wait();
  if (warp >= rows) return;
  const float* xr = x + (size_t)warp * D;
  float v[kMaxPerLane];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < kMaxPerLane; ++i) {
    const int c = lane + i * 32;
    v[i] = c < D ? xr[c] : 0.f;
    s += v[i];
  }
  sc::pdl_launch();  // after the row is loaded
  const float mean = sc::warp_sum(s) / (float)D;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < kMaxPerLane; ++i) {
    const int c = lane + i * 32;
    const float d = c < D ? v[i] - mean : 0.f;
    q += d * d;
  }
  const float var = sc::warp_sum(q) / (float)(D - 1);  // unbiased, torch.std default
  const float inv = 1.f / (sqrtf(var) + eps);
  OutT* yr = y + (size_t)warp * D;
#pragma unroll
  for (int i = 0; i < kMaxPerLane; ++i) {
    const int c = lane + i * 32;
    if (c < D) yr[c] = sc::from_f32<OutT>(__ldg(a + c) * (v[i] - mean) * inv + __ldg(b + c));
  }
}

template <typename OutT>
__global__ void __launch_bounds__(128) embed_pe_kernel(const int* __restrict__ tokens, const float* __restrict__ table,
                                                       const float* __restrict__ mask, int mask_mode,
                                                       const float* __restrict__ uniforms, unsigned long long seed,
                                                       unsigned long long stream_id, const float* __restrict__ pe,
                                                       OutT* __restrict__ out, int rows, int D, int V, int T, int pos0,
                                                       float scale) {
  const int r = blockIdx.x;
  sc::pdl_launch();
  sc::pdl_wait();
  if (r >= rows) return;
  int tok = tokens[r];
  tok = tok < 0 ? 0 : (tok >= V ? V - 1 : tok);
  const int pos = pos0 + (r % T);
  const sc::Philox ph(seed);
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    const size_t e = (size_t)tok * D + c;
    float w = __ldg(table + e);
    if (mask_mode == SC_MASK_ROUND) w *= sc::mask_round(__ldg(mask + e));
    else if (mask_mode == SC_MASK_RAW) w *= __ldg(mask + e);
    else if (mask_mode == SC_MASK_UNIFORM) w = (__ldg(uniforms + e) < sc::sigmoidf_(__ldg(mask + e))) ? w : 0.f;
    else if (mask_mode == SC_MASK_BERNOULLI) {
      uint4 rr = ph(e >> 2, stream_id);
      uint32_t bits = (e & 3) == 0 ? rr.x : (e & 3) == 1 ? rr.y : (e & 3) == 2 ? rr.z : rr.w;
      w = (sc::u24(bits) < sc::sigmoidf_(__ldg(mask + e))) ? w : 0.f;
    }
    out[(size_t)r * D + c] = sc::from_f32<OutT>(w * scale + __ldg(pe + (size_t)pos * D + c));
  }
}

// Decode-step embedding for the LayerNorm-folded path: one warp per row writes the fp32 residual stream, its bf16
// copy (the next GEMM's TMA operand) and the per-32-column (sum, M2) statistics sc_linear_ln consumes.  D % 32 == 0.
template <int kMaxPerLane>
__global__ void __launch_bounds__(256) embed_pe_stats_kernel(const int* __restrict__ tokens, const float* __restrict__ table,
                                                             const float* __restrict__ pe, float* __restrict__ x32,
                                                             __nv_bfloat16* __restrict__ xb, float* __restrict__ stats,
                                                             int rows, int D, int V, int T, int pos0, float scale) {
  const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  sc::pdl_launch();
  sc::pdl_wait();
  if (r >= rows) return;
  int tok = tokens[r];
  tok = tok < 0 ? 0 : (tok >= V ? V - 1 : tok);
  const int pos = pos0 + (r % T);
  const int chunks = D >> 5;
#pragma unroll
  for (int i = 0; i < kMaxPerLane; ++i) {
    if (i >= chunks) break;
    const int c = lane + i * 32;
    const float v = __ldg(table + (size_t)tok * D + c) * scale + __ldg(pe + (size_t)pos * D + c);
    x32[(size_t)r * D + c] = v;
    xb[(size_t)r * D + c] = __float2bfloat16_rn(v);
    const float s = sc::warp_sum(v);
    const float dlt = v - s * (1.f / 32.f);
    const float m2 = sc::warp_sum(dlt * dlt);
    if (lane == 0) *(float2*)(stats + ((size_t)r * chunks + i) * 2) = make_float2(s, m2);
  }
}

template <typename OutT>
__global__ void __launch_bounds__(256) apply_mask_kernel(const float* __restrict__ w, const float* __restrict__ mask,
                                                         int mask_mode, const float* __restrict__ uniforms,
                                                         unsigned long long seed, unsigned long long stream_id,
                                                         OutT* __restrict__ out, size_t n) {
  const sc::Philox ph(seed);
  // 4 elements per thread per iteration (n4 groups); tail handled by the last group guard
  const size_t n4 = (n + 3) / 4;
  for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < n4; g += (size_t)gridDim.x * blockDim.x) {
    const size_t e0 = g * 4;
    uint4 rr = make_uint4(0, 0, 0, 0);
    if (mask_mode == SC_MASK_BERNOULLI) rr = ph(g, stream_id);
    const uint32_t bits[4] = {rr.x, rr.y, rr.z, rr.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const size_t e = e0 + i;
      if (e >= n) break;
      float v = w[e];
      if (mask_mode == SC_MASK_ROUND) v *= sc::mask_round(mask[e]);
      else if (mask_mode == SC_MASK_RAW) v *= mask[e];
      else if (mask_mode == SC_MASK_UNIFORM) v = (uniforms[e] < sc::sigmoidf_(mask[e])) ? v : 0.f;
      else if (mask_mode == SC_MASK_BERNOULLI) v = (sc::u24(bits[i]) < sc::sigmoidf_(mask[e])) ? v : 0.f;
      out[e] = sc::from_f32<OutT>(v);
    }
  }
}

__global__ void __launch_bounds__(256) mask_count_kernel(const float* __restrict__ s, size_t n, unsigned long long* out) {
  unsigned int c = 0;
  // 16-byte loads, four in flight per thread (the scalar loop ran at a third of the HBM rate); s is 16-byte aligned
  const size_t n4 = ((uintptr_t)s & 15) == 0 ? n / 4 : 0;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < n4; g += 4 * stride) {
    float4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = g + u * stride < n4 ? __ldg((const float4*)s + g + u * stride) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int u = 0; u < 4; ++u)
      c += (v[u].x > SC_BINARIZE_THRESHOLD) + (v[u].y > SC_BINARIZE_THRESHOLD) + (v[u].z > SC_BINARIZE_THRESHOLD) +
           (v[u].w > SC_BINARIZE_THRESHOLD);
  }
  for (size_t i = n4 * 4 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    c += s[i] > SC_BINARIZE_THRESHOLD ? 1u : 0u;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  __shared__ unsigned int part[8];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int t = 0;
    for (int i = 0; i < 8; ++i) t += part[i];
    if (t) atomicAdd(out, (unsigned long long)t);
  }
}

// Fused ingest: reads fp32 features straight out of PINNED HOST memory (mapped into the device address space
// under UVA) over PCIe and stores the bf16 GEMM operand with a streaming hint - no fp32 staging buffer in HBM, no separate
// cast pass, half the L2 footprint of a cudaMemcpyAsync + cast.  Few CTAs, many 16-byte requests in flight per thread.
__global__ void __launch_bounds__(256) ingest_f32_bf16_kernel(const float* __restrict__ host, __nv_bfloat16* __restrict__ y, size_t n4) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < n4; g += 8 * stride) {
    float4 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u)
      if (g + u * stride < n4) v[u] = __ldcs((const float4*)host + g + u * stride);
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      if (g + u * stride < n4) {
        __nv_bfloat162 a = __floats2bfloat162_rn(v[u].x, v[u].y), b = __floats2bfloat162_rn(v[u].z, v[u].w);
        uint2 o; o.x = *(uint32_t*)&a; o.y = *(uint32_t*)&b;
        __stcs((uint2*)y + g + u * stride, o);
      }
    }
  }
}

__global__ void __launch_bounds__(256) cast_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, size_t n) {
  const size_t n4 = n / 4;
  for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < n4; g += (size_t)gridDim.x * blockDim.x) {
    float4 v = ((const float4*)x)[g];
    __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
    uint2 o; o.x = *(uint32_t*)&a; o.y = *(uint32_t*)&b;
    ((uint2*)y)[g] = o;
  }
  for (size_t i = n4 * 4 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    y[i] = __float2bfloat16_rn(x[i]);
}

__global__ void __launch_bounds__(256) mask_rows_kernel(float* __restrict__ x, const float* __restrict__ m, int rows, int D) {
  const size_t n = (size_t)rows * D;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    if (m[i / D] == 0.f) x[i] = 0.f;
}

int grid_for(size_t work_items, int block) {
  size_t g = (work_items + block - 1) / block;
  const size_t cap = 148 * 16;  // persistent-ish grid-stride: 16 CTAs per SM
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace

extern "C" {

int sc_layernorm(const float* x, const float* a, const float* b, void* y, int y_dtype, int rows, int D, float eps,
                 cudaStream_t stream) {
  SC_CHECK(rows > 0 && D > 1, SC_ERR_SHAPE, "sc_layernorm: rows=%d D=%d", rows, D);
  SC_CHECK(D <= 2048, SC_ERR_UNSUPPORTED, "sc_layernorm: D=%d > 2048", D);
  const int blocks = (rows + 7) / 8;
#define LN_LAUNCH(T, P) sc::launch_pdl(layernorm_kernel<T, P>, dim3(blocks), dim3(256), 0, stream, x, a, b, (T*)y, rows, D, eps)
  if (y_dtype == SC_F32) {
    if (D <= 128) LN_LAUNCH(float, 4); else if (D <= 512) LN_LAUNCH(float, 16); else LN_LAUNCH(float, 64);
  } else if (y_dtype == SC_BF16) {
    if (D <= 128) LN_LAUNCH(__nv_bfloat16, 4); else if (D <= 512) LN_LAUNCH(__nv_bfloat16, 16); else LN_LAUNCH(__nv_bfloat16, 64);
  } else {
    SC_CHECK(false, SC_ERR_DTYPE, "sc_layernorm: bad dtype %d", y_dtype);
  }
#undef LN_LAUNCH
  SC_LAUNCH_CHECK("sc_layernorm");
  return SC_OK;
}

int sc_embed_pe(const int* tokens, const float* table, const float* mask, int mask_mode, const float* uniforms,
                unsigned long long seed, unsigned long long stream_id, const float* pe, void* out, int out_dtype,
                int rows, int D, int V, int T, int pos0, float scale, cudaStream_t stream) {
  SC_CHECK(rows > 0 && D > 0 && V > 0 && T > 0, SC_ERR_SHAPE, "sc_embed_pe: rows=%d D=%d V=%d T=%d", rows, D, V, T);
  SC_CHECK(mask_mode == SC_MASK_NONE || mask != nullptr, SC_ERR_SHAPE, "sc_embed_pe: mask missing");
  if (out_dtype == SC_F32)
    sc::launch_pdl(embed_pe_kernel<float>, dim3(rows), dim3(128), 0, stream, tokens, table, mask, mask_mode, uniforms, seed,
                   stream_id, pe, (float*)out, rows, D, V, T, pos0, scale);
  else if (out_dtype == SC_BF16)
    sc::launch_pdl(embed_pe_kernel<__nv_bfloat16>, dim3(rows), dim3(128), 0, stream, tokens, table, mask, mask_mode, uniforms,
                   seed, stream_id, pe, (__nv_bfloat16*)out, rows, D, V, T, pos0, scale);
  else
    SC_CHECK(false, SC_ERR_DTYPE, "sc_embed_pe: bad dtype %d", out_dtype);
  SC_LAUNCH_CHECK("sc_embed_pe");
  return SC_OK;
}

int sc_embed_pe_stats(const int* tokens, const float* table, const float* pe, float* x32, void* x_bf16, float* stats, int rows,
                      int D, int V, int T, int pos0, float scale, cudaStream_t stream) {
  SC_CHECK(rows > 0 && D > 0 && V > 0 && T > 0, SC_ERR_SHAPE, "sc_embed_pe_stats: rows=%d D=%d V=%d T=%d", rows, D, V, T);
  SC_CHECK(D % 32 == 0 && D <= 2048, SC_ERR_UNSUPPORTED, "sc_embed_pe_stats: D=%d must be a multiple of 32, <= 2048", D);
  const int blocks = (rows + 7) / 8;
  if (D <= 512)
    sc::launch_pdl(embed_pe_stats_kernel<16>, dim3(blocks), dim3(256), 0, stream, tokens, table, pe, x32, (__nv_bfloat16*)x_bf16,
                   stats, rows, D, V, T, pos0, scale);
  else
    sc::launch_pdl(embed_pe_stats_kernel<64>, dim3(blocks), dim3(256), 0, stream, tokens, table, pe, x32, (__nv_bfloat16*)x_bf16,
                   stats, rows, D, V, T, pos0, scale);
  SC_LAUNCH_CHECK("sc_embed_pe_stats");
  return SC_OK;
}

int sc_apply_mask(const float* w, const float* mask, int mask_mode, const float* uniforms, unsigned long long seed,
                  unsigned long long stream_id, void* out, int out_dtype, size_t n, cudaStream_t stream) {
  SC_CHECK(n > 0, SC_ERR_SHAPE, "sc_apply_mask: n=0");
  SC_CHECK(mask_mode == SC_MASK_NONE || mask != nullptr, SC_ERR_SHAPE, "sc_apply_mask: mask missing");
  const int g = grid_for((n + 3) / 4, 256);
  if (out_dtype == SC_F32)
    apply_mask_kernel<float><<<g, 256, 0, stream>>>(w, mask, mask_mode, uniforms, seed, stream_id, (float*)out, n);
  else if (out_dtype == SC_BF16)
    apply_mask_kernel<__nv_bfloat16><<<g, 256, 0, stream>>>(w, mask, mask_mode, uniforms, seed, stream_id,
                                                            (__nv_bfloat16*)out, n);
  else
    SC_CHECK(false, SC_ERR_DTYPE, "sc_apply_mask: bad dtype %d", out_dtype);
  SC_LAUNCH_CHECK("sc_apply_mask");
  return SC_OK;
}

int sc_mask_count(const float* logits, size_t n, unsigned long long* count_out, cudaStream_t stream) {
  SC_CHECK(n > 0 && count_out != nullptr, SC_ERR_SHAPE, "sc_mask_count: bad args");
  mask_count_kernel<<<grid_for(n, 256), 256, 0, stream>>>(logits, n, count_out);
  SC_LAUNCH_CHECK("sc_mask_count");
  return SC_OK;
}

int sc_mask_rows(float* x, const float* row_mask, int rows, int D, cudaStream_t stream) {
  SC_CHECK(rows > 0 && D > 0, SC_ERR_SHAPE, "sc_mask_rows: rows=%d D=%d", rows, D);
  mask_rows_kernel<<<grid_for((size_t)rows * D, 256), 256, 0, stream>>>(x, row_mask, rows, D);
  SC_LAUNCH_CHECK("sc_mask_rows");
  return SC_OK;
}

int sc_ingest_f32_bf16(const float* pinned_host, void* y, size_t n, int ctas, cudaStream_t stream) {
  SC_CHECK(n > 0 && n % 4 == 0, SC_ERR_SHAPE, "sc_ingest_f32_bf16: n=%zu must be a positive multiple of 4", n);
  SC_CHECK(((uintptr_t)pinned_host & 15) == 0 && ((uintptr_t)y & 7) == 0, SC_ERR_ALIGN, "sc_ingest_f32_bf16: alignment");
  cudaPointerAttributes at;
  SC_CHECK(cudaPointerGetAttributes(&at, pinned_host) == cudaSuccess && at.type == cudaMemoryTypeHost && at.devicePointer != nullptr,
           SC_ERR_UNSUPPORTED, "sc_ingest_f32_bf16: the source must be pinned host memory mapped into the device address space");
  if (ctas <= 0) ctas = 64;
  ingest_f32_bf16_kernel<<<ctas, 256, 0, stream>>>((const float*)at.devicePointer, (__nv_bfloat16*)y, n / 4);
  SC_LAUNCH_CHECK("sc_ingest_f32_bf16");
  return SC_OK;
}

int sc_cast_f32_bf16(const float* x, void* y, size_t n, cudaStream_t stream) {
  SC_CHECK(n > 0, SC_ERR_SHAPE, "sc_cast: n=0");
  SC_CHECK(((uintptr_t)x & 15) == 0 && ((uintptr_t)y & 7) == 0, SC_ERR_ALIGN, "sc_cast: alignment");
  cast_kernel<<<grid_for(n / 4 + 1, 256), 256, 0, stream>>>(x, (__nv_bfloat16*)y, n);
  SC_LAUNCH_CHECK("sc_cast_f32_bf16");
  return SC_OK;
}

}  // extern "C"
