// Training-side row / elementwise kernels (K2 helpers, K9 backward, K10).
//   sc_prep_grad           : G (.) relu/dropout mask -> G (T) and G^T (T)   (operands of dgrad / wgrad GEMMs)
//   sc_transpose           : X[rows,cols] -> X^T (cast)                       (wgrad operand)
//   sc_apply_mask_transposed: (W (.) mask)^T -> [K,N]                         (dgrad operand; sampler.py straight-through)
//   sc_mask_grad           : dWm -> dW = dWm (.) m, dS = dWm (.) W (.) sigmoid'(S) [+ sparsity-loss term]
//                            (sparse_caption/pruning/sampler.py:15-17,32-34; prune.py:249-258)
//   sc_colsum              : bias gradient
//   sc_layernorm_bwd       : backward of a*(x-mean)/(std_unbiased+eps)+b      (models/transformer.py:329-341)
//   sc_logsoftmax_nll      : log_softmax + LanguageModelCriterion fwd/bwd     (transformer.py:413; utils/losses.py:32-43)
//   sc_log_softmax         : plain log-probs for the `_forward` API
//   sc_embedding_bwd       : scatter-add of token-row gradients
//   sc_adam_clip           : clip_grad_value_ + Adam on a flat buffer         (utils/optim.py:116-126,187-191)
#include "sc_common.cuh"

namespace {

template <typename T> __device__ __forceinline__ T cvt(float v) { return sc::from_f32<T>(v); }

// ---------------------------------------------------------------------------------------------
// prep_grad: out[r,c] = g[r,c] * keep(r,c) * scale ; outT[c,r] = same.  keep = (h[r,c] != 0) when h is given
// (ReLU and/or dropout already folded into the saved activation), else Philox dropout mask when p > 0.
// ---------------------------------------------------------------------------------------------
template <typename HT, typename OT>
__global__ void __launch_bounds__(256) prep_grad_kernel(const float* __restrict__ g, const HT* __restrict__ h, OT* __restrict__ out,
                                                        OT* __restrict__ outT, int ldT, int rows, int cols, float scale, float p,
                                                        unsigned long long seed, unsigned long long stream,
                                                        float* __restrict__ colsum) {
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  const sc::Philox ph(seed);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = r0 + ty + i * 8, c = c0 + tx;
    float v = 0.f;
    if (r < rows && c < cols) {
      const size_t e = (size_t)r * cols + c;
      v = g[e];
      if (h) v = (sc::to_f32<HT>(h[e]) != 0.f) ? v * scale : 0.f;
      else if (p > 0.f) v *= sc::keep_scale(ph, e, stream, p);
      else v *= scale;
      if (out) out[e] = cvt<OT>(v);
    }
    tile[ty + i * 8][tx] = v;
  }
  if (!outT && !colsum) return;
  __syncthreads();
  if (colsum && ty == 0 && c0 + tx < cols) {
    // bias gradient: column sums of the masked gradient (in the precision the GEMMs consume), one atomic per tile column
    float t = 0.f;
#pragma unroll
    for (int r = 0; r < 32; ++r) t += sc::to_f32<OT>(cvt<OT>(tile[r][tx]));
    atomicAdd(colsum + c0 + tx, t);
  }
  if (!outT) return;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = c0 + ty + i * 8, r = r0 + tx;
    if (r < rows && c < cols) outT[(size_t)c * ldT + r] = cvt<OT>(tile[tx][ty + i * 8]);
  }
}

template <typename IT, typename OT>
__global__ void __launch_bounds__(256) transpose_kernel(const IT* __restrict__ x, OT* __restrict__ y, int ldT, int rows, int cols) {
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = r0 + ty + i * 8, c = c0 + tx;
    tile[ty + i * 8][tx] = (r < rows && c < cols) ? sc::to_f32<IT>(x[(size_t)r * cols + c]) : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = c0 + ty + i * 8, r = r0 + tx;
    if (r < rows && c < cols) y[(size_t)c * ldT + r] = cvt<OT>(tile[tx][ty + i * 8]);
  }
}

using sc::mask_value;

template <typename OT>
__global__ void __launch_bounds__(256) apply_mask_t_kernel(const float* __restrict__ w, const float* __restrict__ mask, int mode,
                                                           const float* __restrict__ uni, unsigned long long seed,
                                                           unsigned long long stream, OT* __restrict__ outT, int N, int Kd,
                                                           OT* __restrict__ out) {
  __shared__ float tile[32][33];
  const int k0 = blockIdx.x * 32, n0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const sc::Philox ph(seed);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int n = n0 + ty + i * 8, k = k0 + tx;
    float v = 0.f;
    if (n < N && k < Kd) {
      const size_t e = (size_t)n * Kd + k;
      v = w[e] * mask_value(mode, mask ? mask[e] : 0.f, uni ? uni[e] : 0.f, ph, e, stream);
      if (out) out[e] = cvt<OT>(v);  // the forward operand W (.) m from the same pass (and the same mask sample)
    }
    tile[ty + i * 8][tx] = v;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int k = k0 + ty + i * 8, n = n0 + tx;
    if (n < N && k < Kd) outT[(size_t)k * N + n] = cvt<OT>(tile[tx][ty + i * 8]);
  }
}

// dW = dWm * m ; dS = dWm * W * sigmoid'(S) (or * 1 when bypass / raw) + sp_coeff * sigmoid'(S)
__global__ void __launch_bounds__(256) mask_grad_kernel(const float* __restrict__ dwm, const float* __restrict__ w,
                                                        const float* __restrict__ s, int mode, const float* __restrict__ uni,
                                                        unsigned long long seed, unsigned long long stream, int bypass,
                                                        float sp_coeff, float* __restrict__ dw, float* __restrict__ ds,
                                                        int accumulate, size_t n) {
  const sc::Philox ph(seed);
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (size_t)gridDim.x * blockDim.x) {
    const float g = dwm[e];
    const float sv = s ? s[e] : 0.f;
    const float m = mask_value(mode, sv, uni ? uni[e] : 0.f, ph, e, stream);
    float gw, gs;
    sc::mask_grad_elem(mode, g, w[e], sv, m, bypass, sp_coeff, gw, gs);
    if (dw) dw[e] = (accumulate ? dw[e] : 0.f) + gw;
    if (ds) ds[e] = (accumulate ? ds[e] : 0.f) + gs;
  }
}

// Split-K reduction fused with the straight-through mask gradient: dWm = sum_s part[s]; dW = dWm (.) m;
// dS = dWm (.) W (.) sigmoid'(S) + sp_coeff sigmoid'(S).  One float4 per thread, every operand streamed once, coalesced.
__global__ void __launch_bounds__(256) mask_grad_reduce_kernel(const float* __restrict__ part, int splits, size_t stride,
                                                               const float* __restrict__ w, const float* __restrict__ s, int mode,
                                                               const float* __restrict__ uni, unsigned long long seed,
                                                               unsigned long long stream, int bypass, float sp_coeff,
                                                               float* __restrict__ dw, float* __restrict__ ds, int accumulate,
                                                               size_t n4) {
  const sc::Philox ph(seed);
  sc::pdl_wait();
  for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < n4; g += (size_t)gridDim.x * blockDim.x) {
    const size_t e = g * 4;
    float4 acc = *(const float4*)(part + e);
    for (int k = 1; k < splits; ++k) {
      const float4 p = *(const float4*)(part + (size_t)k * stride + e);
      acc.x += p.x; acc.y += p.y; acc.z += p.z; acc.w += p.w;
    }
    const float4 w4 = __ldg((const float4*)(w + e));
    float4 s4 = make_float4(0.f, 0.f, 0.f, 0.f), u4 = s4;
    if (s) s4 = __ldg((const float4*)(s + e));
    if (uni) u4 = __ldg((const float4*)(uni + e));
    const float gv[4] = {acc.x, acc.y, acc.z, acc.w}, wv[4] = {w4.x, w4.y, w4.z, w4.w};
    const float sv[4] = {s4.x, s4.y, s4.z, s4.w}, uv[4] = {u4.x, u4.y, u4.z, u4.w};
    float m[4];
    if (mode == SC_MASK_BERNOULLI) {
      sc::bernoulli4(ph, g, stream, sv, m);
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) m[i] = sc::mask_value(mode, sv[i], uv[i], ph, e + i, stream);
    }
    float gw[4], gs[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) sc::mask_grad_elem(mode, gv[i], wv[i], sv[i], m[i], bypass, sp_coeff, gw[i], gs[i]);
    if (dw) {
      float4 o = make_float4(gw[0], gw[1], gw[2], gw[3]);
      if (accumulate) { const float4 p = *(const float4*)(dw + e); o.x += p.x; o.y += p.y; o.z += p.z; o.w += p.w; }
      *(float4*)(dw + e) = o;
    }
    if (ds) {
      float4 o = make_float4(gs[0], gs[1], gs[2], gs[3]);
      if (accumulate) { const float4 p = *(const float4*)(ds + e); o.x += p.x; o.y += p.y; o.z += p.z; o.w += p.w; }
      *(float4*)(ds + e) = o;
    }
  }
}

// column sums of a [rows, cols] matrix: one CTA per 32 columns x row slice, partial sums added atomically
template <typename T>
__global__ void __launch_bounds__(256) colsum_split_kernel(const T* __restrict__ x, float* __restrict__ out, int rows, int cols,
                                                           int rows_per_cta) {
  __shared__ float part[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx;
  const int r0 = blockIdx.y * rows_per_cta, r1 = min(rows, r0 + rows_per_cta);
  float s = 0.f;
  if (c < cols)
    for (int r = r0 + ty; r < r1; r += 8) s += sc::to_f32<T>(x[(size_t)r * cols + c]);
  part[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && c < cols) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += part[i][tx];
    atomicAdd(out + c, t);
  }
}

// column sums of a [rows, cols] matrix: one CTA per 32 columns
template <typename T>
__global__ void __launch_bounds__(256) colsum_kernel(const T* __restrict__ x, float* __restrict__ out, int rows, int cols,
                                                     int accumulate) {
  __shared__ float part[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx;
  float s = 0.f;
  if (c < cols)
    for (int r = ty; r < rows; r += 8) s += sc::to_f32<T>(x[(size_t)r * cols + c]);
  part[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && c < cols) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += part[i][tx];
    out[c] = (accumulate ? out[c] : 0.f) + t;
  }
}

// LayerNorm backward: one warp per row.  y = a * (x - mu) * r + b, r = 1/(sigma + eps), sigma unbiased.
//   dx_i = r (g_i - mean(g)) - r^2 (sum_j g_j c_j) c_i / ((D-1) sigma),   g = dy * a,  c = x - mu
//   da += sum_rows dy * c * r ; db += sum_rows dy   (smem reduction per CTA, one atomicAdd per column per CTA)
template <typename GT, int kMaxPerLane>
__global__ void __launch_bounds__(256) layernorm_bwd_kernel(const float* __restrict__ x, const float* __restrict__ a,
                                                            const GT* __restrict__ dy, const float* __restrict__ dres,
                                                            float* __restrict__ dx, float* __restrict__ da,
                                                            float* __restrict__ db, int rows, int D, float eps) {
  extern __shared__ float s_acc[];  // [2][D]
  for (int i = threadIdx.x; i < 2 * D; i += blockDim.x) s_acc[i] = 0.f;
  __syncthreads();
  const int warp_in_cta = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // da / db partial sums of the rows this warp handles stay in registers; shared-memory and global atomics happen once
  // per warp / per CTA (the per-row version serialised 532 CTAs x 1024 global atomics on the same addresses)
  float acc_a[kMaxPerLane], acc_b[kMaxPerLane];
#pragma unroll
  for (int i = 0; i < kMaxPerLane; ++i) { acc_a[i] = 0.f; acc_b[i] = 0.f; }
  for (int row = blockIdx.x * 8 + warp_in_cta; row < rows; row += gridDim.x * 8) {
    const float* xr = x + (size_t)row * D;
    float c[kMaxPerLane], g[kMaxPerLane];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < kMaxPerLane; ++i) {
      const int col = lane + i * 32;
      c[i] = col < D ? xr[col] : 0.f;
      s += c[i];
    }
    const float mu = sc::warp_sum(s) / (float)D;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < kMaxPerLane; ++i) {
      const int col = lane + i * 32;
      c[i] = col < D ? c[i] - mu : 0.f;
      q += c[i] * c[i];
    }
    const float sigma = sqrtf(sc::warp_sum(q) / (float)(D - 1));
    const float r = 1.f / (sigma + eps);
    float sg = 0.f, sgc = 0.f;
#pragma unroll
    for (int i = 0; i < kMaxPerLane; ++i) {
      const int col = lane + i * 32;
      float dyv = 0.f;
      if (col < D) {
        dyv = sc::to_f32<GT>(dy[(size_t)row * D + col]);
        acc_a[i] += dyv * c[i] * r;
        acc_b[i] += dyv;
        g[i] = dyv * a[col];
      } else {
        g[i] = 0.f;
      }
      sg += g[i];
      sgc += g[i] * c[i];
    }
    sg = sc::warp_sum(sg) / (float)D;
    sgc = sc::warp_sum(sgc);
    const float k2 = (sigma > 0.f) ? r * r * sgc / ((float)(D - 1) * sigma) : 0.f;
#pragma unroll
    for (int i = 0; i < kMaxPerLane; ++i) {
      const int col = lane + i * 32;
      if (col < D) {
        float v = r * (g[i] - sg) - k2 * c[i];
        if (dres) v += dres[(size_t)row * D + col];
        dx[(size_t)row * D + col] = v;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < kMaxPerLane; ++i) {
    const int col = lane + i * 32;
    if (col < D) {
      atomicAdd(&s_acc[col], acc_a[i]);
      atomicAdd(&s_acc[D + col], acc_b[i]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < D; i += blockDim.x) {
    atomicAdd(&da[i], s_acc[i]);
    atomicAdd(&db[i], s_acc[D + i]);
  }
}

// log-softmax + masked NLL, forward and backward in one pass over the logits (one CTA per row).
//   loss_sum += -w_r * logp[target_r]   ;   dlogits = (softmax - onehot) * w_r * inv_norm   (written in GT)
template <typename GT>
__global__ void __launch_bounds__(256) logsoftmax_nll_kernel(const float* __restrict__ logits, const int* __restrict__ target,
                                                             const float* __restrict__ weight, const float* __restrict__ inv_norm,
                                                             float* __restrict__ loss_sum, GT* __restrict__ dlogits,
                                                             float* __restrict__ logprobs, int V) {
  __shared__ float s_red[8];
  const int r = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* x = logits + (size_t)r * V;
  float mx = -INFINITY;
  for (int i = tid; i < V; i += 256) mx = fmaxf(mx, x[i]);
  mx = sc::warp_max(mx);
  if (lane == 0) s_red[warp] = mx;
  __syncthreads();
  mx = s_red[0];
#pragma unroll
  for (int w = 1; w < 8; ++w) mx = fmaxf(mx, s_red[w]);
  __syncthreads();
  float se = 0.f;
  for (int i = tid; i < V; i += 256) se += expf(x[i] - mx);
  se = sc::warp_sum(se);
  if (lane == 0) s_red[warp] = se;
  __syncthreads();
  se = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) se += s_red[w];
  const float ls = logf(se);
  if (logprobs) {
    for (int i = tid; i < V; i += 256) logprobs[(size_t)r * V + i] = (x[i] - mx) - ls;
  }
  if (target) {
    const int tg = target[r];
    const float w = weight ? weight[r] : 1.f;
    if (tid == 0 && loss_sum && w != 0.f) atomicAdd(loss_sum, -w * ((x[tg] - mx) - ls));
    if (dlogits) {
      const float sc_ = w * (inv_norm ? *inv_norm : 1.f);
      for (int i = tid; i < V; i += 256) {
        const float p = expf((x[i] - mx) - ls);
        dlogits[(size_t)r * V + i] = cvt<GT>((p - (i == tg ? 1.f : 0.f)) * sc_);
      }
    }
  }
}

__global__ void __launch_bounds__(128) embedding_bwd_kernel(const int* __restrict__ tokens, const float* __restrict__ dy,
                                                            float* __restrict__ dtable, int rows, int D, int V, float scale) {
  const int r = blockIdx.x;
  int tok = tokens[r];
  tok = tok < 0 ? 0 : (tok >= V ? V - 1 : tok);
  for (int c = threadIdx.x; c < D; c += blockDim.x) atomicAdd(&dtable[(size_t)tok * D + c], dy[(size_t)r * D + c] * scale);
}

// clip_grad_value_(clip) then Adam (torch.optim.Adam semantics: L2 weight decay added to the gradient,
// bias-corrected moments, eps added outside the sqrt).  grad_scale: e.g. 1/world after a sum all-reduce.
__global__ void __launch_bounds__(256) adam_clip_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                        float* __restrict__ v, size_t n, float lr, float b1, float b2, float eps,
                                                        float wd, float clip, float grad_scale, float bc1, float bc2_sqrt,
                                                        const float* __restrict__ sig_coeff, const float* __restrict__ dyn) {
  // sig_coeff (mask-logit group only): gradient of the sparsity loss, coeff * sigmoid'(S), with S = the parameter itself
  const float sc_ = sig_coeff ? *sig_coeff : 0.f;
  if (dyn) { lr = dyn[0]; bc1 = dyn[1]; bc2_sqrt = dyn[2]; }  // per-step values of a captured graph
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    float gi = g[i] * grad_scale;
    const float pi = p[i];
    if (sc_ != 0.f) { const float sg = sc::sigmoidf_(pi); gi += sc_ * sg * (1.f - sg); }
    if (clip > 0.f) gi = fminf(fmaxf(gi, -clip), clip);
    if (wd != 0.f) gi += wd * pi;
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi; v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] = pi - (lr / bc1) * (mi / denom);
  }
}

// PruningMixin.compute_sparsity_loss (pruning/prune.py:228-269) from the binarized-mask count:
//   out[0] = |target - sparsity| ; out[1] = d(scaled loss)/d(nnz) = sign(target - sparsity) / total * scale ; out[2] = sparsity
__global__ void sparsity_coeff_kernel(const unsigned long long* count, double total, float target, float scale,
                                      const float* scale_dev, float* out) {
  if (scale_dev) scale = *scale_dev;
  const double sparsity = 1.0 - (double)(*count) / total;
  const double diff = (double)target - sparsity;
  out[0] = (float)fabs(diff);
  out[1] = (float)((diff >= 0 ? 1.0 : -1.0) / total * (double)scale);
  out[2] = (float)sparsity;
}

int grid_for(size_t n, int block) {
  size_t g = (n + block - 1) / block;
  const size_t cap = 148 * 16;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace

extern "C" {

int sc_prep_grad(const float* g, const void* h, int h_dtype, void* out, void* outT, int ldT, int out_dtype, int rows, int cols,
                 float scale, float dropout_p, unsigned long long seed, unsigned long long stream_id, float* colsum_accum,
                 cudaStream_t stream) {
  SC_CHECK(rows > 0 && cols > 0, SC_ERR_SHAPE, "sc_prep_grad: rows=%d cols=%d", rows, cols);
  dim3 grid((cols + 31) / 32, (rows + 31) / 32);
#define PG(HT, OT) prep_grad_kernel<HT, OT><<<grid, 256, 0, stream>>>(g, (const HT*)h, (OT*)out, (OT*)outT, ldT, rows, cols, scale, dropout_p, seed, stream_id, colsum_accum)
  if (out_dtype == SC_BF16) { if (h_dtype == SC_BF16) PG(__nv_bfloat16, __nv_bfloat16); else PG(float, __nv_bfloat16); }
  else if (out_dtype == SC_F32) { if (h_dtype == SC_BF16) PG(__nv_bfloat16, float); else PG(float, float); }
  else SC_CHECK(false, SC_ERR_DTYPE, "sc_prep_grad: bad dtype");
#undef PG
  SC_LAUNCH_CHECK("sc_prep_grad");
  return SC_OK;
}

int sc_transpose(const void* x, int x_dtype, void* y, int ldT, int y_dtype, int rows, int cols, cudaStream_t stream) {
  SC_CHECK(rows > 0 && cols > 0 && ldT >= rows, SC_ERR_SHAPE, "sc_transpose: rows=%d cols=%d ldT=%d", rows, cols, ldT);
  dim3 grid((cols + 31) / 32, (rows + 31) / 32);
  if (x_dtype == SC_F32 && y_dtype == SC_F32) transpose_kernel<float, float><<<grid, 256, 0, stream>>>((const float*)x, (float*)y, ldT, rows, cols);
  else if (x_dtype == SC_F32 && y_dtype == SC_BF16) transpose_kernel<float, __nv_bfloat16><<<grid, 256, 0, stream>>>((const float*)x, (__nv_bfloat16*)y, ldT, rows, cols);
  else if (x_dtype == SC_BF16 && y_dtype == SC_BF16) transpose_kernel<__nv_bfloat16, __nv_bfloat16><<<grid, 256, 0, stream>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)y, ldT, rows, cols);
  else if (x_dtype == SC_BF16 && y_dtype == SC_F32) transpose_kernel<__nv_bfloat16, float><<<grid, 256, 0, stream>>>((const __nv_bfloat16*)x, (float*)y, ldT, rows, cols);
  else SC_CHECK(false, SC_ERR_DTYPE, "sc_transpose: bad dtypes %d %d", x_dtype, y_dtype);
  SC_LAUNCH_CHECK("sc_transpose");
  return SC_OK;
}

int sc_apply_mask_transposed(const float* w, const float* mask, int mask_mode, const float* uniforms, unsigned long long seed,
                             unsigned long long stream_id, void* outT, int out_dtype, int N, int K, void* out_plain,
                             cudaStream_t stream) {
  SC_CHECK(N > 0 && K > 0, SC_ERR_SHAPE, "sc_apply_mask_transposed: N=%d K=%d", N, K);
  SC_CHECK(mask_mode == SC_MASK_NONE || mask != nullptr, SC_ERR_SHAPE, "sc_apply_mask_transposed: mask missing");
  dim3 grid((K + 31) / 32, (N + 31) / 32);
  if (out_dtype == SC_BF16) apply_mask_t_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(w, mask, mask_mode, uniforms, seed, stream_id, (__nv_bfloat16*)outT, N, K, (__nv_bfloat16*)out_plain);
  else if (out_dtype == SC_F32) apply_mask_t_kernel<float><<<grid, 256, 0, stream>>>(w, mask, mask_mode, uniforms, seed, stream_id, (float*)outT, N, K, (float*)out_plain);
  else SC_CHECK(false, SC_ERR_DTYPE, "sc_apply_mask_transposed: bad dtype");
  SC_LAUNCH_CHECK("sc_apply_mask_transposed");
  return SC_OK;
}

int sc_mask_grad(const float* dwm, const float* w, const float* mask, int mask_mode, const float* uniforms,
                 unsigned long long seed, unsigned long long stream_id, int bypass_sigmoid_grad, float sparsity_coeff,
                 float* dw, float* ds, int accumulate, size_t n, cudaStream_t stream) {
  SC_CHECK(n > 0, SC_ERR_SHAPE, "sc_mask_grad: n=0");
  mask_grad_kernel<<<grid_for(n, 256), 256, 0, stream>>>(dwm, w, mask, mask_mode, uniforms, seed, stream_id, bypass_sigmoid_grad,
                                                         sparsity_coeff, dw, ds, accumulate, n);
  SC_LAUNCH_CHECK("sc_mask_grad");
  return SC_OK;
}

// second kernel of the two-kernel weight gradient (sc_linear_wgrad with a workspace)
int sc_mask_grad_reduce_launch(const float* part, int splits, size_t stride, const float* w, const float* mask, int mask_mode,
                               const float* uniforms, unsigned long long seed, unsigned long long stream_id, int bypass,
                               float sp_coeff, float* dw, float* ds, int accumulate, size_t n, cudaStream_t stream) {
  SC_CHECK(n > 0 && n % 4 == 0, SC_ERR_SHAPE, "sc_mask_grad_reduce: n=%zu must be a positive multiple of 4", n);
  const size_t n4 = n / 4;
  cudaError_t e = sc::launch_pdl(mask_grad_reduce_kernel, dim3(grid_for(n4, 256)), dim3(256), 0, stream, part, splits, stride, w, mask,
                                 mask_mode, uniforms, seed, stream_id, bypass, sp_coeff, dw, ds, accumulate, n4);
  SC_CHECK(e == cudaSuccess, (int)e, "sc_mask_grad_reduce: %s", cudaGetErrorString(e));
  SC_LAUNCH_CHECK("sc_mask_grad_reduce");
  return SC_OK;
}

int sc_colsum(const void* x, int dtype, float* out, int rows, int cols, int accumulate, cudaStream_t stream) {
  SC_CHECK(rows > 0 && cols > 0, SC_ERR_SHAPE, "sc_colsum: rows=%d cols=%d", rows, cols);
  if (rows >= 512) {
    // tall matrices (bias gradients over thousands of tokens): split the rows over ~2 CTAs per SM
    const int gx = (cols + 31) / 32;
    int gy = (2 * 148 + gx - 1) / gx;
    int rpc = (rows + gy - 1) / gy;
    rpc = (rpc + 7) / 8 * 8;
    gy = (rows + rpc - 1) / rpc;
    if (!accumulate) cudaMemsetAsync(out, 0, (size_t)cols * sizeof(float), stream);
    if (dtype == SC_F32) colsum_split_kernel<float><<<dim3(gx, gy), 256, 0, stream>>>((const float*)x, out, rows, cols, rpc);
    else if (dtype == SC_BF16) colsum_split_kernel<__nv_bfloat16><<<dim3(gx, gy), 256, 0, stream>>>((const __nv_bfloat16*)x, out, rows, cols, rpc);
    else SC_CHECK(false, SC_ERR_DTYPE, "sc_colsum: bad dtype");
    SC_LAUNCH_CHECK("sc_colsum");
    return SC_OK;
  }
  const int grid = (cols + 31) / 32;
  if (dtype == SC_F32) colsum_kernel<float><<<grid, 256, 0, stream>>>((const float*)x, out, rows, cols, accumulate);
  else if (dtype == SC_BF16) colsum_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>((const __nv_bfloat16*)x, out, rows, cols, accumulate);
  else SC_CHECK(false, SC_ERR_DTYPE, "sc_colsum: bad dtype");
  SC_LAUNCH_CHECK("sc_colsum");
  return SC_OK;
}

int sc_layernorm_bwd(const float* x, const float* a, const void* dy, int dy_dtype, const float* dres, float* dx, float* da,
                     float* db, int rows, int D, float eps, cudaStream_t stream) {
  SC_CHECK(rows > 0 && D > 1 && D <= 2048, SC_ERR_SHAPE, "sc_layernorm_bwd: rows=%d D=%d", rows, D);
  int blocks = (rows + 7) / 8;
  if (blocks > 148 * 2) blocks = 148 * 2;
  const size_t smem = 2 * (size_t)D * sizeof(float);
#define LNB(T, P) layernorm_bwd_kernel<T, P><<<blocks, 256, smem, stream>>>(x, a, (const T*)dy, dres, dx, da, db, rows, D, eps)
  if (dy_dtype == SC_F32) { if (D <= 128) LNB(float, 4); else if (D <= 512) LNB(float, 16); else LNB(float, 64); }
  else if (dy_dtype == SC_BF16) { if (D <= 128) LNB(__nv_bfloat16, 4); else if (D <= 512) LNB(__nv_bfloat16, 16); else LNB(__nv_bfloat16, 64); }
  else SC_CHECK(false, SC_ERR_DTYPE, "sc_layernorm_bwd: bad dtype");
#undef LNB
  SC_LAUNCH_CHECK("sc_layernorm_bwd");
  return SC_OK;
}

int sc_logsoftmax_nll(const float* logits, const int* target, const float* weight, const float* inv_norm, float* loss_sum,
                      void* dlogits, int d_dtype, float* logprobs, int rows, int V, cudaStream_t stream) {
  SC_CHECK(rows > 0 && V > 0, SC_ERR_SHAPE, "sc_logsoftmax_nll: rows=%d V=%d", rows, V);
  if (d_dtype == SC_BF16)
    logsoftmax_nll_kernel<__nv_bfloat16><<<rows, 256, 0, stream>>>(logits, target, weight, inv_norm, loss_sum, (__nv_bfloat16*)dlogits, logprobs, V);
  else if (d_dtype == SC_F32)
    logsoftmax_nll_kernel<float><<<rows, 256, 0, stream>>>(logits, target, weight, inv_norm, loss_sum, (float*)dlogits, logprobs, V);
  else SC_CHECK(false, SC_ERR_DTYPE, "sc_logsoftmax_nll: bad dtype");
  SC_LAUNCH_CHECK("sc_logsoftmax_nll");
  return SC_OK;
}

int sc_embedding_bwd(const int* tokens, const float* dy, float* dtable, int rows, int D, int V, float scale, cudaStream_t stream) {
  SC_CHECK(rows > 0 && D > 0 && V > 0, SC_ERR_SHAPE, "sc_embedding_bwd: bad shape");
  embedding_bwd_kernel<<<rows, 128, 0, stream>>>(tokens, dy, dtable, rows, D, V, scale);
  SC_LAUNCH_CHECK("sc_embedding_bwd");
  return SC_OK;
}

int sc_sparsity_coeff(const unsigned long long* count, double total, float target, float scale, const float* scale_dev,
                      float* out3, cudaStream_t stream) {
  SC_CHECK(count && out3 && total > 0, SC_ERR_SHAPE, "sc_sparsity_coeff: bad args");
  sparsity_coeff_kernel<<<1, 1, 0, stream>>>(count, total, target, scale, scale_dev, out3);
  SC_LAUNCH_CHECK("sc_sparsity_coeff");
  return SC_OK;
}

int sc_adam_clip(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, size_t n, float lr, float beta1, float beta2,
                 float eps, float weight_decay, float clip_value, float grad_scale, int step, const float* sigmoid_grad_coeff,
                 const float* dyn, cudaStream_t stream) {
  SC_CHECK(n > 0 && step >= 1, SC_ERR_SHAPE, "sc_adam_clip: n=%zu step=%d", n, step);
  const float bc1 = 1.f - powf(beta1, (float)step);
  const float bc2 = sqrtf(1.f - powf(beta2, (float)step));
  adam_clip_kernel<<<grid_for(n, 256), 256, 0, stream>>>(param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, weight_decay,
                                                         clip_value, grad_scale, bc1, bc2, sigmoid_grad_coeff, dyn);
  SC_LAUNCH_CHECK("sc_adam_clip");
  return SC_OK;
}

}  // extern "C"
