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


// Row-major-only variant (the bf16 trainer feeds dy to the weight-gradient GEMM as MN-major tiles, so no transposed copy
// is needed): 16-byte loads / 8-byte bf16 stores, CTA = 64 rows x 128 columns, one Philox draw per 4 elements, bias
// gradient accumulated in registers over the thread's 8 rows -> shared memory -> one vector reduction per 4 columns.
template <typename HT>
__global__ void __launch_bounds__(256) prep_grad_vec_kernel(const float* __restrict__ g, const HT* __restrict__ h,
                                                            __nv_bfloat16* __restrict__ out, int rows, int cols, float scale, float p,
                                                            unsigned long long seed, unsigned long long stream,
                                                            float* __restrict__ colsum) {
  __shared__ float4 s_cs[8][32];
  sc::pdl_launch();
  sc::pdl_wait();
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 128 + tx * 4;
  const int r0 = blockIdx.y * 64 + ty;
  const sc::Philox ph(seed);
  const float keep = p > 0.f ? 1.f / (1.f - p) : 1.f;
  float4 gv[8];
  float hk[8][4];
  bool ok[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = r0 + 8 * i;
    ok[i] = r < rows && c < cols;
    gv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    hk[i][0] = hk[i][1] = hk[i][2] = hk[i][3] = 1.f;
    if (ok[i]) {
      const size_t e = (size_t)r * cols + c;
      gv[i] = *(const float4*)(g + e);
      if (h) {
        if (sizeof(HT) == 2) {
          const uint2 hv = *(const uint2*)((const __nv_bfloat16*)h + e);
          hk[i][0] = __uint_as_float(hv.x << 16); hk[i][1] = __uint_as_float(hv.x & 0xffff0000u);
          hk[i][2] = __uint_as_float(hv.y << 16); hk[i][3] = __uint_as_float(hv.y & 0xffff0000u);
        } else {
          const float4 hv = *(const float4*)((const float*)h + e);
          hk[i][0] = hv.x; hk[i][1] = hv.y; hk[i][2] = hv.z; hk[i][3] = hv.w;
        }
      }
    }
  }
  float cs[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    if (!ok[i]) continue;
    const size_t e = (size_t)(r0 + 8 * i) * cols + c;
    float v[4] = {gv[i].x, gv[i].y, gv[i].z, gv[i].w};
    if (h) {
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = hk[i][j] != 0.f ? v[j] * scale : 0.f;
    } else if (p > 0.f) {
      const uint4 rr = ph(e >> 2, stream);
      v[0] *= sc::u24(rr.x) >= p ? keep : 0.f; v[1] *= sc::u24(rr.y) >= p ? keep : 0.f;
      v[2] *= sc::u24(rr.z) >= p ? keep : 0.f; v[3] *= sc::u24(rr.w) >= p ? keep : 0.f;
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] *= scale;
    }
    const __nv_bfloat162 lo = __floats2bfloat162_rn(v[0], v[1]), hi = __floats2bfloat162_rn(v[2], v[3]);
    uint2 o; o.x = *(const uint32_t*)&lo; o.y = *(const uint32_t*)&hi;
    *(uint2*)(out + e) = o;
    // the bias gradient sums the values the GEMMs consume (bf16-rounded)
    cs[0] += __low2float(lo); cs[1] += __high2float(lo); cs[2] += __low2float(hi); cs[3] += __high2float(hi);
  }
  if (!colsum) return;
  s_cs[ty][tx] = make_float4(cs[0], cs[1], cs[2], cs[3]);
  __syncthreads();
  if (ty == 0 && c < cols) {
    float4 t = s_cs[0][tx];
#pragma unroll
    for (int w = 1; w < 8; ++w) { const float4 q = s_cs[w][tx]; t.x += q.x; t.y += q.y; t.z += q.z; t.w += q.w; }
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(colsum + c), "f"(t.x), "f"(t.y), "f"(t.z), "f"(t.w) : "memory");
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


// All masked weights of a training step in ONE launch (the per-tensor version was 68 launches of 6-10 us for 512x512
// tensors).  desc i = {w, s, u, out, outT, N, K, stream, tile_start, tiles_k} (10 x 64-bit words, device memory);
// CTA = one 64 x 64 tile of one tensor, found by binary search over tile_start.  One Philox draw per 4 elements
// (bernoulli4: the same sample mask_value() regenerates element-wise in the weight-gradient epilogue).
struct MaskDesc {
  const float* w; const float* s; const float* u; void* out; void* outT;
  long long N, K; unsigned long long stream; long long tile_start, tiles_k;
};

template <typename OT>
__global__ void __launch_bounds__(256, 4) apply_mask_batched_kernel(const MaskDesc* __restrict__ descs, int n_desc, int mode,
                                                                 unsigned long long seed, unsigned long long stream_base) {
  __shared__ float tile[64][65];
  int lo = 0, hi = n_desc - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (descs[mid].tile_start <= (long long)blockIdx.x) lo = mid; else hi = mid - 1;
  }
  const MaskDesc d = descs[lo];
  const int local = (int)((long long)blockIdx.x - d.tile_start);
  const int N = (int)d.N, Kd = (int)d.K;
  const int k0 = (local % (int)d.tiles_k) * 64, n0 = (local / (int)d.tiles_k) * 64;
  const unsigned long long stream = stream_base + d.stream;
  const sc::Philox ph(seed);
  const int tid = threadIdx.x;
  const int c4 = (tid & 15) * 4;
  OT* out = (OT*)d.out;
  // issue all loads of the thread first (4 rows x {w, s, u})
  float4 wv[4], sv[4], uv[4];
  bool ok[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int n = n0 + (tid >> 4) + 16 * i, k = k0 + c4;
    ok[i] = n < N && k < Kd;
    wv[i] = sv[i] = uv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ok[i]) {
      const size_t e = (size_t)n * Kd + k;
      wv[i] = __ldg((const float4*)(d.w + e));
      if (mode != SC_MASK_NONE) sv[i] = __ldg((const float4*)(d.s + e));
      if (mode == SC_MASK_UNIFORM) uv[i] = __ldg((const float4*)(d.u + e));
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = (tid >> 4) + 16 * i;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (ok[i]) {
      const size_t e = (size_t)(n0 + r) * Kd + k0 + c4;
      const float sa[4] = {sv[i].x, sv[i].y, sv[i].z, sv[i].w}, ua[4] = {uv[i].x, uv[i].y, uv[i].z, uv[i].w};
      float m[4];
      if (mode == SC_MASK_BERNOULLI) {
        sc::bernoulli4(ph, e >> 2, stream, sa, m);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) m[j] = sc::mask_value(mode, sa[j], ua[j], ph, e + j, stream);
      }
      v[0] = wv[i].x * m[0]; v[1] = wv[i].y * m[1]; v[2] = wv[i].z * m[2]; v[3] = wv[i].w * m[3];
      if (out) {
        if (sizeof(OT) == 2) {
          const __nv_bfloat162 lo2 = __floats2bfloat162_rn(v[0], v[1]), hi2 = __floats2bfloat162_rn(v[2], v[3]);
          uint2 o; o.x = *(const uint32_t*)&lo2; o.y = *(const uint32_t*)&hi2;
          *(uint2*)((__nv_bfloat16*)d.out + e) = o;
        } else {
          *(float4*)((float*)d.out + e) = make_float4(v[0], v[1], v[2], v[3]);
        }
      }
    }
    tile[r][c4] = v[0]; tile[r][c4 + 1] = v[1]; tile[r][c4 + 2] = v[2]; tile[r][c4 + 3] = v[3];
  }
  if (!d.outT) return;
  __syncthreads();
  // transposed store: a warp writes 64 consecutive n of one k row (two per lane)
  const int lane = tid & 31, wq = tid >> 5;
  OT* outT = (OT*)d.outT;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int kk = wq + 8 * i, k = k0 + kk, n = n0 + 2 * lane;
    if (k >= Kd || n >= N) continue;
    const float a = tile[2 * lane][kk], b = tile[2 * lane + 1][kk];
    OT* o = outT + (size_t)k * N + n;
    if ((N & 1) == 0) {
      if (sizeof(OT) == 2) { const __nv_bfloat162 pr = __floats2bfloat162_rn(a, b); *(uint32_t*)o = *(const uint32_t*)&pr; }
      else *(float2*)o = make_float2(a, b);
    } else {
      o[0] = cvt<OT>(a);
      if (n + 1 < N) o[1] = cvt<OT>(b);
    }
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

// bf16, cols % 8 == 0: 16-byte loads (8 columns per thread), CTA = 256 columns x a row slice, 8 row lanes reduced through
// shared memory, one vector reduction per 4 columns (the 32-column version read 64-byte row segments: 95 us for the
// [4250, 10000] generator gradient)
__global__ void __launch_bounds__(256) colsum_bf16_vec_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ out, int rows,
                                                              int cols, int rows_per_cta) {
  __shared__ float part[8][32][8];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 256 + tx * 8;
  const int r0 = blockIdx.y * rows_per_cta, r1 = min(rows, r0 + rows_per_cta);
  float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (c < cols) {
    for (int r = r0 + ty; r < r1; r += 32) {
      // four rows in flight per thread
      uint4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int rr = r + 8 * u;
        v[u] = rr < r1 ? *(const uint4*)(x + (size_t)rr * cols + c) : make_uint4(0u, 0u, 0u, 0u);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const uint32_t w[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
        for (int i = 0; i < 4; ++i) { s[2 * i] += __uint_as_float(w[i] << 16); s[2 * i + 1] += __uint_as_float(w[i] & 0xffff0000u); }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) part[ty][tx][i] = s[i];
  __syncthreads();
  // thread -> (column group tx, half: 4 of its 8 columns); ty < 2 does the final sums
  if (ty < 2 && c < cols) {
    float t[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int w = 0; w < 8; ++w)
#pragma unroll
      for (int i = 0; i < 4; ++i) t[i] += part[w][tx][ty * 4 + i];
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(out + c + ty * 4), "f"(t[0]), "f"(t[1]), "f"(t[2]), "f"(t[3]) : "memory");
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


// D = 512 (d_model): one row per warp, 16-byte accesses, every load of the row (x, dy, residual gradient) issued before
// the first reduction; parameter gradients go warp -> shared memory -> one vector reduction per 4 columns and CTA.
// Optional fused gradient preparation for the NEXT linear of the backward chain (the one whose output gradient is dx, i.e.
// o-proj / cross o-proj / ff2: y = x + dropout(linear(.))): gb = bf16(dx (.) dropout keep mask) - the mask the forward
// GEMM epilogue drew, regenerated from (seed, stream, element) - and its column sums (that linear's bias gradient).
struct LnbNext {
  __nv_bfloat16* gb; float* colsum; float p; unsigned long long seed, stream;
};
template <typename GT>
__global__ void __launch_bounds__(256) layernorm_bwd512_kernel(const float* __restrict__ x, const float* __restrict__ a,
                                                               const GT* __restrict__ dy, const float* __restrict__ dres,
                                                               float* __restrict__ dx, float* __restrict__ da,
                                                               float* __restrict__ db, int rows, float eps, const LnbNext nx) {
  constexpr int D = 512;
  __shared__ float4 s_part[3][8][128];
  sc::pdl_launch();
  sc::pdl_wait();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + warp;
  float4 pa[4], pb[4], pc[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) pa[i] = pb[i] = pc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (row < rows) {
    const size_t base = (size_t)row * D;
    float c[4][4], g[4][4], dyv[4][4], rs[4][4];
    float4 av[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int col = 4 * (lane + 32 * i);
      const float4 xv = *(const float4*)(x + base + col);
      c[i][0] = xv.x; c[i][1] = xv.y; c[i][2] = xv.z; c[i][3] = xv.w;
      if (sizeof(GT) == 2) {
        const uint2 dv = *(const uint2*)((const __nv_bfloat16*)dy + base + col);
        dyv[i][0] = __uint_as_float(dv.x << 16); dyv[i][1] = __uint_as_float(dv.x & 0xffff0000u);
        dyv[i][2] = __uint_as_float(dv.y << 16); dyv[i][3] = __uint_as_float(dv.y & 0xffff0000u);
      } else {
        const float4 dv = *(const float4*)((const float*)dy + base + col);
        dyv[i][0] = dv.x; dyv[i][1] = dv.y; dyv[i][2] = dv.z; dyv[i][3] = dv.w;
      }
      rs[i][0] = rs[i][1] = rs[i][2] = rs[i][3] = 0.f;
      if (dres) { const float4 rv = *(const float4*)(dres + base + col); rs[i][0] = rv.x; rs[i][1] = rv.y; rs[i][2] = rv.z; rs[i][3] = rv.w; }
      av[i] = __ldg((const float4*)(a + col));
    }
    float sm = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) sm += c[i][j];
    const float mu = sc::warp_sum(sm) / (float)D;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) { c[i][j] -= mu; q += c[i][j] * c[i][j]; }
    const float sigma = sqrtf(sc::warp_sum(q) / (float)(D - 1));
    const float r = 1.f / (sigma + eps);
    float sg = 0.f, sgc = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float aa[4] = {av[i].x, av[i].y, av[i].z, av[i].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        g[i][j] = dyv[i][j] * aa[j];
        sg += g[i][j];
        sgc += g[i][j] * c[i][j];
      }
      pa[i] = make_float4(dyv[i][0] * c[i][0] * r, dyv[i][1] * c[i][1] * r, dyv[i][2] * c[i][2] * r, dyv[i][3] * c[i][3] * r);
      pb[i] = make_float4(dyv[i][0], dyv[i][1], dyv[i][2], dyv[i][3]);
    }
    sg = sc::warp_sum(sg) / (float)D;
    sgc = sc::warp_sum(sgc);
    const float k2 = (sigma > 0.f) ? r * r * sgc / ((float)(D - 1) * sigma) : 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int col = 4 * (lane + 32 * i);
      float4 o;
      o.x = r * (g[i][0] - sg) - k2 * c[i][0] + rs[i][0];
      o.y = r * (g[i][1] - sg) - k2 * c[i][1] + rs[i][1];
      o.z = r * (g[i][2] - sg) - k2 * c[i][2] + rs[i][2];
      o.w = r * (g[i][3] - sg) - k2 * c[i][3] + rs[i][3];
      *(float4*)(dx + base + col) = o;
      if (nx.gb) {
        float v[4] = {o.x, o.y, o.z, o.w};
        if (nx.p > 0.f) {
          const sc::Philox ph(nx.seed);
          const float keep = 1.f / (1.f - nx.p);
          const uint4 rr = ph((base + col) >> 2, nx.stream);
          v[0] *= sc::u24(rr.x) >= nx.p ? keep : 0.f; v[1] *= sc::u24(rr.y) >= nx.p ? keep : 0.f;
          v[2] *= sc::u24(rr.z) >= nx.p ? keep : 0.f; v[3] *= sc::u24(rr.w) >= nx.p ? keep : 0.f;
        }
        const __nv_bfloat162 lo = __floats2bfloat162_rn(v[0], v[1]), hi = __floats2bfloat162_rn(v[2], v[3]);
        uint2 ob; ob.x = *(const uint32_t*)&lo; ob.y = *(const uint32_t*)&hi;
        *(uint2*)(nx.gb + base + col) = ob;
        pc[i] = make_float4(__low2float(lo), __high2float(lo), __low2float(hi), __high2float(hi));
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    s_part[0][warp][lane + 32 * i] = pa[i]; s_part[1][warp][lane + 32 * i] = pb[i];
    if (nx.colsum) s_part[2][warp][lane + 32 * i] = pc[i];
  }
  __syncthreads();
  for (int job = threadIdx.x; job < (nx.colsum ? 384 : 256); job += 256) {
    const int which = job >> 7, c4 = job & 127;
    float4 t = s_part[which][0][c4];
#pragma unroll
    for (int w = 1; w < 8; ++w) { const float4 q4 = s_part[which][w][c4]; t.x += q4.x; t.y += q4.y; t.z += q4.z; t.w += q4.w; }
    float* dst = (which == 0 ? da : which == 1 ? db : nx.colsum) + 4 * c4;
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(t.x), "f"(t.y), "f"(t.z), "f"(t.w) : "memory");
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

// backward of a plain log_softmax (OutputEmbedding, transformer.py:405-413) for callers that take the loss outside
// (LanguageModelCriterion / RewardCriterion on the returned log-probs): dx = dy - exp(lp) * sum(dy)
__global__ void __launch_bounds__(256) logsoftmax_bwd_kernel(const float* __restrict__ lp, const float* __restrict__ dy,
                                                             float* __restrict__ dx, int V) {
  __shared__ float s_red[8];
  const int r = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* g = dy + (size_t)r * V;
  const float* l = lp + (size_t)r * V;
  float s = 0.f;
  for (int i = tid; i < V; i += 256) s += g[i];
  s = sc::warp_sum(s);
  if (lane == 0) s_red[warp] = s;
  __syncthreads();
  s = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) s += s_red[w];
  for (int i = tid; i < V; i += 256) dx[(size_t)r * V + i] = g[i] - expf(l[i]) * s;
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

// Whole-model optimizer step in ONE launch, with the straight-through epilogue of the masked weights inside it.
// The backward leaves dWm = d loss / d (W (.) m) in the flat gradient buffer (that is what a data-parallel run all-reduces:
// half the bytes of dW and dS, which are both elementwise functions of dWm, W, S and the step's mask sample).  Per element
// of a masked tensor:  m = mask sample regenerated from Philox(seed, stream, element) (sampler.py:10-34);
//   dW = dWm * m ; dS = dWm * W * sigmoid'(S) (* 1 when bypass) + sparsity_coeff * sigmoid'(S)  (prune.py:249-258)
// then clip_grad_value_ + Adam for the weight group AND the mask-logit group (utils/optim.py:116-126,187-191;
// scripts/train_n_prune_transformer.py:67-82).  Unmasked parameters (biases, LayerNorm) take the plain update.
// desc = {w_off, s_off (-1: unmasked), n, stream, blk_start} (5 x 64-bit words); CTA = a chunk of kStChunk elements.
struct StDesc { long long w_off, s_off, n; unsigned long long stream; long long blk_start; };
constexpr int kStChunk = 2048;

struct StArgs {
  float* w; const float* g; float* mw; float* vw;      // weight group, flat
  float* s; float* ms; float* vs; const float* u;      // mask-logit group, flat (u: injected uniforms or NULL)
  int mode, bypass, update_s;
  unsigned long long seed, stream_base;
  float lr_w, eps_w, wd_w, lr_s, eps_s, b1, b2, clip, grad_scale, bc1, bc2_sqrt;
  const float* sig_coeff; const float* dyn;
};

__device__ __forceinline__ float adam_update(float p, float gi, float& m, float& v, float lr, float eps, float wd, float b1, float b2,
                                             float clip, float bc1, float bc2_sqrt) {
  if (clip > 0.f) gi = fminf(fmaxf(gi, -clip), clip);
  if (wd != 0.f) gi += wd * p;
  m = b1 * m + (1.f - b1) * gi;
  v = b2 * v + (1.f - b2) * gi * gi;
  return p - (lr / bc1) * (m / (sqrtf(v) / bc2_sqrt + eps));
}

__global__ void __launch_bounds__(256) adam_clip_st_kernel(const StDesc* __restrict__ descs, int n_desc, StArgs a) {
  int lo = 0, hi = n_desc - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (descs[mid].blk_start <= (long long)blockIdx.x) lo = mid; else hi = mid - 1;
  }
  const StDesc d = descs[lo];
  const long long e0 = ((long long)blockIdx.x - d.blk_start) * kStChunk;
  float lr_w = a.lr_w, lr_s = a.lr_s, bc1 = a.bc1, bc2 = a.bc2_sqrt;
  if (a.dyn) { lr_w = a.dyn[0]; bc1 = a.dyn[1]; bc2 = a.dyn[2]; lr_s = a.dyn[3]; }
  const float spc = a.sig_coeff ? *a.sig_coeff : 0.f;
  const sc::Philox ph(a.seed);
  const unsigned long long stream = a.stream_base + d.stream;
  const bool vec = ((d.w_off | d.n) & 3) == 0 && (d.s_off < 0 || (d.s_off & 3) == 0);
  if (vec) {
    for (long long e = e0 + (long long)threadIdx.x * 4; e < e0 + kStChunk && e < d.n; e += 256 * 4) {
      const size_t iw = (size_t)(d.w_off + e);
      float4 g4 = *(const float4*)(a.g + iw), w4 = *(const float4*)(a.w + iw), mw4 = *(const float4*)(a.mw + iw), vw4 = *(const float4*)(a.vw + iw);
      float g[4] = {g4.x * a.grad_scale, g4.y * a.grad_scale, g4.z * a.grad_scale, g4.w * a.grad_scale};
      float w[4] = {w4.x, w4.y, w4.z, w4.w}, mw[4] = {mw4.x, mw4.y, mw4.z, mw4.w}, vw[4] = {vw4.x, vw4.y, vw4.z, vw4.w};
      if (d.s_off >= 0) {
        const size_t is = (size_t)(d.s_off + e);
        float4 s4 = *(const float4*)(a.s + is);
        float sv[4] = {s4.x, s4.y, s4.z, s4.w}, m[4];
        if (a.mode == SC_MASK_BERNOULLI) {
          sc::bernoulli4(ph, (uint64_t)e >> 2, stream, sv, m);
        } else {
          float uv[4] = {0.f, 0.f, 0.f, 0.f};
          if (a.mode == SC_MASK_UNIFORM) { const float4 u4 = *(const float4*)(a.u + is); uv[0] = u4.x; uv[1] = u4.y; uv[2] = u4.z; uv[3] = u4.w; }
#pragma unroll
          for (int j = 0; j < 4; ++j) m[j] = sc::mask_value(a.mode, sv[j], uv[j], ph, (size_t)e + j, stream);
        }
        float dw[4], ds[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) sc::mask_grad_elem(a.mode, g[j], w[j], sv[j], m[j], a.bypass, spc, dw[j], ds[j]);
        if (a.update_s) {
          float4 ms4 = *(const float4*)(a.ms + is), vs4 = *(const float4*)(a.vs + is);
          float ms[4] = {ms4.x, ms4.y, ms4.z, ms4.w}, vs[4] = {vs4.x, vs4.y, vs4.z, vs4.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) sv[j] = adam_update(sv[j], ds[j], ms[j], vs[j], lr_s, a.eps_s, 0.f, a.b1, a.b2, a.clip, bc1, bc2);
          *(float4*)(a.s + is) = make_float4(sv[0], sv[1], sv[2], sv[3]);
          *(float4*)(a.ms + is) = make_float4(ms[0], ms[1], ms[2], ms[3]);
          *(float4*)(a.vs + is) = make_float4(vs[0], vs[1], vs[2], vs[3]);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) g[j] = dw[j];
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) w[j] = adam_update(w[j], g[j], mw[j], vw[j], lr_w, a.eps_w, a.wd_w, a.b1, a.b2, a.clip, bc1, bc2);
      *(float4*)(a.w + iw) = make_float4(w[0], w[1], w[2], w[3]);
      *(float4*)(a.mw + iw) = make_float4(mw[0], mw[1], mw[2], mw[3]);
      *(float4*)(a.vw + iw) = make_float4(vw[0], vw[1], vw[2], vw[3]);
    }
    return;
  }
  for (long long e = e0 + threadIdx.x; e < e0 + kStChunk && e < d.n; e += 256) {
    const size_t iw = (size_t)(d.w_off + e);
    float g = a.g[iw] * a.grad_scale, w = a.w[iw], mw = a.mw[iw], vw = a.vw[iw];
    if (d.s_off >= 0) {
      const size_t is = (size_t)(d.s_off + e);
      float sv = a.s[is];
      const float m = sc::mask_value(a.mode, sv, a.mode == SC_MASK_UNIFORM ? a.u[is] : 0.f, ph, (size_t)e, stream);
      float dw, ds;
      sc::mask_grad_elem(a.mode, g, w, sv, m, a.bypass, spc, dw, ds);
      if (a.update_s) {
        float ms = a.ms[is], vs = a.vs[is];
        a.s[is] = adam_update(sv, ds, ms, vs, lr_s, a.eps_s, 0.f, a.b1, a.b2, a.clip, bc1, bc2);
        a.ms[is] = ms; a.vs[is] = vs;
      }
      g = dw;
    }
    a.w[iw] = adam_update(w, g, mw, vw, lr_w, a.eps_w, a.wd_w, a.b1, a.b2, a.clip, bc1, bc2);
    a.mw[iw] = mw; a.vw[iw] = vw;
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
  if (out_dtype == SC_BF16 && outT == nullptr && out != nullptr && cols % 4 == 0 && (((uintptr_t)g | (uintptr_t)h) & 15) == 0 &&
      ((uintptr_t)out & 7) == 0 && (colsum_accum == nullptr || ((uintptr_t)colsum_accum & 15) == 0)) {
    dim3 vgrid((cols + 127) / 128, (rows + 63) / 64);
    if (h_dtype == SC_BF16 && h != nullptr)
      sc::launch_pdl_aux(prep_grad_vec_kernel<__nv_bfloat16>, vgrid, dim3(256), 0, stream, g, (const __nv_bfloat16*)h, (__nv_bfloat16*)out, rows, cols, scale, dropout_p, seed, stream_id, colsum_accum);
    else
      sc::launch_pdl_aux(prep_grad_vec_kernel<float>, vgrid, dim3(256), 0, stream, g, (const float*)h, (__nv_bfloat16*)out, rows, cols, scale, dropout_p, seed, stream_id, colsum_accum);
    SC_LAUNCH_CHECK("sc_prep_grad");
    return SC_OK;
  }
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

int sc_apply_mask_batched(const void* descs, int n_desc, long total_tiles, int mask_mode, unsigned long long seed,
                          unsigned long long stream_base, int out_dtype, cudaStream_t stream) {
  SC_CHECK(descs != nullptr && n_desc > 0 && total_tiles > 0 && total_tiles < (1L << 31), SC_ERR_SHAPE,
           "sc_apply_mask_batched: n_desc=%d total_tiles=%ld", n_desc, total_tiles);
  static_assert(sizeof(MaskDesc) == 80, "descriptor = 10 x 64-bit words");
  if (out_dtype == SC_BF16)
    apply_mask_batched_kernel<__nv_bfloat16><<<(unsigned)total_tiles, 256, 0, stream>>>((const MaskDesc*)descs, n_desc, mask_mode, seed, stream_base);
  else if (out_dtype == SC_F32)
    apply_mask_batched_kernel<float><<<(unsigned)total_tiles, 256, 0, stream>>>((const MaskDesc*)descs, n_desc, mask_mode, seed, stream_base);
  else SC_CHECK(false, SC_ERR_DTYPE, "sc_apply_mask_batched: bad dtype");
  SC_LAUNCH_CHECK("sc_apply_mask_batched");
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
    if (dtype == SC_BF16 && cols % 8 == 0 && ((uintptr_t)x & 15) == 0 && ((uintptr_t)out & 15) == 0) {
      const int vx = (cols + 255) / 256;
      int vy = (3 * 148 + vx - 1) / vx;
      int vr = ((rows + vy - 1) / vy + 31) / 32 * 32;
      vy = (rows + vr - 1) / vr;
      colsum_bf16_vec_kernel<<<dim3(vx, vy), 256, 0, stream>>>((const __nv_bfloat16*)x, out, rows, cols, vr);
      SC_LAUNCH_CHECK("sc_colsum");
      return SC_OK;
    }
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

int sc_layernorm_bwd_fused(const float* x, const float* a, const void* dy, int dy_dtype, const float* dres, float* dx, float* da,
                           float* db, int rows, int D, float eps, void* next_gb, float* next_colsum, float next_dropout_p,
                           unsigned long long seed, unsigned long long stream_id, cudaStream_t stream);

int sc_layernorm_bwd(const float* x, const float* a, const void* dy, int dy_dtype, const float* dres, float* dx, float* da,
                     float* db, int rows, int D, float eps, cudaStream_t stream) {
  return sc_layernorm_bwd_fused(x, a, dy, dy_dtype, dres, dx, da, db, rows, D, eps, nullptr, nullptr, 0.f, 0, 0, stream);
}

// LayerNorm backward that also prepares the gradient operand of the next linear in the backward chain:
// next_gb (bf16 [rows, D]) = dx (.) dropout keep mask (Philox(seed, stream_id, element), p = next_dropout_p) and
// next_colsum (fp32 [D], accumulated) += its column sums.  Served for D == 512; next_gb == NULL: plain backward.
int sc_layernorm_bwd_fused(const float* x, const float* a, const void* dy, int dy_dtype, const float* dres, float* dx, float* da,
                           float* db, int rows, int D, float eps, void* next_gb, float* next_colsum, float next_dropout_p,
                           unsigned long long seed, unsigned long long stream_id, cudaStream_t stream) {
  SC_CHECK(rows > 0 && D > 1 && D <= 2048, SC_ERR_SHAPE, "sc_layernorm_bwd: rows=%d D=%d", rows, D);
  LnbNext nx;
  nx.gb = (__nv_bfloat16*)next_gb; nx.colsum = next_colsum; nx.p = next_dropout_p; nx.seed = seed; nx.stream = stream_id;
  SC_CHECK(next_gb == nullptr || (D == 512 && ((uintptr_t)next_gb & 7) == 0 && ((uintptr_t)next_colsum & 15) == 0), SC_ERR_UNSUPPORTED,
           "sc_layernorm_bwd_fused: the fused gradient preparation needs D == 512 and aligned outputs");
  if (D == 512 && ((((uintptr_t)x | (uintptr_t)a | (uintptr_t)dy | (uintptr_t)dres | (uintptr_t)dx | (uintptr_t)da | (uintptr_t)db) & 15) == 0) &&
      (dy_dtype == SC_F32 || dy_dtype == SC_BF16)) {
    const int nb = (rows + 7) / 8;
    if (dy_dtype == SC_F32) sc::launch_pdl_aux(layernorm_bwd512_kernel<float>, dim3(nb), dim3(256), 0, stream, x, a, (const float*)dy, dres, dx, da, db, rows, eps, nx);
    else sc::launch_pdl_aux(layernorm_bwd512_kernel<__nv_bfloat16>, dim3(nb), dim3(256), 0, stream, x, a, (const __nv_bfloat16*)dy, dres, dx, da, db, rows, eps, nx);
    SC_LAUNCH_CHECK("sc_layernorm_bwd");
    return SC_OK;
  }
  SC_CHECK(next_gb == nullptr, SC_ERR_UNSUPPORTED, "sc_layernorm_bwd_fused: unaligned operands");
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

int sc_logsoftmax_bwd(const float* logprobs, const float* dy, float* dx, int rows, int V, cudaStream_t stream) {
  SC_CHECK(rows > 0 && V > 0 && logprobs && dy && dx, SC_ERR_SHAPE, "sc_logsoftmax_bwd: rows=%d V=%d", rows, V);
  logsoftmax_bwd_kernel<<<rows, 256, 0, stream>>>(logprobs, dy, dx, V);
  SC_LAUNCH_CHECK("sc_logsoftmax_bwd");
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

int sc_adam_clip_st_chunk(void) { return kStChunk; }

int sc_adam_clip_st(const void* descs, int n_desc, long total_blocks, float* w, const float* grad_wm, float* m_w, float* v_w, float* s,
                    float* m_s, float* v_s, const float* uniforms, int mask_mode, int bypass_sigmoid_grad, int update_logits,
                    unsigned long long seed, unsigned long long stream_base, float lr_w, float eps_w, float weight_decay_w, float lr_s,
                    float eps_s, float beta1, float beta2, float clip_value, float grad_scale, int step,
                    const float* sigmoid_grad_coeff, const float* dyn, cudaStream_t stream) {
  SC_CHECK(descs != nullptr && n_desc > 0 && total_blocks > 0 && total_blocks < (1L << 31) && step >= 1, SC_ERR_SHAPE,
           "sc_adam_clip_st: n_desc=%d blocks=%ld step=%d", n_desc, total_blocks, step);
  SC_CHECK(mask_mode != SC_MASK_UNIFORM || uniforms != nullptr, SC_ERR_SHAPE, "sc_adam_clip_st: uniforms missing");
  static_assert(sizeof(StDesc) == 40, "descriptor = 5 x 64-bit words");
  StArgs a;
  a.w = w; a.g = grad_wm; a.mw = m_w; a.vw = v_w; a.s = s; a.ms = m_s; a.vs = v_s; a.u = uniforms;
  a.mode = mask_mode; a.bypass = bypass_sigmoid_grad; a.update_s = update_logits; a.seed = seed; a.stream_base = stream_base;
  a.lr_w = lr_w; a.eps_w = eps_w; a.wd_w = weight_decay_w; a.lr_s = lr_s; a.eps_s = eps_s; a.b1 = beta1; a.b2 = beta2;
  a.clip = clip_value; a.grad_scale = grad_scale;
  a.bc1 = 1.f - powf(beta1, (float)step); a.bc2_sqrt = sqrtf(1.f - powf(beta2, (float)step));
  a.sig_coeff = sigmoid_grad_coeff; a.dyn = dyn;
  adam_clip_st_kernel<<<(unsigned)total_blocks, 256, 0, stream>>>((const StDesc*)descs, n_desc, a);
  SC_LAUNCH_CHECK("sc_adam_clip_st");
  return SC_OK;
}

}  // extern "C"
