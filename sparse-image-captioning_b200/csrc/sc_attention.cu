// Training-time attention (teacher forcing): forward that keeps the probabilities, and the matching backward.
// One pair of kernels serves the three attentions of the model:
//   decoder self-attention  (causal + key padding)      sparse_caption/models/transformer.py:230-295, relation_transformer.py:356-361
//   decoder cross-attention (S captions x T positions share the K/V of their image: "encoder run once")  transformer.py:255-256
//   encoder box attention   (additive log-geometry bias) sparse_caption/models/relation_transformer.py:258-293
// plus the geometry-bias kernels (BoxRelationalEmbedding + WG + ReLU + log, relation_transformer.py:179-183,196-256)
// with the gradient to the WG weights/biases.
//
// A "group" g owns Tq query rows [g*Tq, (g+1)*Tq) and Tk key rows [g*Tk, (g+1)*Tk).  CTA = (group, head); K,V (and in
// the backward Q, dO, P, dS) live in shared memory; lane = key for score-shaped work, lane = feature for PV-shaped
// work; all reductions are warp shuffles; the backward needs no atomics (phase A per query row, phase B per key row).
#include "sc_common.cuh"

namespace {

constexpr int kMaxPass = 4;  // Tk <= 128

struct AttnArgs {
  const void* q; const void* k; const void* v; int ldq, ldk, ldv;
  const float* key_valid;  // [G, Tk] 0 = masked, or nullptr
  const float* bias;       // [G, h, Tq, Tk] additive, or nullptr
  float* probs;            // [G, h, Tq, Tk] softmax output (pre-dropout), saved for the backward
  void* out; int ldo;
  int G, Tq, Tk, h, dk, causal_T;
  float dropout_p; unsigned long long seed, stream;
  // backward
  const float* d_out; int ldd;  // [G*Tq, ldd] fp32
  float* dq; float* dkk; float* dv; int ldgq, ldgk, ldgv;
  float* dbias;                  // [G, h, Tq, Tk] or nullptr
};

using sc::keep_scale;

// Stage `rows` rows of `dk` elements (global row stride ld, 16-byte aligned) into fp32 shared memory with row stride kst.
// 16-byte loads, four in flight per thread: the scalar version of this loop was latency-bound (ncu: >40 % of the
// forward kernel's stall samples were long-scoreboard waits on 2-byte loads).
template <typename T>
__device__ __forceinline__ void stage_f32(float* dst, int kst, const T* src, size_t ld, int rows, int dk, int tid, int nthr) {
  constexpr int kVec = 16 / (int)sizeof(T);
  const int vpr = dk / kVec;
  const int total = rows * vpr;
  if ((dk % kVec) != 0 || (ld % kVec) != 0 || (((uintptr_t)src) & 15) != 0) {
    for (int e = tid; e < rows * dk; e += nthr) {
      const int j = e / dk, d = e - j * dk;
      dst[j * kst + d] = sc::to_f32<T>(src[(size_t)j * ld + d]);
    }
    return;
  }
  for (int e0 = 0; e0 < total; e0 += 4 * nthr) {
    uint4 buf[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int e = e0 + u * nthr + tid;
      if (e < total) {
        const int j = e / vpr, c = e - j * vpr;
        buf[u] = *(const uint4*)(src + (size_t)j * ld + c * kVec);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int e = e0 + u * nthr + tid;
      if (e < total) {
        const int j = e / vpr, c = e - j * vpr;
        float* o = dst + j * kst + c * kVec;
        if (sizeof(T) == 2) {
          const uint32_t w[4] = {buf[u].x, buf[u].y, buf[u].z, buf[u].w};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            o[2 * i] = __uint_as_float(w[i] << 16);
            o[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
          }
        } else {
          o[0] = __uint_as_float(buf[u].x); o[1] = __uint_as_float(buf[u].y);
          o[2] = __uint_as_float(buf[u].z); o[3] = __uint_as_float(buf[u].w);
        }
      }
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(128) attn_fwd_kernel(const AttnArgs a) {
  extern __shared__ float sm[];
  const int g = blockIdx.x, hh = blockIdx.y;
  const int Tq = a.Tq, Tk = a.Tk, dk = a.dk, kst = dk + 1;
  float* sK = sm;
  float* sV = sK + Tk * kst;
  float* sQ = sV + Tk * kst;       // [Tq][dk]: every query row of the group, staged once
  float* sB = sQ + Tq * dk;        // [Tq][Tk] additive bias (encoder) or unused
  float* sM = sB + (a.bias ? Tq * Tk : 0);  // [Tk] key validity
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const T* qp = (const T*)a.q; const T* kp = (const T*)a.k; const T* vp = (const T*)a.v;
  // all global reads of the CTA are issued up front with 16-byte loads; the row loop below only touches shared memory
  // (ncu on the per-row version: the q / bias loads of every row were exposed long-scoreboard stalls)
  stage_f32<T>(sK, kst, kp + (size_t)g * Tk * a.ldk + hh * dk, a.ldk, Tk, dk, tid, 128);
  stage_f32<T>(sV, kst, vp + (size_t)g * Tk * a.ldv + hh * dk, a.ldv, Tk, dk, tid, 128);
  stage_f32<T>(sQ, dk, qp + (size_t)g * Tq * a.ldq + hh * dk, a.ldq, Tq, dk, tid, 128);
  if (a.bias) {
    const float* bp = a.bias + ((size_t)g * a.h + hh) * Tq * Tk;
    for (int e = tid; e < Tq * Tk; e += 128) sB[e] = bp[e];
  }
  for (int j = tid; j < Tk; j += 128) sM[j] = a.key_valid ? a.key_valid[(size_t)g * Tk + j] : 1.f;
  __syncthreads();
  const float sqrt_dk = sqrtf((float)dk);
  const sc::Philox ph(a.seed);
  for (int i = warp; i < Tq; i += 4) {
    const size_t qrow = (size_t)g * Tq + i;
    const float* sq = sQ + i * dk;
    const size_t pbase = (((size_t)g * a.h + hh) * Tq + i) * Tk;
    const int climit = a.causal_T > 0 ? (i % a.causal_T) : Tk;
    float s_[kMaxPass];
    float mx = -INFINITY;
#pragma unroll
    for (int ps = 0; ps < kMaxPass; ++ps) {
      const int j = ps * 32 + lane;
      float s = -INFINITY;
      if (j < Tk) {
        float dot = 0.f;
        for (int d = 0; d < dk; ++d) dot = fmaf(sq[d], sK[j * kst + d], dot);
        s = dot / sqrt_dk;
        const bool masked = (sM[j] == 0.f) || (a.causal_T > 0 && j > climit);
        if (masked) s = -1e9f;
        if (a.bias) s += sB[i * Tk + j];
      }
      s_[ps] = s;
      mx = fmaxf(mx, s);
    }
    mx = sc::warp_max(mx);
    float sum = 0.f;
#pragma unroll
    for (int ps = 0; ps < kMaxPass; ++ps) {
      const int j = ps * 32 + lane;
      s_[ps] = j < Tk ? expf(s_[ps] - mx) : 0.f;
      sum += s_[ps];
    }
    const float inv = 1.f / sc::warp_sum(sum);
#pragma unroll
    for (int ps = 0; ps < kMaxPass; ++ps) {
      const int j = ps * 32 + lane;
      if (j < Tk) {
        const float p = s_[ps] * inv;
        if (a.probs) a.probs[pbase + j] = p;
        s_[ps] = p * keep_scale(ph, pbase + j, a.stream, a.dropout_p);
      }
    }
    float o0 = 0.f, o1 = 0.f;
#pragma unroll
    for (int ps = 0; ps < kMaxPass; ++ps) {
      if (ps * 32 >= Tk) break;
      const int lim = min(32, Tk - ps * 32);
      for (int l = 0; l < lim; ++l) {
        const float pj = __shfl_sync(0xffffffffu, s_[ps], l);
        const int j = ps * 32 + l;
        if (lane < dk) o0 = fmaf(pj, sV[j * kst + lane], o0);
        if (lane + 32 < dk) o1 = fmaf(pj, sV[j * kst + lane + 32], o1);
      }
    }
    T* op = (T*)a.out;
    if (lane < dk) op[qrow * a.ldo + hh * dk + lane] = sc::from_f32<T>(o0);
    if (lane + 32 < dk) op[qrow * a.ldo + hh * dk + lane + 32] = sc::from_f32<T>(o1);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) attn_bwd_kernel(const AttnArgs a) {
  extern __shared__ float sm[];
  const int g = blockIdx.x, hh = blockIdx.y;
  const int Tq = a.Tq, Tk = a.Tk, dk = a.dk, kst = dk + 1, pst = Tk + 1;
  float* sK = sm;
  float* sV = sK + Tk * kst;
  float* sQ = sV + Tk * kst;
  float* sdO = sQ + Tq * kst;
  float* sPd = sdO + Tq * kst;
  float* sdS = sPd + Tq * pst;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const T* qp = (const T*)a.q; const T* kp = (const T*)a.k; const T* vp = (const T*)a.v;
  stage_f32<T>(sK, kst, kp + (size_t)g * Tk * a.ldk + hh * dk, a.ldk, Tk, dk, tid, 256);
  stage_f32<T>(sV, kst, vp + (size_t)g * Tk * a.ldv + hh * dk, a.ldv, Tk, dk, tid, 256);
  stage_f32<T>(sQ, kst, qp + (size_t)g * Tq * a.ldq + hh * dk, a.ldq, Tq, dk, tid, 256);
  stage_f32<float>(sdO, kst, a.d_out + (size_t)g * Tq * a.ldd + hh * dk, a.ldd, Tq, dk, tid, 256);
  {
    // saved probabilities of the whole (group, head) -> sPd, coalesced, before the row loop
    const float* pp = a.probs + ((size_t)g * a.h + hh) * Tq * Tk;
    for (int e = tid; e < Tq * Tk; e += 256) { const int i = e / Tk, j = e - i * Tk; sPd[i * pst + j] = pp[e]; }
  }
  __syncthreads();
  const float inv_sqrt = 1.f / sqrtf((float)dk);
  const sc::Philox ph(a.seed);
  // phase A: one warp per query row
  for (int i = warp; i < Tq; i += 8) {
    const size_t pbase = (((size_t)g * a.h + hh) * Tq + i) * Tk;
    float P_[kMaxPass], dP_[kMaxPass];
    float delta = 0.f;
#pragma unroll
    for (int ps = 0; ps < kMaxPass; ++ps) {
      const int j = ps * 32 + lane;
      P_[ps] = 0.f; dP_[ps] = 0.f;
      if (j < Tk) {
        float dot = 0.f;
        for (int d = 0; d < dk; ++d) dot = fmaf(sdO[i * kst + d], sV[j * kst + d], dot);
        const float p = sPd[i * pst + j];
        const float m = keep_scale(ph, pbase + j, a.stream, a.dropout_p);
        sPd[i * pst + j] = p * m;
        P_[ps] = p;
        dP_[ps] = dot * m;
        delta += dot * m * p;
      }
    }
    delta = sc::warp_sum(delta);
#pragma unroll
    for (int ps = 0; ps < kMaxPass; ++ps) {
      const int j = ps * 32 + lane;
      if (j < Tk) {
        const float ds = P_[ps] * (dP_[ps] - delta);
        dP_[ps] = ds;
        sdS[i * pst + j] = ds;
        if (a.dbias) a.dbias[pbase + j] = ds;
      }
    }
    float g0 = 0.f, g1 = 0.f;
#pragma unroll
    for (int ps = 0; ps < kMaxPass; ++ps) {
      if (ps * 32 >= Tk) break;
      const int lim = min(32, Tk - ps * 32);
      for (int l = 0; l < lim; ++l) {
        const float dsj = __shfl_sync(0xffffffffu, dP_[ps], l);
        const int j = ps * 32 + l;
        if (lane < dk) g0 = fmaf(dsj, sK[j * kst + lane], g0);
        if (lane + 32 < dk) g1 = fmaf(dsj, sK[j * kst + lane + 32], g1);
      }
    }
    const size_t row = (size_t)g * Tq + i;
    if (lane < dk) a.dq[row * a.ldgq + hh * dk + lane] = g0 * inv_sqrt;
    if (lane + 32 < dk) a.dq[row * a.ldgq + hh * dk + lane + 32] = g1 * inv_sqrt;
  }
  __syncthreads();
  // phase B: one warp per key row
  for (int j = warp; j < Tk; j += 8) {
    float k0 = 0.f, k1 = 0.f, v0 = 0.f, v1 = 0.f;
    for (int i = 0; i < Tq; ++i) {
      const float ds = sdS[i * pst + j], pd = sPd[i * pst + j];
      if (lane < dk) { k0 = fmaf(ds, sQ[i * kst + lane], k0); v0 = fmaf(pd, sdO[i * kst + lane], v0); }
      if (lane + 32 < dk) { k1 = fmaf(ds, sQ[i * kst + lane + 32], k1); v1 = fmaf(pd, sdO[i * kst + lane + 32], v1); }
    }
    const size_t row = (size_t)g * Tk + j;
    if (lane < dk) { a.dkk[row * a.ldgk + hh * dk + lane] = k0 * inv_sqrt; a.dv[row * a.ldgv + hh * dk + lane] = v0; }
    if (lane + 32 < dk) { a.dkk[row * a.ldgk + hh * dk + lane + 32] = k1 * inv_sqrt; a.dv[row * a.ldgv + hh * dk + lane + 32] = v1; }
  }
}

// ---- geometry bias: bias[b,h,i,j] = log(max(relu(WG_h . emb(i,j) + b_h), 1e-6)) ----
constexpr int kMaxHeads = 8;

__device__ __forceinline__ void pair_deltas(const float* boxes, int b, int N, int i, int j, float (&delta)[4]) {
  const float4 bi = *(const float4*)(boxes + ((size_t)b * N + i) * 4);
  const float4 bj = *(const float4*)(boxes + ((size_t)b * N + j) * 4);
  const float cxi = (bi.x + bi.z) * 0.5f, cyi = (bi.y + bi.w) * 0.5f, wi = (bi.z - bi.x) + 1.0f, hi = (bi.w - bi.y) + 1.0f;
  const float cxj = (bj.x + bj.z) * 0.5f, cyj = (bj.y + bj.w) * 0.5f, wj = (bj.z - bj.x) + 1.0f, hj = (bj.w - bj.y) + 1.0f;
  delta[0] = logf(fmaxf(fabsf((cxi - cxj) / wi), 1e-3f));
  delta[1] = logf(fmaxf(fabsf((cyi - cyj) / hi), 1e-3f));
  delta[2] = logf(wi / wj);
  delta[3] = logf(hi / hj);
}

struct DimMat { float v[8]; };

__global__ void __launch_bounds__(256) box_bias_fwd_kernel(const float* __restrict__ boxes, const float* __restrict__ wg_w,
                                                           const float* __restrict__ wg_b, float* __restrict__ bias, int B, int N,
                                                           int h, int trig, DimMat dm) {
  __shared__ float s_wg[kMaxHeads * 64 + kMaxHeads];
  sc::pdl_launch();
  sc::pdl_wait();
  const int dim_g = trig ? 64 : 4;
  for (int i = threadIdx.x; i < h * dim_g; i += 256) s_wg[i] = wg_w[i];
  for (int i = threadIdx.x; i < h; i += 256) s_wg[h * dim_g + i] = wg_b[i];
  __syncthreads();
  const long total = (long)B * N * N;
  for (long p = (long)blockIdx.x * 256 + threadIdx.x; p < total; p += (long)gridDim.x * 256) {
    const int b = (int)(p / (N * N));
    const int r = (int)(p - (long)b * N * N);
    const int i = r / N, j = r - i * N;
    float delta[4];
    pair_deltas(boxes, b, N, i, j, delta);
    float acc[kMaxHeads];
#pragma unroll
    for (int hh = 0; hh < kMaxHeads; ++hh) acc[hh] = 0.f;
    if (trig) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float p100 = 100.0f * delta[c];
#pragma unroll
        for (int f = 0; f < 8; ++f) {
          float sv, cv;
          sincosf(p100 * dm.v[f], &sv, &cv);
#pragma unroll
          for (int hh = 0; hh < kMaxHeads; ++hh)
            if (hh < h) acc[hh] += sv * s_wg[hh * 64 + c * 8 + f] + cv * s_wg[hh * 64 + 32 + c * 8 + f];
        }
      }
    } else {
#pragma unroll
      for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int hh = 0; hh < kMaxHeads; ++hh)
          if (hh < h) acc[hh] += delta[c] * s_wg[hh * 4 + c];
    }
#pragma unroll
    for (int hh = 0; hh < kMaxHeads; ++hh)
      if (hh < h) {
        const float gg = fmaxf(acc[hh] + s_wg[h * dim_g + hh], 0.f);
        bias[(((size_t)b * h + hh) * N + i) * N + j] = logf(fmaxf(gg, 1e-6f));
      }
  }
}

// BoxRelationalEmbedding itself (relation_transformer.py:196-256), for callers of the static method: emb[b,i,j,:] =
// [sin(100 * delta_c / wave^(f/8)) for c in (x,y,w,h), f in 0..7] ++ [cos(same)]  (dim_g = 64), or the 4 deltas (dim_g = 4).
__global__ void __launch_bounds__(256) box_embedding_kernel(const float* __restrict__ boxes, float* __restrict__ emb, int B, int N,
                                                            int trig, DimMat dm) {
  const long total = (long)B * N * N;
  for (long p = (long)blockIdx.x * 256 + threadIdx.x; p < total; p += (long)gridDim.x * 256) {
    const int b = (int)(p / (N * N));
    const int r = (int)(p - (long)b * N * N);
    const int i = r / N, j = r - i * N;
    float delta[4];
    pair_deltas(boxes, b, N, i, j, delta);
    if (trig) {
      float* o = emb + (size_t)p * 64;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float p100 = 100.0f * delta[c];
#pragma unroll
        for (int f = 0; f < 8; ++f) {
          float sv, cv;
          sincosf(p100 * dm.v[f], &sv, &cv);
          o[c * 8 + f] = sv;
          o[32 + c * 8 + f] = cv;
        }
      }
    } else {
      *(float4*)(emb + (size_t)p * 4) = make_float4(delta[0], delta[1], delta[2], delta[3]);
    }
  }
}

// box_attention's additive term from the relu'd geometry weights w_g: log(max(w_g, 1e-6)) (relation_transformer.py:283-286),
// and its gradient d w_g = d bias / w_g where the clamp is inactive.
__global__ void __launch_bounds__(256) log_clamp_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                        float* __restrict__ out, size_t n, float lo) {
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) {
    const float v = x[i];
    out[i] = dy ? (v > lo ? dy[i] / v : 0.f) : logf(fmaxf(v, lo));
  }
}

// dWG[h,f] += sum_pairs dpre * emb_f ; db[h] += sum_pairs dpre, with dpre = dbias * exp(-bias) where the clamp/ReLU
// were inactive (bias > log 1e-6).  CTA per (image, slice of pairs_per_cta pairs): one CTA per image left two thirds
// of the SMs idle at 50 images; 32 pairs of embedding at a time in shared memory.
__global__ void __launch_bounds__(256) box_bias_bwd_kernel(const float* __restrict__ boxes, const float* __restrict__ bias,
                                                           const float* __restrict__ dbias, float* __restrict__ dwg_w,
                                                           float* __restrict__ dwg_b, int B, int N, int h, int trig, DimMat dm,
                                                           int pairs_per_cta) {
  __shared__ float s_emb[32][65];
  __shared__ float s_dpre[kMaxHeads][33];
  sc::pdl_launch();
  sc::pdl_wait();
  const int b = blockIdx.x;
  const int tid = threadIdx.x;
  const int dim_g = trig ? 64 : 4;
  const int hh = tid >> 5, f0 = tid & 31;  // thread owns (head hh, features f0 and f0+32)
  float a0 = 0.f, a1 = 0.f, ab = 0.f;
  const int pairs = N * N;
  const int p_end = min(pairs, ((int)blockIdx.y + 1) * pairs_per_cta);
  for (int p0 = (int)blockIdx.y * pairs_per_cta; p0 < p_end; p0 += 32) {
    __syncthreads();
    {  // 256 threads fill 32 pairs x 8 (c,f-octet) slots: thread -> (pair = tid/8, c = (tid%8)/2, half = tid%2)
      const int pl = tid >> 3, sub = tid & 7;
      const int p = p0 + pl;
      if (p < pairs) {
        const int i = p / N, j = p - i * N;
        float delta[4];
        pair_deltas(boxes, b, N, i, j, delta);
        if (trig) {
          const int c = sub >> 1, fb = (sub & 1) * 4;
          const float p100 = 100.0f * delta[c];
#pragma unroll
          for (int f = 0; f < 4; ++f) {
            float sv, cv;
            sincosf(p100 * dm.v[fb + f], &sv, &cv);
            s_emb[pl][c * 8 + fb + f] = sv;
            s_emb[pl][32 + c * 8 + fb + f] = cv;
          }
        } else if (sub < 4) {
          s_emb[pl][sub] = delta[sub];
        }
      }
      // dpre for (head = tid/32, pair = tid%32)
      const int hd = tid >> 5, pq = tid & 31;
      float dp = 0.f;
      if (hd < h && p0 + pq < pairs) {
        const size_t e = ((size_t)b * h + hd) * pairs + p0 + pq;
        const float bv = bias[e];
        dp = (bv > -13.815510f) ? dbias[e] * expf(-bv) : 0.f;
      }
      s_dpre[hd][pq] = dp;
    }
    __syncthreads();
    if (hh < h) {
      const int lim = min(32, pairs - p0);
      for (int l = 0; l < lim; ++l) {
        const float dp = s_dpre[hh][l];
        if (f0 < dim_g) a0 = fmaf(dp, s_emb[l][f0], a0);
        if (f0 + 32 < dim_g) a1 = fmaf(dp, s_emb[l][f0 + 32], a1);
        if (f0 == 0) ab += dp;
      }
    }
  }
  if (hh < h) {
    if (f0 < dim_g) atomicAdd(&dwg_w[hh * dim_g + f0], a0);
    if (f0 + 32 < dim_g) atomicAdd(&dwg_w[hh * dim_g + f0 + 32], a1);
    if (f0 == 0) atomicAdd(&dwg_b[hh], ab);
  }
}

DimMat make_dim_mat(float wave_len) {
  DimMat d;
  for (int f = 0; f < 8; ++f) d.v[f] = 1.0f / powf(wave_len, (float)f / 8.0f);
  return d;
}

int check_common(const char* name, int G, int Tq, int Tk, int h, int dk) {
  SC_CHECK(G > 0 && Tq > 0 && Tk > 0 && h > 0, SC_ERR_SHAPE, "%s: G=%d Tq=%d Tk=%d h=%d", name, G, Tq, Tk, h);
  SC_CHECK(Tk <= 32 * kMaxPass, SC_ERR_UNSUPPORTED, "%s: Tk=%d > %d", name, Tk, 32 * kMaxPass);
  SC_CHECK(dk >= 1 && dk <= 64, SC_ERR_UNSUPPORTED, "%s: d_k=%d not in [1,64]", name, dk);
  return SC_OK;
}

}  // namespace

// tensor-path kernels (sc_mma_attention_train.cu): SC_ERR_UNSUPPORTED = shape not served, nothing launched
int sc_attn_train_fwd_mma_launch(const void* q, const void* k, const void* v, int ldq, int ldk, int ldv, const float* key_valid,
                                 const float* bias, float* probs, void* out, int ldo, int G, int Tq, int Tk, int h, int dk,
                                 int causal_T, float dropout_p, unsigned long long seed, unsigned long long stream_id,
                                 cudaStream_t stream);
int sc_attn_train_bwd_mma_launch(const void* q, const void* k, const void* v, int ldq, int ldk, int ldv, const float* probs,
                                 const float* d_out, int ldd, float* dq, float* dk_, float* dv, int ldgq, int ldgk, int ldgv,
                                 float* dbias, int G, int Tq, int Tk, int h, int dk, float dropout_p, unsigned long long seed,
                                 unsigned long long stream_id, cudaStream_t stream);

int sc_attn_train_bwd_mma_launch2(const void* q, const void* k, const void* v, int ldq, int ldk, int ldv, const float* probs,
                                  const float* d_out, int ldd, void* dq, void* dk_, void* dv, int out_bf16, int ldgq, int ldgk, int ldgv,
                                  float* bq, float* bk, float* bv, float* dbias, int G, int Tq, int Tk, int h, int dk, float dropout_p,
                                  unsigned long long seed, unsigned long long stream_id, cudaStream_t stream);

extern "C" {

// sc_attention_bwd with the gradient preparation of the following q / k / v projections fused in (bf16, d_k = 64 tensor path
// only; SC_ERR_UNSUPPORTED otherwise, nothing launched): dq / dk / dv are written as bf16 and bq / bk / bv (fp32 [h * d_k],
// accumulated) receive their column sums, i.e. the projections' bias gradients.
int sc_attention_bwd_bf16out(const void* q, const void* k, const void* v, int ldq, int ldk, int ldv, const float* probs,
                             const float* d_out, int ldd, void* dq, void* dk_, void* dv, int ldgq, int ldgk, int ldgv, float* bq,
                             float* bk, float* bv, float* dbias, int G, int Tq, int Tk, int h, int dk, float dropout_p,
                             unsigned long long seed, unsigned long long stream_id, cudaStream_t stream) {
  int rc = check_common("sc_attention_bwd_bf16out", G, Tq, Tk, h, dk);
  if (rc) return rc;
  SC_CHECK(probs != nullptr, SC_ERR_SHAPE, "sc_attention_bwd_bf16out: saved probabilities missing");
  rc = sc_attn_train_bwd_mma_launch2(q, k, v, ldq, ldk, ldv, probs, d_out, ldd, dq, dk_, dv, 1, ldgq, ldgk, ldgv, bq, bk, bv, dbias, G,
                                     Tq, Tk, h, dk, dropout_p, seed, stream_id, stream);
  SC_CHECK(rc != SC_ERR_UNSUPPORTED, SC_ERR_UNSUPPORTED, "sc_attention_bwd_bf16out: shape / alignment not served by the tensor path");
  return rc;
}

int sc_attention_fwd(const void* q, const void* k, const void* v, int ldq, int ldk, int ldv, int dtype, const float* key_valid,
                     const float* bias, float* probs, void* out, int ldo, int G, int Tq, int Tk, int h, int dk, int causal_T,
                     float dropout_p, unsigned long long seed, unsigned long long stream_id, cudaStream_t stream) {
  int rc = check_common("sc_attention_fwd", G, Tq, Tk, h, dk);
  if (rc) return rc;
  if (dtype == SC_BF16) {  // (probs may be NULL: the inference encoder uses the same kernel without saving them)
    rc = sc_attn_train_fwd_mma_launch(q, k, v, ldq, ldk, ldv, key_valid, bias, probs, out, ldo, G, Tq, Tk, h, dk, causal_T,
                                      dropout_p, seed, stream_id, stream);
    if (rc != SC_ERR_UNSUPPORTED) return rc;
  }
  AttnArgs a = {};
  a.q = q; a.k = k; a.v = v; a.ldq = ldq; a.ldk = ldk; a.ldv = ldv; a.key_valid = key_valid; a.bias = bias; a.probs = probs;
  a.out = out; a.ldo = ldo; a.G = G; a.Tq = Tq; a.Tk = Tk; a.h = h; a.dk = dk; a.causal_T = causal_T;
  a.dropout_p = dropout_p; a.seed = seed; a.stream = stream_id;
  const size_t smem = sizeof(float) * (2 * (size_t)Tk * (dk + 1) + (size_t)Tq * dk + (bias ? (size_t)Tq * Tk : 0) + Tk);
  SC_CHECK(smem <= 200 * 1024, SC_ERR_UNSUPPORTED, "sc_attention_fwd: Tq=%d Tk=%d need %zu bytes of shared memory", Tq, Tk, smem);
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(attn_fwd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(attn_fwd_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    attr = true;
  }
  dim3 grid(G, h);
  if (dtype == SC_F32) attn_fwd_kernel<float><<<grid, 128, smem, stream>>>(a);
  else if (dtype == SC_BF16) attn_fwd_kernel<__nv_bfloat16><<<grid, 128, smem, stream>>>(a);
  else SC_CHECK(false, SC_ERR_DTYPE, "sc_attention_fwd: bad dtype %d", dtype);
  SC_LAUNCH_CHECK("sc_attention_fwd");
  return SC_OK;
}

int sc_attention_bwd(const void* q, const void* k, const void* v, int ldq, int ldk, int ldv, int dtype, const float* probs,
                     const float* d_out, int ldd, float* dq, float* dk_, float* dv, int ldgq, int ldgk, int ldgv, float* dbias,
                     int G, int Tq, int Tk, int h, int dk, float dropout_p, unsigned long long seed,
                     unsigned long long stream_id, cudaStream_t stream) {
  int rc = check_common("sc_attention_bwd", G, Tq, Tk, h, dk);
  if (rc) return rc;
  SC_CHECK(probs != nullptr, SC_ERR_SHAPE, "sc_attention_bwd: saved probabilities missing");
  if (dtype == SC_BF16) {
    rc = sc_attn_train_bwd_mma_launch(q, k, v, ldq, ldk, ldv, probs, d_out, ldd, dq, dk_, dv, ldgq, ldgk, ldgv, dbias, G, Tq, Tk,
                                      h, dk, dropout_p, seed, stream_id, stream);
    if (rc != SC_ERR_UNSUPPORTED) return rc;
  }
  AttnArgs a = {};
  a.q = q; a.k = k; a.v = v; a.ldq = ldq; a.ldk = ldk; a.ldv = ldv; a.probs = const_cast<float*>(probs);
  a.G = G; a.Tq = Tq; a.Tk = Tk; a.h = h; a.dk = dk; a.dropout_p = dropout_p; a.seed = seed; a.stream = stream_id;
  a.d_out = d_out; a.ldd = ldd; a.dq = dq; a.dkk = dk_; a.dv = dv; a.ldgq = ldgq; a.ldgk = ldgk; a.ldgv = ldgv; a.dbias = dbias;
  const size_t smem = sizeof(float) * (2 * (size_t)Tk * (dk + 1) + 2 * (size_t)Tq * (dk + 1) + 2 * (size_t)Tq * (Tk + 1));
  SC_CHECK(smem <= 200 * 1024, SC_ERR_UNSUPPORTED, "sc_attention_bwd: Tq=%d Tk=%d need %zu bytes of shared memory", Tq, Tk, smem);
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(attn_bwd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(attn_bwd_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    attr = true;
  }
  dim3 grid(G, h);
  if (dtype == SC_F32) attn_bwd_kernel<float><<<grid, 256, smem, stream>>>(a);
  else if (dtype == SC_BF16) attn_bwd_kernel<__nv_bfloat16><<<grid, 256, smem, stream>>>(a);
  else SC_CHECK(false, SC_ERR_DTYPE, "sc_attention_bwd: bad dtype %d", dtype);
  SC_LAUNCH_CHECK("sc_attention_bwd");
  return SC_OK;
}

int sc_box_bias_fwd(const float* boxes, const float* wg_w, const float* wg_b, float* bias, int B, int N, int h, int trig,
                    float wave_len, cudaStream_t stream) {
  SC_CHECK(B > 0 && N > 0 && h >= 1 && h <= kMaxHeads, SC_ERR_UNSUPPORTED, "sc_box_bias_fwd: B=%d N=%d h=%d", B, N, h);
  SC_CHECK(((uintptr_t)boxes & 15) == 0, SC_ERR_ALIGN, "sc_box_bias_fwd: boxes must be 16-byte aligned");
  long blocks = ((long)B * N * N + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  sc::launch_pdl_aux(box_bias_fwd_kernel, dim3((int)blocks), dim3(256), 0, stream, boxes, wg_w, wg_b, bias, B, N, h, trig, make_dim_mat(wave_len));
  SC_LAUNCH_CHECK("sc_box_bias_fwd");
  return SC_OK;
}

int sc_box_bias_bwd(const float* boxes, const float* bias, const float* dbias, float* dwg_w, float* dwg_b, int B, int N, int h,
                    int trig, float wave_len, cudaStream_t stream) {
  SC_CHECK(B > 0 && N > 0 && h >= 1 && h <= kMaxHeads, SC_ERR_UNSUPPORTED, "sc_box_bias_bwd: B=%d N=%d h=%d", B, N, h);
  // ~3 CTAs per SM; slices are multiples of the 32-pair staging step
  const int pairs = N * N;
  int slices = (3 * 148 + B - 1) / B;
  int ppc = ((pairs + slices - 1) / slices + 31) / 32 * 32;
  slices = (pairs + ppc - 1) / ppc;
  sc::launch_pdl_aux(box_bias_bwd_kernel, dim3(B, slices), dim3(256), 0, stream, boxes, bias, dbias, dwg_w, dwg_b, B, N, h, trig, make_dim_mat(wave_len), ppc);
  SC_LAUNCH_CHECK("sc_box_bias_bwd");
  return SC_OK;
}

int sc_box_embedding(const float* boxes, float* emb, int B, int N, int trig, float wave_len, cudaStream_t stream) {
  SC_CHECK(B > 0 && N > 0, SC_ERR_SHAPE, "sc_box_embedding: B=%d N=%d", B, N);
  SC_CHECK(((uintptr_t)boxes & 15) == 0 && ((uintptr_t)emb & 15) == 0, SC_ERR_ALIGN, "sc_box_embedding: 16-byte alignment");
  long blocks = ((long)B * N * N + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  box_embedding_kernel<<<(int)blocks, 256, 0, stream>>>(boxes, emb, B, N, trig, make_dim_mat(wave_len));
  SC_LAUNCH_CHECK("sc_box_embedding");
  return SC_OK;
}

int sc_log_clamp(const float* x, const float* dy, float* out, size_t n, float lo, cudaStream_t stream) {
  SC_CHECK(n > 0 && x && out, SC_ERR_SHAPE, "sc_log_clamp: bad args");
  size_t blocks = (n + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  log_clamp_kernel<<<(int)blocks, 256, 0, stream>>>(x, dy, out, n, lo);
  SC_LAUNCH_CHECK("sc_log_clamp");
  return SC_OK;
}

}  // extern "C"
