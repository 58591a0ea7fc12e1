// Training-time attention on warp-level tensor cores (bf16 operands, fp32 accumulate): forward that keeps the
// probabilities, and the matching backward.  Serves the three attentions of the teacher-forcing step
//   decoder self-attention  (causal + key padding)   sparse_caption/models/transformer.py:230-295
//   decoder cross-attention (S captions x T positions share the K/V of their image)  transformer.py:255-256
//   encoder box attention   (additive log-geometry bias)  sparse_caption/models/relation_transformer.py:258-293
// for d_k = 64 and key sets of at most 128 entries; sc_attention_fwd / sc_attention_bwd (sc_attention.cu) route here
// and keep their shared-memory FMA kernels for fp32 verification mode and other shapes.
//
// CTA = (group, head).  Q / K / V (and dO in the backward) are staged once as bf16 rows of 144 bytes; the products run
// on mma.sync.m16n8k16 fed by ldmatrix:
//   forward   warp = 16-query m-tile:  S = Q K^T -> mask / bias / softmax in registers -> P (fp32, saved) -> dropout ->
//             O = P V with P re-used as the A fragment straight from the accumulator registers
//   backward  phase A, warp = m-tile:  dP = dO V^T ; dS = P (.) (dP (.) m - delta) ; dQ = dS K ;  dS and P (.) m -> smem
//             phase B, warp = 16-key tile:  dV = (P (.) m)^T dO ; dK = dS^T Q   (A fragments by ldmatrix.trans of the
//             [query][key] tiles written in phase A, so no transposed copy is ever made and no atomics are needed)
// The shared-memory FMA version of these kernels was LSU-bound (ncu: 20 % warp occupancy, 2.7e5 bank conflicts per launch).
#include "sc_common.cuh"

namespace {

constexpr int kDk = 64;
constexpr int kPitch = (kDk + 8) * 2;  // 144-byte rows: the 8 row addresses of an ldmatrix phase hit 8 distinct 16-byte bank groups

struct TrainAttnArgs {
  const __nv_bfloat16* q; const __nv_bfloat16* k; const __nv_bfloat16* v; int ldq, ldk, ldv;
  const float* key_valid;  // [G, Tk] 0 = masked, or nullptr
  const float* bias;       // [G, h, Tq, Tk] additive, or nullptr
  float* probs;            // [G, h, Tq, Tk] softmax output (pre-dropout)
  __nv_bfloat16* out; int ldo;
  int G, Tq, Tk, h, causal_T;
  float dropout_p; unsigned long long seed, stream;
  const float* d_out; int ldd;
  float* dq; float* dkk; float* dv; int ldgq, ldgk, ldgv;
  float* dbias;
  // fused gradient preparation for the q / k / v projections that come next in the backward chain: the gradients leave as
  // bf16 (the GEMM operand dtype) and their column sums (the projections' bias gradients, [h * 64] each) are accumulated
  __nv_bfloat16* dq16; __nv_bfloat16* dk16; __nv_bfloat16* dv16;
  float* bq; float* bk; float* bv;
};

// Column sums of an accumulator-layout tile (rows g / g + 8 of the quad layout, already masked to valid rows and rounded
// to bf16) added into acc[n][c]; after the three shuffles lanes 0..3 (g == 0) hold the sums of columns n * 8 + 2 t + c.
__device__ __forceinline__ void colsum_tile(float (&acc)[kDk / 8][2], const float (&v)[kDk / 8][4]) {
#pragma unroll
  for (int n = 0; n < kDk / 8; ++n)
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      float x = v[n][c] + v[n][2 + c];
      x += __shfl_xor_sync(0xffffffffu, x, 4);
      x += __shfl_xor_sync(0xffffffffu, x, 8);
      x += __shfl_xor_sync(0xffffffffu, x, 16);
      acc[n][c] += x;
    }
}
__device__ __forceinline__ float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *(uint32_t*)&t;
}

// rows [0, rows_pad) of a [*, 64] bf16 slice (row stride ld elements) -> smem rows of kPitch bytes; rows >= n zeroed
__device__ __forceinline__ void stage_bf16(const __nv_bfloat16* __restrict__ src, size_t ld, int n, int rows_pad,
                                           unsigned char* dst, int tid, int nthr) {
  for (int idx = tid; idx < rows_pad * 8; idx += nthr) {
    const int r = idx >> 3, c = idx & 7;
    void* d = dst + r * kPitch + c * 16;
    if (r < n) cp_async16(d, src + (size_t)r * ld + c * 8);
    else *(uint4*)d = make_uint4(0u, 0u, 0u, 0u);
  }
}
// same for an fp32 source (the incoming output gradient), converted to bf16 on the way
__device__ __forceinline__ void stage_f32_as_bf16(const float* __restrict__ src, size_t ld, int n, int rows_pad,
                                                  unsigned char* dst, int tid, int nthr) {
  for (int idx = tid; idx < rows_pad * 8; idx += nthr) {
    const int r = idx >> 3, c = idx & 7;
    uint4 o = make_uint4(0u, 0u, 0u, 0u);
    if (r < n) {
      const float4 x0 = *(const float4*)(src + (size_t)r * ld + c * 8);
      const float4 x1 = *(const float4*)(src + (size_t)r * ld + c * 8 + 4);
      o.x = pack_bf16(x0.x, x0.y); o.y = pack_bf16(x0.z, x0.w); o.z = pack_bf16(x1.x, x1.y); o.w = pack_bf16(x1.z, x1.w);
    }
    *(uint4*)(dst + r * kPitch + c * 16) = o;
  }
}

// acc[2NT][4] = A[m0.., 0..63] * B[0..16NT, 0..63]^T  (both operands row-major [rows][64] in smem: Q K^T, dO V^T)
template <int NT>
__device__ __forceinline__ void mma_abt(float (&acc)[2 * NT][4], const unsigned char* sA, int m0, const unsigned char* sB, int lane) {
#pragma unroll
  for (int n = 0; n < 2 * NT; ++n) { acc[n][0] = 0.f; acc[n][1] = 0.f; acc[n][2] = 0.f; acc[n][3] = 0.f; }
  const uint32_t abase = smem_u32(sA) + (uint32_t)((m0 + (lane & 7) + ((lane >> 3) & 1) * 8) * kPitch + (lane >> 4) * 16);
  const uint32_t bbase = smem_u32(sB) + (uint32_t)(((lane & 7) + (lane >> 4) * 8) * kPitch + ((lane >> 3) & 1) * 16);
#pragma unroll
  for (int ks = 0; ks < kDk / 16; ++ks) {
    uint32_t a[4];
    ldsm_x4(abase + ks * 32, a);
#pragma unroll
    for (int np = 0; np < NT; ++np) {
      uint32_t b[4];
      ldsm_x4(bbase + np * 16 * kPitch + ks * 32, b);
      mma_bf16(acc[2 * np], a, b[0], b[1]);
      mma_bf16(acc[2 * np + 1], a, b[2], b[3]);
    }
  }
}

// o[8][4] = P[16 x 16NT] (accumulator-layout registers, re-packed as A fragments) * B[0..16NT, 0..63]  (P V, dS K)
template <int NT>
__device__ __forceinline__ void mma_pb(float (&o)[kDk / 8][4], const float (&p)[2 * NT][4], const unsigned char* sB, int lane) {
#pragma unroll
  for (int n = 0; n < kDk / 8; ++n) { o[n][0] = 0.f; o[n][1] = 0.f; o[n][2] = 0.f; o[n][3] = 0.f; }
  const uint32_t bbase = smem_u32(sB) + (uint32_t)(((lane & 7) + ((lane >> 3) & 1) * 8) * kPitch + (lane >> 4) * 16);
#pragma unroll
  for (int kk = 0; kk < NT; ++kk) {
    uint32_t a[4];
    a[0] = pack_bf16(p[2 * kk][0], p[2 * kk][1]);
    a[1] = pack_bf16(p[2 * kk][2], p[2 * kk][3]);
    a[2] = pack_bf16(p[2 * kk + 1][0], p[2 * kk + 1][1]);
    a[3] = pack_bf16(p[2 * kk + 1][2], p[2 * kk + 1][3]);
#pragma unroll
    for (int dp = 0; dp < kDk / 16; ++dp) {
      uint32_t b[4];
      ldsm_x4_trans(bbase + kk * 16 * kPitch + dp * 32, b);
      mma_bf16(o[2 * dp], a, b[0], b[1]);
      mma_bf16(o[2 * dp + 1], a, b[2], b[3]);
    }
  }
}

template <int NT>
__global__ void __launch_bounds__(256) attn_train_fwd_mma_kernel(const TrainAttnArgs a) {
  extern __shared__ __align__(16) unsigned char smem_x[];
  sc::pdl_launch();
  sc::pdl_wait();
  const int gi = blockIdx.x, hh = blockIdx.y;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nthr = blockDim.x, nwarps = nthr >> 5;
  const int Tq = a.Tq, Tk = a.Tk;
  const int MT = (Tq + 15) >> 4;
  constexpr int kKeys = 16 * NT;
  unsigned char* sQ = smem_x;
  unsigned char* sK = sQ + (size_t)MT * 16 * kPitch;
  unsigned char* sV = sK + kKeys * kPitch;
  float* sM = (float*)(sV + kKeys * kPitch);  // key validity [kKeys]
  stage_bf16(a.q + (size_t)gi * Tq * a.ldq + hh * kDk, a.ldq, Tq, MT * 16, sQ, tid, nthr);
  stage_bf16(a.k + (size_t)gi * Tk * a.ldk + hh * kDk, a.ldk, Tk, kKeys, sK, tid, nthr);
  stage_bf16(a.v + (size_t)gi * Tk * a.ldv + hh * kDk, a.ldv, Tk, kKeys, sV, tid, nthr);
  for (int j = tid; j < kKeys; j += nthr) sM[j] = (j < Tk && a.key_valid) ? a.key_valid[(size_t)gi * Tk + j] : 1.f;
  cp_async_wait_all();
  __syncthreads();
  const int g = lane >> 2, t = lane & 3;
  const sc::Philox ph(a.seed);
  const size_t pgh = ((size_t)gi * a.h + hh) * Tq;
  for (int mt = warp; mt < MT; mt += nwarps) {
    const int m0 = mt * 16;
    const int r0 = m0 + g, r1 = r0 + 8;
    const bool ok0 = r0 < Tq, ok1 = r1 < Tq;
    const size_t pb0 = (pgh + r0) * Tk, pb1 = (pgh + r1) * Tk;
    // the additive bias of this lane's elements is requested before the tensor-core product
    float bz[2 * NT][4];
    if (a.bias) {
#pragma unroll
      for (int n = 0; n < 2 * NT; ++n)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const int col = n * 8 + 2 * t + c;
          bz[n][c] = (ok0 && col < Tk) ? a.bias[pb0 + col] : 0.f;
          bz[n][2 + c] = (ok1 && col < Tk) ? a.bias[pb1 + col] : 0.f;
        }
    }
    float s[2 * NT][4];
    mma_abt<NT>(s, sQ, m0, sK, lane);
    const int cl0 = a.causal_T > 0 ? (r0 % a.causal_T) : Tk, cl1 = a.causal_T > 0 ? (r1 % a.causal_T) : Tk;
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int n = 0; n < 2 * NT; ++n) {
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int col = n * 8 + 2 * t + c;
        float v0 = s[n][c] * 0.125f, v1 = s[n][2 + c] * 0.125f;  // / sqrt(64)
        if (col < Tk) {
          const bool dead = sM[col] == 0.f;
          if (dead || col > cl0) v0 = -1e9f;
          if (dead || col > cl1) v1 = -1e9f;
          if (a.bias) { v0 += bz[n][c]; v1 += bz[n][2 + c]; }
        } else {
          v0 = -INFINITY; v1 = -INFINITY;
        }
        s[n][c] = v0; s[n][2 + c] = v1;
        mx0 = fmaxf(mx0, v0); mx1 = fmaxf(mx1, v1);
      }
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
    for (int n = 0; n < 2 * NT; ++n) {
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const float p0 = expf(s[n][c] - mx0), p1 = expf(s[n][2 + c] - mx1);
        s[n][c] = p0; s[n][2 + c] = p1;
        sum0 += p0; sum1 += p1;
      }
    }
    sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1); sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
    sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1); sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);
    const float inv0 = 1.f / sum0, inv1 = 1.f / sum1;
#pragma unroll
    for (int n = 0; n < 2 * NT; ++n) {
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int col = n * 8 + 2 * t + c;
        float p0 = s[n][c] * inv0, p1 = s[n][2 + c] * inv1;
        if (col < Tk) {
          if (ok0) { if (a.probs) a.probs[pb0 + col] = p0; p0 *= sc::keep_scale(ph, pb0 + col, a.stream, a.dropout_p); }
          if (ok1) { if (a.probs) a.probs[pb1 + col] = p1; p1 *= sc::keep_scale(ph, pb1 + col, a.stream, a.dropout_p); }
        }
        s[n][c] = p0; s[n][2 + c] = p1;
      }
    }
    float o[kDk / 8][4];
    mma_pb<NT>(o, s, sV, lane);
    // the output tile replaces this warp's own Q rows (no other warp reads them), then leaves in 16-byte rows
    __syncwarp();
#pragma unroll
    for (int n = 0; n < kDk / 8; ++n) {
      *(uint32_t*)(sQ + (m0 + g) * kPitch + (n * 8 + 2 * t) * 2) = pack_bf16(o[n][0], o[n][1]);
      *(uint32_t*)(sQ + (m0 + g + 8) * kPitch + (n * 8 + 2 * t) * 2) = pack_bf16(o[n][2], o[n][3]);
    }
    __syncwarp();
    const int nrows = min(16, Tq - m0);
    __nv_bfloat16* dst = a.out + ((size_t)gi * Tq + m0) * a.ldo + hh * kDk;
    for (int idx = lane; idx < nrows * 8; idx += 32) {
      const int r = idx >> 3, c = idx & 7;
      *(uint4*)(dst + (size_t)r * a.ldo + c * 8) = *(const uint4*)(sQ + (m0 + r) * kPitch + c * 16);
    }
  }
}

template <int NT, bool kOut16>
__global__ void __launch_bounds__(256) attn_train_bwd_mma_kernel(const TrainAttnArgs a) {
  extern __shared__ __align__(16) unsigned char smem_x[];
  sc::pdl_launch();
  sc::pdl_wait();
  const int gi = blockIdx.x, hh = blockIdx.y;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nthr = blockDim.x, nwarps = nthr >> 5;
  const int Tq = a.Tq, Tk = a.Tk;
  const int MT = (Tq + 15) >> 4;
  constexpr int kKeys = 16 * NT;
  constexpr int kPk = 32 * NT + 16;  // pitch of the [query][key] bf16 tiles: an odd multiple of 16 bytes
  unsigned char* sQ = smem_x;
  unsigned char* sdO = sQ + (size_t)MT * 16 * kPitch;
  unsigned char* sK = sdO + (size_t)MT * 16 * kPitch;
  unsigned char* sV = sK + kKeys * kPitch;
  unsigned char* sPD = sV + kKeys * kPitch;               // P (.) dropout mask
  unsigned char* sDS = sPD + (size_t)MT * 16 * kPk;       // dS
  stage_bf16(a.q + (size_t)gi * Tq * a.ldq + hh * kDk, a.ldq, Tq, MT * 16, sQ, tid, nthr);
  stage_bf16(a.k + (size_t)gi * Tk * a.ldk + hh * kDk, a.ldk, Tk, kKeys, sK, tid, nthr);
  stage_bf16(a.v + (size_t)gi * Tk * a.ldv + hh * kDk, a.ldv, Tk, kKeys, sV, tid, nthr);
  stage_f32_as_bf16(a.d_out + (size_t)gi * Tq * a.ldd + hh * kDk, a.ldd, Tq, MT * 16, sdO, tid, nthr);
  cp_async_wait_all();
  __syncthreads();
  const int g = lane >> 2, t = lane & 3;
  const sc::Philox ph(a.seed);
  const size_t pgh = ((size_t)gi * a.h + hh) * Tq;
  float csq[kOut16 ? kDk / 8 : 1][2];  // bias-gradient partial sums of this warp's query tiles (lanes 0..3)
  if constexpr (kOut16) {
#pragma unroll
    for (int n = 0; n < kDk / 8; ++n) { csq[n][0] = 0.f; csq[n][1] = 0.f; }
  }
  // ---- phase A: one warp per 16-query tile ----
  for (int mt = warp; mt < MT; mt += nwarps) {
    const int m0 = mt * 16;
    const int r0 = m0 + g, r1 = r0 + 8;
    const bool ok0 = r0 < Tq, ok1 = r1 < Tq;
    const size_t pb0 = (pgh + r0) * Tk, pb1 = (pgh + r1) * Tk;
    // saved probabilities of this lane's elements: in flight while the tensor cores form dP
    float pr[2 * NT][4];
#pragma unroll
    for (int n = 0; n < 2 * NT; ++n)
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int col = n * 8 + 2 * t + c;
        pr[n][c] = (ok0 && col < Tk) ? a.probs[pb0 + col] : 0.f;
        pr[n][2 + c] = (ok1 && col < Tk) ? a.probs[pb1 + col] : 0.f;
      }
    float s[2 * NT][4];
    mma_abt<NT>(s, sdO, m0, sV, lane);  // dP = dO V^T
    float dl0 = 0.f, dl1 = 0.f;
#pragma unroll
    for (int n = 0; n < 2 * NT; ++n) {
      float pm[4];  // P (.) m of this lane's four elements of the 8-column block
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int col = n * 8 + 2 * t + c;
        float k0 = 0.f, k1 = 0.f;
        if (col < Tk) {
          if (ok0) k0 = sc::keep_scale(ph, pb0 + col, a.stream, a.dropout_p);
          if (ok1) k1 = sc::keep_scale(ph, pb1 + col, a.stream, a.dropout_p);
        }
        const float d0 = s[n][c] * k0, d1 = s[n][2 + c] * k1;  // dP (.) m
        dl0 += d0 * pr[n][c]; dl1 += d1 * pr[n][2 + c];
        s[n][c] = d0; s[n][2 + c] = d1;
        pm[c] = pr[n][c] * k0; pm[2 + c] = pr[n][2 + c] * k1;
      }
      *(uint32_t*)(sPD + (size_t)r0 * kPk + (n * 8 + 2 * t) * 2) = pack_bf16(pm[0], pm[1]);
      *(uint32_t*)(sPD + (size_t)r1 * kPk + (n * 8 + 2 * t) * 2) = pack_bf16(pm[2], pm[3]);
    }
    dl0 += __shfl_xor_sync(0xffffffffu, dl0, 1); dl0 += __shfl_xor_sync(0xffffffffu, dl0, 2);
    dl1 += __shfl_xor_sync(0xffffffffu, dl1, 1); dl1 += __shfl_xor_sync(0xffffffffu, dl1, 2);
#pragma unroll
    for (int n = 0; n < 2 * NT; ++n) {
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int col = n * 8 + 2 * t + c;
        const float e0 = pr[n][c] * (s[n][c] - dl0), e1 = pr[n][2 + c] * (s[n][2 + c] - dl1);  // dS
        s[n][c] = e0; s[n][2 + c] = e1;
        if (a.dbias && col < Tk) {
          if (ok0) a.dbias[pb0 + col] = e0;
          if (ok1) a.dbias[pb1 + col] = e1;
        }
      }
      *(uint32_t*)(sDS + (size_t)r0 * kPk + (n * 8 + 2 * t) * 2) = pack_bf16(s[n][0], s[n][1]);
      *(uint32_t*)(sDS + (size_t)r1 * kPk + (n * 8 + 2 * t) * 2) = pack_bf16(s[n][2], s[n][3]);
    }
    float o[kDk / 8][4];
    mma_pb<NT>(o, s, sK, lane);  // dQ = dS K
    if constexpr (kOut16) {
#pragma unroll
      for (int n = 0; n < kDk / 8; ++n) {
        const int col = hh * kDk + n * 8 + 2 * t;
        o[n][0] = ok0 ? bf16_round(o[n][0] * 0.125f) : 0.f; o[n][1] = ok0 ? bf16_round(o[n][1] * 0.125f) : 0.f;
        o[n][2] = ok1 ? bf16_round(o[n][2] * 0.125f) : 0.f; o[n][3] = ok1 ? bf16_round(o[n][3] * 0.125f) : 0.f;
        if (ok0) *(uint32_t*)(a.dq16 + ((size_t)gi * Tq + r0) * a.ldgq + col) = pack_bf16(o[n][0], o[n][1]);
        if (ok1) *(uint32_t*)(a.dq16 + ((size_t)gi * Tq + r1) * a.ldgq + col) = pack_bf16(o[n][2], o[n][3]);
      }
      if (a.bq) colsum_tile(csq, o);
    } else {
#pragma unroll
      for (int n = 0; n < kDk / 8; ++n) {
        const int col = hh * kDk + n * 8 + 2 * t;
        if (ok0) *(float2*)(a.dq + ((size_t)gi * Tq + r0) * a.ldgq + col) = make_float2(o[n][0] * 0.125f, o[n][1] * 0.125f);
        if (ok1) *(float2*)(a.dq + ((size_t)gi * Tq + r1) * a.ldgq + col) = make_float2(o[n][2] * 0.125f, o[n][3] * 0.125f);
      }
    }
  }
  if constexpr (kOut16) {
    if (a.bq && lane < 4) {
#pragma unroll
      for (int n = 0; n < kDk / 8; ++n)
#pragma unroll
        for (int c = 0; c < 2; ++c)
          if (csq[n][c] != 0.f) atomicAdd(a.bq + hh * kDk + n * 8 + 2 * lane + c, csq[n][c]);
    }
  }
  __syncthreads();
  // ---- phase B: one warp per 16-key tile; the contraction runs over the queries ----
  for (int nt = warp; nt < NT; nt += nwarps) {
    const int j0 = nt * 16;
    float accK[kDk / 8][4], accV[kDk / 8][4];
#pragma unroll
    for (int n = 0; n < kDk / 8; ++n) {
      accK[n][0] = accK[n][1] = accK[n][2] = accK[n][3] = 0.f;
      accV[n][0] = accV[n][1] = accV[n][2] = accV[n][3] = 0.f;
    }
    // A = (tile)^T: matrix mi = lane >> 3 of the x4 load covers stored rows i0 + (mi >> 1) * 8.., columns j0 + (mi & 1) * 8..
    const uint32_t a_off = (uint32_t)(((lane & 7) + ((lane >> 4) & 1) * 8) * kPk + (j0 + ((lane >> 3) & 1) * 8) * 2);
    const uint32_t b_off = (uint32_t)(((lane & 7) + ((lane >> 3) & 1) * 8) * kPitch + (lane >> 4) * 16);
    for (int it = 0; it < MT; ++it) {
      uint32_t apd[4], ads[4];
      ldsm_x4_trans(smem_u32(sPD) + (uint32_t)(it * 16 * kPk) + a_off, apd);
      ldsm_x4_trans(smem_u32(sDS) + (uint32_t)(it * 16 * kPk) + a_off, ads);
#pragma unroll
      for (int dp = 0; dp < kDk / 16; ++dp) {
        uint32_t b[4];
        ldsm_x4_trans(smem_u32(sdO) + (uint32_t)(it * 16 * kPitch + dp * 32) + b_off, b);
        mma_bf16(accV[2 * dp], apd, b[0], b[1]);
        mma_bf16(accV[2 * dp + 1], apd, b[2], b[3]);
        ldsm_x4_trans(smem_u32(sQ) + (uint32_t)(it * 16 * kPitch + dp * 32) + b_off, b);
        mma_bf16(accK[2 * dp], ads, b[0], b[1]);
        mma_bf16(accK[2 * dp + 1], ads, b[2], b[3]);
      }
    }
    const int jr0 = j0 + g, jr1 = jr0 + 8;
    if constexpr (kOut16) {
      const bool k0 = jr0 < Tk, k1 = jr1 < Tk;
#pragma unroll
      for (int n = 0; n < kDk / 8; ++n) {
        const int col = hh * kDk + n * 8 + 2 * t;
        accK[n][0] = k0 ? bf16_round(accK[n][0] * 0.125f) : 0.f; accK[n][1] = k0 ? bf16_round(accK[n][1] * 0.125f) : 0.f;
        accK[n][2] = k1 ? bf16_round(accK[n][2] * 0.125f) : 0.f; accK[n][3] = k1 ? bf16_round(accK[n][3] * 0.125f) : 0.f;
        accV[n][0] = k0 ? bf16_round(accV[n][0]) : 0.f; accV[n][1] = k0 ? bf16_round(accV[n][1]) : 0.f;
        accV[n][2] = k1 ? bf16_round(accV[n][2]) : 0.f; accV[n][3] = k1 ? bf16_round(accV[n][3]) : 0.f;
        if (k0) {
          *(uint32_t*)(a.dk16 + ((size_t)gi * Tk + jr0) * a.ldgk + col) = pack_bf16(accK[n][0], accK[n][1]);
          *(uint32_t*)(a.dv16 + ((size_t)gi * Tk + jr0) * a.ldgv + col) = pack_bf16(accV[n][0], accV[n][1]);
        }
        if (k1) {
          *(uint32_t*)(a.dk16 + ((size_t)gi * Tk + jr1) * a.ldgk + col) = pack_bf16(accK[n][2], accK[n][3]);
          *(uint32_t*)(a.dv16 + ((size_t)gi * Tk + jr1) * a.ldgv + col) = pack_bf16(accV[n][2], accV[n][3]);
        }
      }
      if (a.bk) {
        float ck[kDk / 8][2], cv[kDk / 8][2];
#pragma unroll
        for (int n = 0; n < kDk / 8; ++n) { ck[n][0] = ck[n][1] = cv[n][0] = cv[n][1] = 0.f; }
        colsum_tile(ck, accK);
        colsum_tile(cv, accV);
        if (lane < 4) {
#pragma unroll
          for (int n = 0; n < kDk / 8; ++n)
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              atomicAdd(a.bk + hh * kDk + n * 8 + 2 * lane + c, ck[n][c]);
              atomicAdd(a.bv + hh * kDk + n * 8 + 2 * lane + c, cv[n][c]);
            }
        }
      }
    } else {
#pragma unroll
    for (int n = 0; n < kDk / 8; ++n) {
      const int col = hh * kDk + n * 8 + 2 * t;
      if (jr0 < Tk) {
        *(float2*)(a.dkk + ((size_t)gi * Tk + jr0) * a.ldgk + col) = make_float2(accK[n][0] * 0.125f, accK[n][1] * 0.125f);
        *(float2*)(a.dv + ((size_t)gi * Tk + jr0) * a.ldgv + col) = make_float2(accV[n][0], accV[n][1]);
      }
      if (jr1 < Tk) {
        *(float2*)(a.dkk + ((size_t)gi * Tk + jr1) * a.ldgk + col) = make_float2(accK[n][2] * 0.125f, accK[n][3] * 0.125f);
        *(float2*)(a.dv + ((size_t)gi * Tk + jr1) * a.ldgv + col) = make_float2(accV[n][2], accV[n][3]);
      }
    }
    }
  }
}

bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }

}  // namespace

// Tensor-path entry points used by sc_attention_fwd / sc_attention_bwd (sc_attention.cu).  Return SC_ERR_UNSUPPORTED
// (nothing launched) when the shape is not served; the caller then runs its generic kernel.
int sc_attn_train_fwd_mma_launch(const void* q, const void* k, const void* v, int ldq, int ldk, int ldv, const float* key_valid,
                                 const float* bias, float* probs, void* out, int ldo, int G, int Tq, int Tk, int h, int dk,
                                 int causal_T, float dropout_p, unsigned long long seed, unsigned long long stream_id,
                                 cudaStream_t stream) {
  if (dk != kDk || Tk > 128 || Tq > 1024) return SC_ERR_UNSUPPORTED;
  if (!aligned16(q) || !aligned16(k) || !aligned16(v) || !aligned16(out) || (ldq | ldk | ldv | ldo) % 8 != 0) return SC_ERR_UNSUPPORTED;
  TrainAttnArgs a = {};
  a.q = (const __nv_bfloat16*)q; a.k = (const __nv_bfloat16*)k; a.v = (const __nv_bfloat16*)v; a.ldq = ldq; a.ldk = ldk; a.ldv = ldv;
  a.key_valid = key_valid; a.bias = bias; a.probs = probs; a.out = (__nv_bfloat16*)out; a.ldo = ldo;
  a.G = G; a.Tq = Tq; a.Tk = Tk; a.h = h; a.causal_T = causal_T; a.dropout_p = dropout_p; a.seed = seed; a.stream = stream_id;
  const int MT = (Tq + 15) / 16, NT = (Tk + 15) / 16;
  const size_t smem = (size_t)(MT * 16 + 2 * NT * 16) * kPitch + (size_t)NT * 16 * sizeof(float);
  if (smem > 200 * 1024) return SC_ERR_UNSUPPORTED;
  const int warps = MT < 8 ? MT : 8;
  dim3 grid(G, h);
#define F_CASE(NTV)                                                                                                     \
  case NTV: {                                                                                                           \
    static bool attr = false;                                                                                           \
    if (!attr) {                                                                                                        \
      cudaFuncSetAttribute(attn_train_fwd_mma_kernel<NTV>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);    \
      attr = true;                                                                                                      \
    }                                                                                                                   \
    sc::launch_pdl_aux(attn_train_fwd_mma_kernel<NTV>, grid, dim3(32 * warps), smem, stream, a);                                              \
  } break
  switch (NT) {
    F_CASE(1); F_CASE(2); F_CASE(3); F_CASE(4); F_CASE(5); F_CASE(6); F_CASE(7); F_CASE(8);
    default: return SC_ERR_UNSUPPORTED;
  }
#undef F_CASE
  SC_LAUNCH_CHECK("sc_attention_fwd(mma)");
  return SC_OK;
}

int sc_attn_train_bwd_mma_launch2(const void* q, const void* k, const void* v, int ldq, int ldk, int ldv, const float* probs,
                                  const float* d_out, int ldd, void* dq, void* dk_, void* dv, int out_bf16, int ldgq, int ldgk, int ldgv,
                                  float* bq, float* bk, float* bv, float* dbias, int G, int Tq, int Tk, int h, int dk, float dropout_p,
                                  unsigned long long seed, unsigned long long stream_id, cudaStream_t stream);

int sc_attn_train_bwd_mma_launch(const void* q, const void* k, const void* v, int ldq, int ldk, int ldv, const float* probs,
                                 const float* d_out, int ldd, float* dq, float* dk_, float* dv, int ldgq, int ldgk, int ldgv,
                                 float* dbias, int G, int Tq, int Tk, int h, int dk, float dropout_p, unsigned long long seed,
                                 unsigned long long stream_id, cudaStream_t stream) {
  return sc_attn_train_bwd_mma_launch2(q, k, v, ldq, ldk, ldv, probs, d_out, ldd, dq, dk_, dv, 0, ldgq, ldgk, ldgv, nullptr, nullptr,
                                       nullptr, dbias, G, Tq, Tk, h, dk, dropout_p, seed, stream_id, stream);
}

// out_bf16 != 0: dq / dk / dv are bf16 buffers (the operand dtype of the projection GEMMs that follow in the backward chain)
// and bq / bk / bv (fp32 [h * 64], accumulated, may be NULL) receive their column sums = the projections' bias gradients.
int sc_attn_train_bwd_mma_launch2(const void* q, const void* k, const void* v, int ldq, int ldk, int ldv, const float* probs,
                                  const float* d_out, int ldd, void* dq_, void* dk_, void* dv_, int out_bf16, int ldgq, int ldgk, int ldgv,
                                  float* bq, float* bk, float* bv, float* dbias, int G, int Tq, int Tk, int h, int dk, float dropout_p,
                                  unsigned long long seed, unsigned long long stream_id, cudaStream_t stream) {
  float* dq = out_bf16 ? nullptr : (float*)dq_;
  float* dv = out_bf16 ? nullptr : (float*)dv_;
  if (dk != kDk || Tk > 128 || Tq > 1024) return SC_ERR_UNSUPPORTED;
  if (!aligned16(q) || !aligned16(k) || !aligned16(v) || !aligned16(d_out) || (ldq | ldk | ldv) % 8 != 0 || ldd % 4 != 0) return SC_ERR_UNSUPPORTED;
  if (((uintptr_t)dq_ | (uintptr_t)dk_ | (uintptr_t)dv_) % 8 != 0 || (ldgq | ldgk | ldgv) % 2 != 0) return SC_ERR_UNSUPPORTED;
  if (out_bf16 && ((bq == nullptr) != (bk == nullptr) || (bk == nullptr) != (bv == nullptr))) return SC_ERR_UNSUPPORTED;
  TrainAttnArgs a = {};
  if (out_bf16) { a.dq16 = (__nv_bfloat16*)dq_; a.dk16 = (__nv_bfloat16*)dk_; a.dv16 = (__nv_bfloat16*)dv_; a.bq = bq; a.bk = bk; a.bv = bv; }
  a.q = (const __nv_bfloat16*)q; a.k = (const __nv_bfloat16*)k; a.v = (const __nv_bfloat16*)v; a.ldq = ldq; a.ldk = ldk; a.ldv = ldv;
  a.probs = const_cast<float*>(probs); a.G = G; a.Tq = Tq; a.Tk = Tk; a.h = h; a.dropout_p = dropout_p; a.seed = seed; a.stream = stream_id;
  a.d_out = d_out; a.ldd = ldd; a.dq = dq; a.dkk = out_bf16 ? nullptr : (float*)dk_; a.dv = dv; a.ldgq = ldgq; a.ldgk = ldgk; a.ldgv = ldgv; a.dbias = dbias;
  const int MT = (Tq + 15) / 16, NT = (Tk + 15) / 16;
  const size_t smem = (size_t)(2 * MT * 16 + 2 * NT * 16) * kPitch + (size_t)2 * MT * 16 * (32 * NT + 16);
  if (smem > 200 * 1024) return SC_ERR_UNSUPPORTED;
  int warps = MT > NT ? MT : NT;
  if (warps > 8) warps = 8;
  dim3 grid(G, h);
#define B_CASE(NTV)                                                                                                     \
  case NTV: {                                                                                                           \
    static bool attr = false;                                                                                           \
    if (!attr) {                                                                                                        \
      cudaFuncSetAttribute(attn_train_bwd_mma_kernel<NTV, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);    \
      cudaFuncSetAttribute(attn_train_bwd_mma_kernel<NTV, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);     \
      attr = true;                                                                                                      \
    }                                                                                                                   \
    if (out_bf16) sc::launch_pdl_aux(attn_train_bwd_mma_kernel<NTV, true>, grid, dim3(32 * warps), smem, stream, a);            \
    else sc::launch_pdl_aux(attn_train_bwd_mma_kernel<NTV, false>, grid, dim3(32 * warps), smem, stream, a);                    \
  } break
  switch (NT) {
    B_CASE(1); B_CASE(2); B_CASE(3); B_CASE(4); B_CASE(5); B_CASE(6); B_CASE(7); B_CASE(8);
    default: return SC_ERR_UNSUPPORTED;
  }
#undef B_CASE
  SC_LAUNCH_CHECK("sc_attention_bwd(mma)");
  return SC_OK;
}
