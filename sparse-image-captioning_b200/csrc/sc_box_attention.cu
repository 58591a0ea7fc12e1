// K4: BoxRelationalEmbedding + WG projection + ReLU + log + QK^T/sqrt(dk) + mask + softmax + PV in ONE kernel.
// Reference: sparse_caption/models/relation_transformer.py:148-191 (BoxMultiHeadedAttention.forward),
// :196-256 (BoxRelationalEmbedding), :258-293 (box_attention).
//
// The reference materialises emb[B,N,N,64] fp32 (and 8 x [B*N*N,64] GEMVs) per encoder layer; here a CTA owns
// (image, query block): phase 1 computes the log-geometry bias of every (query,key) pair for all heads straight
// from the 4 box coordinates (sincosf with full range reduction: angles reach +-690 rad) into shared memory;
// phase 2 walks the heads: K_h/V_h staged in shared memory as fp32, one warp per query row, lane = key for the
// scores, warp-shuffle max/sum for the softmax, lane = output dim for PV.  Nothing but Q,K,V,O touches HBM.
#include "sc_common.cuh"

namespace {

constexpr int kMaxHeads = 8;
constexpr int kMaxKeyIters = 4;  // N <= 128

struct BoxArgs {
  const void* q; const void* k; const void* v;  // row (b*N+i), head h, dim d at [row*ld + h*dk + d]
  int ldq, ldk, ldv;
  const float* boxes;     // [B,N,4] x_min,y_min,x_max,y_max
  const float* wg_w;      // [h, dim_g]  (already masked)
  const float* wg_b;      // [h]
  const float* att_mask;  // [B,N] (0 = padded) or nullptr
  void* out; int ldo;
  int B, N, h, dk, QB, trig;
  float dim_mat[8];
};

template <typename T>
__global__ void __launch_bounds__(256) box_attention_kernel(const BoxArgs a) {
  extern __shared__ float sm[];
  const int N = a.N, h = a.h, dk = a.dk, QB = a.QB;
  const int b = blockIdx.x;
  const int q0 = blockIdx.y * QB;
  const int nq = min(QB, N - q0);
  const int dim_g = a.trig ? 64 : 4;
  const int kst = dk + 1;  // odd row stride: lane=key reads are bank-conflict free
  float* s_bias = sm;                         // [h][QB][N]
  float* s_k = s_bias + h * QB * N;           // [N][dk+1]
  float* s_v = s_k + N * kst;                 // [N][dk+1]
  float* s_q = s_v + N * kst;                 // [8 warps][dk]
  float* s_geo = s_q + 8 * dk;                // [N][4] cx, cy, w, h
  float* s_wg = s_geo + N * 4;                // [h][dim_g] + [h]
  float* s_mask = s_wg + h * dim_g + h;       // [N]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  for (int j = tid; j < N; j += 256) {
    const float4 bx = *(const float4*)(a.boxes + ((size_t)b * N + j) * 4);
    s_geo[j * 4 + 0] = (bx.x + bx.z) * 0.5f;
    s_geo[j * 4 + 1] = (bx.y + bx.w) * 0.5f;
    s_geo[j * 4 + 2] = (bx.z - bx.x) + 1.0f;
    s_geo[j * 4 + 3] = (bx.w - bx.y) + 1.0f;
    s_mask[j] = a.att_mask ? a.att_mask[(size_t)b * N + j] : 1.f;
  }
  for (int i = tid; i < h * dim_g; i += 256) s_wg[i] = a.wg_w[i];
  for (int i = tid; i < h; i += 256) s_wg[h * dim_g + i] = a.wg_b[i];
  __syncthreads();

  // ---- phase 1: log(max(relu(WG_h . emb(i,j) + b_h), 1e-6)) for every pair, all heads ----
  for (int p = tid; p < nq * N; p += 256) {
    const int il = p / N, j = p - il * N;
    const int i = q0 + il;
    const float cxi = s_geo[i * 4], cyi = s_geo[i * 4 + 1], wi = s_geo[i * 4 + 2], hi = s_geo[i * 4 + 3];
    const float cxj = s_geo[j * 4], cyj = s_geo[j * 4 + 1], wj = s_geo[j * 4 + 2], hj = s_geo[j * 4 + 3];
    float delta[4];
    delta[0] = logf(fmaxf(fabsf((cxi - cxj) / wi), 1e-3f));
    delta[1] = logf(fmaxf(fabsf((cyi - cyj) / hi), 1e-3f));
    delta[2] = logf(wi / wj);
    delta[3] = logf(hi / hj);
    float acc[kMaxHeads];
#pragma unroll
    for (int hh = 0; hh < kMaxHeads; ++hh) acc[hh] = 0.f;
    if (a.trig) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float p100 = 100.0f * delta[c];
#pragma unroll
        for (int f = 0; f < 8; ++f) {
          float sv, cv;
          sincosf(p100 * a.dim_mat[f], &sv, &cv);
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
        const float g = fmaxf(acc[hh] + s_wg[h * dim_g + hh], 0.f);
        s_bias[(hh * QB + il) * N + j] = logf(fmaxf(g, 1e-6f));
      }
  }

  // ---- phase 2: per head softmax(bias + QK^T/sqrt(dk)) V ----
  const float sqrt_dk = sqrtf((float)dk);
  const T* qp = (const T*)a.q; const T* kp = (const T*)a.k; const T* vp = (const T*)a.v;
  T* op = (T*)a.out;
  for (int hh = 0; hh < h; ++hh) {
    __syncthreads();  // bias ready (first pass) / previous head's K,V no longer read
    for (int e = tid; e < N * dk; e += 256) {
      const int j = e / dk, d = e - j * dk;
      const size_t row = (size_t)b * N + j;
      s_k[j * kst + d] = sc::to_f32<T>(kp[row * a.ldk + hh * dk + d]);
      s_v[j * kst + d] = sc::to_f32<T>(vp[row * a.ldv + hh * dk + d]);
    }
    __syncthreads();
    for (int il = warp; il < nq; il += 8) {
      const size_t qrow = (size_t)b * N + q0 + il;
      for (int d = lane; d < dk; d += 32) s_q[warp * dk + d] = sc::to_f32<T>(qp[qrow * a.ldq + hh * dk + d]);
      __syncwarp();
      float sc_[kMaxKeyIters];
      float mx = -INFINITY;
#pragma unroll
      for (int jj = 0; jj < kMaxKeyIters; ++jj) {
        const int j = lane + jj * 32;
        float s = -INFINITY;
        if (j < N) {
          float dot = 0.f;
          for (int d = 0; d < dk; ++d) dot = fmaf(s_q[warp * dk + d], s_k[j * kst + d], dot);
          s = dot / sqrt_dk;
          if (s_mask[j] == 0.f) s = -1e9f;
          s += s_bias[(hh * QB + il) * N + j];
        }
        sc_[jj] = s;
        mx = fmaxf(mx, s);
      }
      mx = sc::warp_max(mx);
      float sum = 0.f;
#pragma unroll
      for (int jj = 0; jj < kMaxKeyIters; ++jj) {
        const int j = lane + jj * 32;
        sc_[jj] = j < N ? expf(sc_[jj] - mx) : 0.f;
        sum += sc_[jj];
      }
      sum = sc::warp_sum(sum);
      const float inv = 1.f / sum;
      float o0 = 0.f, o1 = 0.f;  // dims lane, lane+32 (dk <= 64)
#pragma unroll
      for (int jj = 0; jj < kMaxKeyIters; ++jj) {
        if (jj * 32 >= N) break;
        const int lim = min(32, N - jj * 32);
        for (int l = 0; l < lim; ++l) {
          const float pj = __shfl_sync(0xffffffffu, sc_[jj], l);
          const int j = jj * 32 + l;
          if (lane < dk) o0 = fmaf(pj, s_v[j * kst + lane], o0);
          if (lane + 32 < dk) o1 = fmaf(pj, s_v[j * kst + lane + 32], o1);
        }
      }
      if (lane < dk) op[qrow * a.ldo + hh * dk + lane] = sc::from_f32<T>(o0 * inv);
      if (lane + 32 < dk) op[qrow * a.ldo + hh * dk + lane + 32] = sc::from_f32<T>(o1 * inv);
      __syncwarp();
    }
  }
}

size_t box_smem_bytes(int N, int h, int dk, int QB, int dim_g) {
  return sizeof(float) * ((size_t)h * QB * N + 2 * (size_t)N * (dk + 1) + 8 * dk + (size_t)N * 4 + h * dim_g + h + N);
}

}  // namespace

extern "C" int sc_box_attention_fwd(const void* q, const void* k, const void* v, int ldq, int ldk, int ldv, int dtype,
                                    const float* boxes, const float* wg_w, const float* wg_b, const float* att_mask,
                                    void* out, int ldo, int B, int N, int h, int dk, int trig, float wave_len,
                                    cudaStream_t stream) {
  SC_CHECK(B > 0 && N > 0, SC_ERR_SHAPE, "sc_box_attention_fwd: B=%d N=%d", B, N);
  SC_CHECK(N <= 32 * kMaxKeyIters, SC_ERR_UNSUPPORTED, "sc_box_attention_fwd: N=%d > %d boxes", N, 32 * kMaxKeyIters);
  SC_CHECK(h >= 1 && h <= kMaxHeads, SC_ERR_UNSUPPORTED, "sc_box_attention_fwd: heads=%d not in [1,%d]", h, kMaxHeads);
  SC_CHECK(dk >= 1 && dk <= 64, SC_ERR_UNSUPPORTED, "sc_box_attention_fwd: d_k=%d not in [1,64]", dk);
  SC_CHECK(((uintptr_t)boxes & 15) == 0, SC_ERR_ALIGN, "sc_box_attention_fwd: boxes must be 16-byte aligned");
  const int dim_g = trig ? 64 : 4;
  int QB = N;
  const size_t budget = 200 * 1024;
  while (QB > 1 && box_smem_bytes(N, h, dk, QB, dim_g) > budget) QB = (QB + 1) / 2;
  const size_t smem = box_smem_bytes(N, h, dk, QB, dim_g);
  SC_CHECK(smem <= budget, SC_ERR_UNSUPPORTED, "sc_box_attention_fwd: shared memory %zu too large", smem);
  BoxArgs a;
  a.q = q; a.k = k; a.v = v; a.ldq = ldq; a.ldk = ldk; a.ldv = ldv;
  a.boxes = boxes; a.wg_w = wg_w; a.wg_b = wg_b; a.att_mask = att_mask; a.out = out; a.ldo = ldo;
  a.B = B; a.N = N; a.h = h; a.dk = dk; a.QB = QB; a.trig = trig;
  // dim_mat = 1 / wave_len^(f/8) in fp32, as torch computes it (relation_transformer.py:236-238)
  for (int f = 0; f < 8; ++f) a.dim_mat[f] = 1.0f / powf(wave_len, (float)f / 8.0f);
  dim3 grid(B, (N + QB - 1) / QB);
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(box_attention_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)budget);
    SC_CHECK(e == cudaSuccess, (int)e, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    e = cudaFuncSetAttribute(box_attention_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)budget);
    SC_CHECK(e == cudaSuccess, (int)e, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_done = true;
  }
  if (dtype == SC_F32) {
    box_attention_kernel<float><<<grid, 256, smem, stream>>>(a);
  } else if (dtype == SC_BF16) {
    box_attention_kernel<__nv_bfloat16><<<grid, 256, smem, stream>>>(a);
  } else {
    SC_CHECK(false, SC_ERR_DTYPE, "sc_box_attention_fwd: bad dtype %d", dtype);
  }
  SC_LAUNCH_CHECK("sc_box_attention_fwd");
  return SC_OK;
}
