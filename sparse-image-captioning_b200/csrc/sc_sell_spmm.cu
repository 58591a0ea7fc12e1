// K3b': Y = X * W^T + bias for 80-99 % unstructured-sparse W in SLICED-ELL form (slab = 32 output features).
// Inference-side product for binarized-mask weights (the reference only *stores* its pruned weights, as COO:
// sparse_caption/pruning/prune.py:200-221, utils/model_utils.py:110-118, and multiplies them densely).
//
// Layout: slab s holds output features n = 32 s .. 32 s + 31; its entries are stored entry-major, feature-minor:
// entry i of feature 32 s + l sits at slab_ptr[s] + 32 i + l, so a warp (lane = feature) reads 32 entries with ONE
// coalesced 128-byte load.  bf16: entry = (column << 16) | bf16(value) (4 bytes); fp32 verification mode: entry =
// {column, fp32 value} (8 bytes).  Rows shorter than the slab's widest row are padded with (column 0, value 0); widths
// are multiples of 4 so the entry loop is unrolled without guards.
//
// CTA = 8 activation rows x a range of slabs.  The 8 rows are staged once in shared memory as fp32 [K][8]: one
// sparse entry then costs a lane two 16-byte shared loads and 8 FMAs (the CSR kernel K3b spends two shuffles and four
// 2-byte loads on 4 FMAs and re-stages a 128-row tile through 32-way bank conflicts).  Outputs are written with
// lane = feature, i.e. 128-byte rows.  No atomics, deterministic.
#include "sc_common.cuh"

namespace {

constexpr int kRows = 8;

struct SellArgs {
  const void* x;             // [M,K] bf16 or fp32
  const int* slab_ptr;       // [slabs + 1], in entries
  const void* entries;       // packed u32 (bf16) or uint2 (fp32)
  const float* bias; const float* residual; void* y; int y_bf16; int relu;
  int M, N, K, slabs, slabs_per_cta;
};

template <typename T>
__global__ void __launch_bounds__(256) sell_spmm_kernel(const SellArgs a) {
  extern __shared__ __align__(16) float xs[];  // [K][kRows]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int m0 = blockIdx.x * kRows;
  sc::pdl_wait();
  {
    // stage: a thread reads 16 bytes of one row and scatters them down the k axis
    constexpr int kVec = 16 / (int)sizeof(T);
    const int vpr = a.K / kVec;  // K % kVec == 0 (checked on the host)
    const T* x = (const T*)a.x;
    for (int idx = tid; idx < kRows * vpr; idx += 256) {
      const int r = idx % kRows, c = idx / kRows;  // consecutive threads: consecutive rows -> consecutive smem words
      float v[8];
      if (m0 + r < a.M) {
        const uint4 raw = *(const uint4*)(x + (size_t)(m0 + r) * a.K + c * kVec);
        if (sizeof(T) == 2) {
          const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
          for (int i = 0; i < 4; ++i) { v[2 * i] = __uint_as_float(w[i] << 16); v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u); }
        } else {
          v[0] = __uint_as_float(raw.x); v[1] = __uint_as_float(raw.y); v[2] = __uint_as_float(raw.z); v[3] = __uint_as_float(raw.w);
        }
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = 0.f;
      }
#pragma unroll
      for (int i = 0; i < kVec; ++i) xs[(size_t)(c * kVec + i) * kRows + r] = v[i];
    }
  }
  __syncthreads();
  sc::pdl_launch();
  const int s_begin = blockIdx.y * a.slabs_per_cta;
  const int s_end = min(a.slabs, s_begin + a.slabs_per_cta);
  for (int s = s_begin + warp; s < s_end; s += 8) {
    const int base = a.slab_ptr[s];
    const int width = (a.slab_ptr[s + 1] - base) >> 5;  // entries per feature, multiple of 4
    float acc[kRows];
#pragma unroll
    for (int r = 0; r < kRows; ++r) acc[r] = 0.f;
    if (sizeof(T) == 2) {
      const uint32_t* e = (const uint32_t*)a.entries + base + lane;
      for (int i = 0; i < width; i += 4) {
        uint32_t en[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) en[u] = __ldg(e + (size_t)(i + u) * 32);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float v = __uint_as_float(en[u] << 16);
          const float4* xp = (const float4*)(xs + (size_t)(en[u] >> 16) * kRows);
          const float4 x0 = xp[0], x1 = xp[1];
          acc[0] = fmaf(v, x0.x, acc[0]); acc[1] = fmaf(v, x0.y, acc[1]); acc[2] = fmaf(v, x0.z, acc[2]); acc[3] = fmaf(v, x0.w, acc[3]);
          acc[4] = fmaf(v, x1.x, acc[4]); acc[5] = fmaf(v, x1.y, acc[5]); acc[6] = fmaf(v, x1.z, acc[6]); acc[7] = fmaf(v, x1.w, acc[7]);
        }
      }
    } else {
      const uint2* e = (const uint2*)a.entries + base + lane;
      for (int i = 0; i < width; i += 4) {
        uint2 en[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) en[u] = __ldg(e + (size_t)(i + u) * 32);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float v = __uint_as_float(en[u].y);
          const float4* xp = (const float4*)(xs + (size_t)en[u].x * kRows);
          const float4 x0 = xp[0], x1 = xp[1];
          acc[0] = fmaf(v, x0.x, acc[0]); acc[1] = fmaf(v, x0.y, acc[1]); acc[2] = fmaf(v, x0.z, acc[2]); acc[3] = fmaf(v, x0.w, acc[3]);
          acc[4] = fmaf(v, x1.x, acc[4]); acc[5] = fmaf(v, x1.y, acc[5]); acc[6] = fmaf(v, x1.z, acc[6]); acc[7] = fmaf(v, x1.w, acc[7]);
        }
      }
    }
    const int n = s * 32 + lane;
    if (n < a.N) {
      const float bz = a.bias ? __ldg(a.bias + n) : 0.f;
#pragma unroll
      for (int r = 0; r < kRows; ++r) {
        const int m = m0 + r;
        if (m >= a.M) break;
        float o = acc[r] + bz;
        if (a.relu) o = fmaxf(o, 0.f);
        if (a.residual) o += a.residual[(size_t)m * a.N + n];  // plain load: written by the predecessor
        if (a.y_bf16) ((__nv_bfloat16*)a.y)[(size_t)m * a.N + n] = __float2bfloat16_rn(o);
        else ((float*)a.y)[(size_t)m * a.N + n] = o;
      }
    }
  }
}

}  // namespace

extern "C" int sc_sell_spmm(const void* x, int dtype, const int* slab_ptr, const void* entries, const float* bias,
                            const float* residual, void* y, int y_dtype, int M, int N, int K, int relu, cudaStream_t stream) {
  SC_CHECK(M > 0 && N > 0 && K > 0, SC_ERR_SHAPE, "sc_sell_spmm: M=%d N=%d K=%d", M, N, K);
  SC_CHECK(dtype == SC_BF16 || dtype == SC_F32, SC_ERR_DTYPE, "sc_sell_spmm: bad x dtype %d", dtype);
  SC_CHECK(y_dtype == SC_F32 || y_dtype == SC_BF16, SC_ERR_DTYPE, "sc_sell_spmm: bad y dtype");
  SC_CHECK(K <= 65536 && K % (dtype == SC_BF16 ? 8 : 4) == 0, SC_ERR_SHAPE, "sc_sell_spmm: K=%d must be <= 65536 and a multiple of %d", K,
           dtype == SC_BF16 ? 8 : 4);
  SC_CHECK(((uintptr_t)x & 15) == 0 && ((uintptr_t)entries & 7) == 0, SC_ERR_ALIGN, "sc_sell_spmm: x must be 16-byte aligned");
  const size_t smem = (size_t)K * kRows * sizeof(float);
  SC_CHECK(smem <= 200 * 1024, SC_ERR_UNSUPPORTED, "sc_sell_spmm: K=%d needs %zu bytes of shared memory", K, smem);
  SellArgs a;
  a.x = x; a.slab_ptr = slab_ptr; a.entries = entries; a.bias = bias; a.residual = residual; a.y = y;
  a.y_bf16 = (y_dtype == SC_BF16); a.relu = relu; a.M = M; a.N = N; a.K = K;
  a.slabs = (N + 31) / 32;
  const int groups = (M + kRows - 1) / kRows;
  // ~3 CTAs per SM, but at least one slab per warp
  int nsplit = (3 * 148 + groups - 1) / groups;
  const int max_split = a.slabs / 8 > 0 ? a.slabs / 8 : 1;
  if (nsplit > max_split) nsplit = max_split;
  if (nsplit < 1) nsplit = 1;
  a.slabs_per_cta = (a.slabs + nsplit - 1) / nsplit;
  dim3 grid(groups, (a.slabs + a.slabs_per_cta - 1) / a.slabs_per_cta);
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(sell_spmm_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(sell_spmm_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    attr = true;
  }
  cudaError_t e;
  if (dtype == SC_BF16) e = sc::launch_pdl(sell_spmm_kernel<__nv_bfloat16>, grid, dim3(256), smem, stream, a);
  else e = sc::launch_pdl(sell_spmm_kernel<float>, grid, dim3(256), smem, stream, a);
  SC_CHECK(e == cudaSuccess, (int)e, "sc_sell_spmm: launch failed: %s", cudaGetErrorString(e));
  SC_LAUNCH_CHECK("sc_sell_spmm");
  return SC_OK;
}
