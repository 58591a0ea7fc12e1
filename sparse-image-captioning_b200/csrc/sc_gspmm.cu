// K3b'' - unstructured-sparse linear on the tensor cores by GATHER: y = act(x W^T + b) + residual for a pruned W [N, K]
// (80 - 99 % zeros; SURVEY.md section 2.3 K3b, north_star (2); reference contract: pruning/prune.py:200-221 stores such weights
// as COO, README.md:89-92 gives their nnz).
//
// Idea: a group of 8 output features owns a list of non-zeros (column c, feature f, value v).  Sixteen of them form ONE
// m16n8k16 MMA step: the A operand (16 activation rows x 16 "k slots") is x[rows, c_0 .. c_15] - gathered by ldmatrix, whose
// eight row pointers per 8x8 tile are simply the addresses of the eight columns in a TRANSPOSED activation slab in shared
// memory - and the B operand (16 k slots x 8 features) has exactly one non-zero per k slot, B[j][f_j] = v_j, built in registers
// from the entry words.  Every non-zero costs 1/16 of (8 ldmatrix.x4 + 8 mma + ~25 scalar instructions) per 128 activation
// rows, i.e. ~2.5 issue slots instead of the ~14 of a scalar FMA loop; what is left is the inherent shared-memory traffic of a
// gather SpMM, nnz x rows x 2 bytes (the smem-bandwidth roofline of this kernel: 148 SMs x 128 B/clk).
//
// CTA = 128 activation rows x (8 warps x G feature groups).  The x slab [128, KC <= 512] is staged transposed ([KC][128 + 8]
// bf16, 272-byte rows: an 8-column ldmatrix phase is conflict-free when the columns differ mod 8 - the host orders the
// entries of a group accordingly).  K > 512 (the feed-forward w_2): K chunks with the accumulators kept across chunks.
#include "sc_common.cuh"

namespace {

constexpr int kRows = 128;                    // activation rows per CTA
constexpr int kKC = 512;                      // K chunk staged at once
constexpr int kStrideB = (kRows + 8) * 2;     // 272 bytes per transposed row
constexpr int kThreads = 256;

struct GsArgs {
  const __nv_bfloat16* x; int ldx;
  const int* grp_ptr;          // [nchunks][ngroups + 1] entry offsets (multiples of 16)
  const unsigned* entries;     // (col_in_chunk << 19) | (feature_in_group << 16) | bf16 bits
  const float* bias; const float* residual;
  void* y; int y_bf16, relu;
  int M, N, K, ngroups, nchunks, G;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ldsm_x4_trans(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// x[r0 .. r0+127, k0 .. k0+kc) -> xT[k][row] (bf16), in four passes of 32 rows through two staging buffers:
//   A  coalesced 16-byte cp.async of the rows (row-major, padded rows: (kc * 2 + 16) bytes);
//   B  ldmatrix.x4.trans over four 8-row x 8-column tiles hands every thread the words {x[2t][k], x[2t+1][k]} - exactly the
//      32-bit words of the transposed slab - stored with 32 distinct banks per instruction.
// The loads of pass i + 2 are in flight while pass i is transposed.
constexpr int kPassRows = 32;
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void ldsm_x4_trans_raw(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr) : "memory");
}

__device__ __forceinline__ void stage_issue(const GsArgs& a, unsigned char* S, int stride_s, int rbase, int k0, int kc, int tid) {
  const int cpr = kc >> 3;  // 16-byte chunks per row
  for (int ch = tid; ch < kPassRows * cpr; ch += kThreads) {
    const int row = ch / cpr, c = ch - row * cpr;
    unsigned char* dst = S + (size_t)row * stride_s + c * 16;
    if (rbase + row < a.M) cp_async16(smem_u32(dst), a.x + (size_t)(rbase + row) * a.ldx + k0 + 8 * c);
    else *(uint4*)dst = make_uint4(0u, 0u, 0u, 0u);
  }
  cp_async_commit();
}

__device__ __forceinline__ void stage_transpose(unsigned char* xT, const unsigned char* S, int stride_s, int pass, int kc, int warp, int lane) {
  const int g = lane >> 2, t = lane & 3;
  const uint32_t src = smem_u32(S) + (uint32_t)(((lane >> 3) * 8 + (lane & 7)) * stride_s);  // matrix i = lane / 8: rows 8i .. 8i+7
  for (int ko = warp; ko < (kc >> 3); ko += kThreads / 32) {
    uint32_t w[4];
    ldsm_x4_trans_raw(src + ko * 16, w);
    unsigned char* dst = xT + (size_t)(ko * 8 + g) * kStrideB + (size_t)(pass * kPassRows + 2 * t) * 2;
#pragma unroll
    for (int i = 0; i < 4; ++i) *(uint32_t*)(dst + i * 16) = w[i];   // rows 8i + 2t, 8i + 2t + 1 of this pass
  }
}

__device__ __forceinline__ void stage_slab(const GsArgs& a, unsigned char* xT, unsigned char* stg, int r0, int k0, int kc, int tid) {
  const int stride_s = kc * 2 + 16;
  unsigned char* S[2] = {stg, stg + (size_t)kPassRows * (kKC * 2 + 16)};
  const int warp = tid >> 5, lane = tid & 31;
  constexpr int kPasses = kRows / kPassRows;
  stage_issue(a, S[0], stride_s, r0, k0, kc, tid);
  stage_issue(a, S[1], stride_s, r0 + kPassRows, k0, kc, tid);
#pragma unroll
  for (int pass = 0; pass < kPasses; ++pass) {
    if (pass + 1 < kPasses) cp_async_wait<1>(); else cp_async_wait<0>();
    __syncthreads();
    stage_transpose(xT, S[pass & 1], stride_s, pass, kc, warp, lane);
    __syncthreads();
    if (pass + 2 < kPasses) stage_issue(a, S[pass & 1], stride_s, r0 + (pass + 2) * kPassRows, k0, kc, tid);
  }
}

// One MMA step: 16 entries held by lanes src0 .. src0 + 15 of `blk`.
__device__ __forceinline__ void mma_step(const unsigned char* xT, unsigned blk, int src0, int lane, float (&acc)[8][4]) {
  const int g = lane >> 2, t = lane & 3;
  const int ei = src0 + ((lane >> 4) << 3) + (lane & 7);   // entry whose column this lane addresses for ldmatrix
  const uint32_t half_off = ((lane >> 3) & 1) * 16;         // rows 0-7 / 8-15 of the m-tile
  const unsigned ea = __shfl_sync(0xffffffffu, blk, ei);
  const uint32_t addr = smem_u32(xT) + (ea >> 19) * kStrideB + half_off;
  // B fragment: b0/b1 = k slots 2t, 2t+1; b2/b3 = k slots 2t+8, 2t+9; column n = g
  const unsigned e00 = __shfl_sync(0xffffffffu, blk, src0 + 2 * t), e01 = __shfl_sync(0xffffffffu, blk, src0 + 2 * t + 1);
  const unsigned e10 = __shfl_sync(0xffffffffu, blk, src0 + 2 * t + 8), e11 = __shfl_sync(0xffffffffu, blk, src0 + 2 * t + 9);
  const uint32_t v00 = (((e00 >> 16) & 7u) == (unsigned)g) ? (e00 & 0xffffu) : 0u;
  const uint32_t v01 = (((e01 >> 16) & 7u) == (unsigned)g) ? (e01 << 16) : 0u;
  const uint32_t v10 = (((e10 >> 16) & 7u) == (unsigned)g) ? (e10 & 0xffffu) : 0u;
  const uint32_t v11 = (((e11 >> 16) & 7u) == (unsigned)g) ? (e11 << 16) : 0u;
  const uint32_t b0 = v00 | v01, b1 = v10 | v11;
#pragma unroll
  for (int mt = 0; mt < 8; ++mt) {
    uint32_t af[4];
    ldsm_x4_trans(addr + mt * 32, af);
    mma_bf16(acc[mt], af, b0, b1);
  }
}

constexpr int kAhead = 4;  // 32-entry blocks in flight per warp (8 MMA steps: more than one L2 round trip of compute)

__device__ __forceinline__ unsigned load_block(const GsArgs& a, int e0, int e1, int b, int lane) {
  const int idx = e0 + b * 32 + lane;
  return idx < e1 ? __ldg(a.entries + idx) : 0u;   // (a zero word: column 0, value 0)
}

// `buf`: the first kAhead blocks of the group, loaded by the caller (possibly before the slab was staged)
__device__ __forceinline__ void process_group(const GsArgs& a, const unsigned char* xT, int e0, int e1, int lane, unsigned (&buf)[kAhead],
                                              float (&acc)[8][4]) {
  const int steps = (e1 - e0) >> 4;
  const int nblk = (steps + 1) >> 1;
  for (int b0 = 0; b0 < nblk; b0 += kAhead) {
#pragma unroll
    for (int j = 0; j < kAhead; ++j) {
      const int b = b0 + j;
      if (b < nblk) {
        const unsigned cur = buf[j];
        buf[j] = load_block(a, e0, e1, b + kAhead, lane);
        mma_step(xT, cur, 0, lane, acc);
        if (2 * b + 1 < steps) mma_step(xT, cur, 16, lane, acc);
      }
    }
  }
}

__device__ __forceinline__ void store_group(const GsArgs& a, int r0, int grp, int lane, float (&acc)[8][4]) {
  const int g = lane >> 2, t = lane & 3;
  const int n = grp * 8 + 2 * t;
  if (n >= a.N) return;
  const bool two = n + 1 < a.N;
  const float bz0 = a.bias ? a.bias[n] : 0.f, bz1 = (a.bias && two) ? a.bias[n + 1] : 0.f;
#pragma unroll
  for (int mt = 0; mt < 8; ++mt) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int m = r0 + mt * 16 + g + 8 * h;
      if (m >= a.M) continue;
      float o0 = acc[mt][2 * h] + bz0, o1 = acc[mt][2 * h + 1] + bz1;
      if (a.relu) { o0 = fmaxf(o0, 0.f); o1 = fmaxf(o1, 0.f); }
      const size_t idx = (size_t)m * a.N + n;
      if (a.residual) { o0 += a.residual[idx]; if (two) o1 += a.residual[idx + 1]; }
      if (a.y_bf16) {
        __nv_bfloat16* y = (__nv_bfloat16*)a.y;
        if (two && (idx & 1) == 0) *(__nv_bfloat162*)(y + idx) = __floats2bfloat162_rn(o0, o1);
        else { y[idx] = __float2bfloat16_rn(o0); if (two) y[idx + 1] = __float2bfloat16_rn(o1); }
      } else {
        float* y = (float*)a.y;
        if (two && (idx & 1) == 0) *(float2*)(y + idx) = make_float2(o0, o1);
        else { y[idx] = o0; if (two) y[idx + 1] = o1; }
      }
    }
  }
}

__global__ void __launch_bounds__(kThreads, 1) gspmm_kernel(const GsArgs a) {
  extern __shared__ __align__(16) unsigned char xT[];
  unsigned char* stg = xT + (size_t)kKC * kStrideB;   // two row-major staging buffers behind the transposed slab
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int r0 = blockIdx.x * kRows;
  const int grp0 = (blockIdx.y * 8 + warp) * a.G;   // this warp's first feature group
  sc::pdl_launch();
  sc::pdl_wait();
  float acc[8][4];
  if (a.nchunks == 1) {
    // the entry words of the first group are requested before the slab is staged: both latencies overlap
    unsigned buf[kAhead];
    int e0 = 0, e1 = 0;
    if (grp0 < a.ngroups) { e0 = a.grp_ptr[grp0]; e1 = a.grp_ptr[grp0 + 1]; }
#pragma unroll
    for (int j = 0; j < kAhead; ++j) buf[j] = load_block(a, e0, e1, j, lane);
    stage_slab(a, xT, stg, r0, 0, a.K, tid);
    for (int gi = 0; gi < a.G; ++gi) {
      const int grp = grp0 + gi;
      if (grp >= a.ngroups) break;
#pragma unroll
      for (int mt = 0; mt < 8; ++mt) { acc[mt][0] = 0.f; acc[mt][1] = 0.f; acc[mt][2] = 0.f; acc[mt][3] = 0.f; }
      process_group(a, xT, e0, e1, lane, buf, acc);
      // next group's first blocks are in flight while this group's tile is written out
      const int ne0 = e1, ne1 = (grp + 1 < a.ngroups && gi + 1 < a.G) ? a.grp_ptr[grp + 2] : e1;
#pragma unroll
      for (int j = 0; j < kAhead; ++j) buf[j] = load_block(a, ne0, ne1, j, lane);
      store_group(a, r0, grp, lane, acc);
      e0 = ne0; e1 = ne1;
    }
    return;
  }
  // K chunks (G == 1): the warp's accumulators live across the chunks; the slab is re-staged between block-wide barriers
  const int grp = grp0;
#pragma unroll
  for (int mt = 0; mt < 8; ++mt) { acc[mt][0] = 0.f; acc[mt][1] = 0.f; acc[mt][2] = 0.f; acc[mt][3] = 0.f; }
  for (int c = 0; c < a.nchunks; ++c) {
    const int k0 = c * kKC;
    unsigned buf[kAhead];
    int e0 = 0, e1 = 0;
    if (grp < a.ngroups) {
      const int* gp = a.grp_ptr + (size_t)c * (a.ngroups + 1);
      e0 = gp[grp]; e1 = gp[grp + 1];
    }
#pragma unroll
    for (int j = 0; j < kAhead; ++j) buf[j] = load_block(a, e0, e1, j, lane);
    if (c > 0) __syncthreads();   // every warp is done reading the previous chunk
    stage_slab(a, xT, stg, r0, k0, min(kKC, a.K - k0), tid);
    if (grp < a.ngroups) process_group(a, xT, e0, e1, lane, buf, acc);
  }
  if (grp < a.ngroups) store_group(a, r0, grp, lane, acc);
}

}  // namespace

extern "C" int sc_gspmm(const void* x, int ldx, const int* grp_ptr, const void* entries, const float* bias, const float* residual,
                        void* y, int y_dtype, int M, int N, int K, int relu, cudaStream_t stream) {
  SC_CHECK(M > 0 && N > 0 && K > 0 && K % 8 == 0 && ldx >= K && ldx % 8 == 0, SC_ERR_SHAPE, "sc_gspmm: M=%d N=%d K=%d ldx=%d (K, ldx multiples of 8)", M, N, K, ldx);
  SC_CHECK(y_dtype == SC_F32 || y_dtype == SC_BF16, SC_ERR_DTYPE, "sc_gspmm: bad y dtype");
  SC_CHECK(((uintptr_t)x & 15) == 0 && ((uintptr_t)entries & 3) == 0 && ((uintptr_t)y & 7) == 0, SC_ERR_ALIGN, "sc_gspmm: x must be 16-byte aligned");
  GsArgs a;
  a.x = (const __nv_bfloat16*)x; a.ldx = ldx; a.grp_ptr = grp_ptr; a.entries = (const unsigned*)entries; a.bias = bias; a.residual = residual;
  a.y = y; a.y_bf16 = (y_dtype == SC_BF16); a.relu = relu; a.M = M; a.N = N; a.K = K;
  a.ngroups = (N + 7) / 8;
  a.nchunks = (K + kKC - 1) / kKC;
  const int row_blocks = (M + kRows - 1) / kRows;
  int G = 1;
  if (a.nchunks == 1) {
    // ~1.5 CTAs per SM at most: fewer feature blocks = fewer re-stagings of the same x slab
    G = (int)(((long)a.ngroups * row_blocks) / (8L * 222));
    G = G < 1 ? 1 : (G > 8 ? 8 : G);
  }
  a.G = G;
  const int fblocks = (a.ngroups + 8 * G - 1) / (8 * G);
  const size_t smem = (size_t)kKC * kStrideB + 2 * (size_t)kPassRows * (kKC * 2 + 16);
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(gspmm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    SC_CHECK(e == cudaSuccess, (int)e, "sc_gspmm: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr = true;
  }
  cudaError_t e = sc::launch_pdl(gspmm_kernel, dim3(row_blocks, fblocks), dim3(kThreads), smem, stream, a);
  SC_CHECK(e == cudaSuccess, (int)e, "sc_gspmm: launch failed: %s", cudaGetErrorString(e));
  SC_LAUNCH_CHECK("sc_gspmm");
  return SC_OK;
}
