// K1 / K3a: Y = X * (W (.) mask)^T + bias  on the 5th-gen tensor cores (tcgen05.mma, TMEM accumulators),
// replacing the reference's  sigmoid -> bernoulli/round -> mul -> F.linear  chain
// (sparse_caption/pruning/masked_layer.py:84-110,134-135; sampler.py:43-66).
//
//   A (activations, bf16 [M,K] row-major)      : TMA (SWIZZLE_128B) -> smem ring
//   B (weights [N,K] row-major), two sources   :
//       kDense  : bf16 weights (already W(.)m, the "densified" inference weights)  -> TMA
//       kMasked : fp32 master weights + fp32 mask logits; 4 transform warps load both with 16-byte
//                 coalesced loads, apply the mask (binarize / Philox-Bernoulli / raw / injected uniforms),
//                 convert to bf16 and write the UMMA operand tile in the 128B-swizzled K-major layout
//                 (the "operand-load prologue"); the masked weight never exists in HBM.
//   D (fp32 [128 x BLOCK_N]) in TMEM; epilogue warps tcgen05.ld it, add bias / residual, ReLU, store.
//
// Warp roles (one CTA per output tile): w0 TMA producer, w1 MMA issuer (one elected lane), w2 TMEM
// allocator, w4-7 epilogue (TMEM lane quarter = warp%4), w8-11 B-transform (kMasked only).
#include "sc_common.cuh"
#include <cuda.h>

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;  // 64 bf16 = 128 B = one swizzle row
constexpr int UMMA_K = 16;
constexpr int kNumTransformWarps = 4;

struct GemmArgs {
  int M, N, K;
  const float* w32;      // kMasked: fp32 weights [N,K]
  const float* mask;     // kMasked: fp32 logits / raw mask / nullptr
  const float* uniforms; // SC_MASK_UNIFORM
  int mask_mode;
  unsigned long long seed, stream_id;
  const float* bias;      // [N] or nullptr
  const float* residual;  // [M,N] fp32 or nullptr
  void* y;
  int y_bf16;
  int relu;
  // forward training: y = dropout(act(acc + bias)) + residual
  float dropout_p; unsigned long long drop_seed, drop_stream;
  // weight-gradient mode: the accumulator tile is dWm[n,k]; the epilogue emits dW and dS (straight-through)
  int wgrad; int bypass; float sp_coeff; int accumulate;
  const float* wg_w; const float* wg_s; const float* wg_u; float* dw; float* ds;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tcgen05_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major, SWIZZLE_128B operand tile: rows of 128 B, 8-row groups 1024 B apart.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);  // start address, 16-byte units
  d |= (uint64_t)1 << 16;                   // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;         // stride byte offset between 8-row groups
  d |= (uint64_t)1 << 46;                   // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                   // SWIZZLE_128B
  return d;
}
// kind::f16, A=B=bf16, D=f32, both K-major
__host__ __device__ constexpr uint32_t make_idesc(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

template <int BLOCK_N, int kStages>
struct Smem {
  static constexpr int kABytes = BLOCK_M * BLOCK_K * 2;
  static constexpr int kBBytes = BLOCK_N * BLOCK_K * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kBarOffset = kStages * kStageBytes;
  static constexpr int kTotal = kBarOffset + 256 + 1024;  // barriers + tmem ptr + alignment slack
};

// kStages: depth of the TMA->MMA ring.  Rings are kept shallow enough (64x4: 96 KB, 128x3: 96 KB) for two CTAs to
// share an SM, so one CTA's epilogue overlaps the other's main loop (scripts/gemm_sweep.py: deeper rings with one
// CTA per SM measured slower for every shape of this model).
template <int BLOCK_N, bool kMasked, int kStages>
__global__ void __launch_bounds__(kMasked ? 384 : 256, 1)
sc_gemm_bf16_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b, const GemmArgs args) {
  using L = Smem<BLOCK_N, kStages>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = (uint64_t*)(smem + L::kBarOffset);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full_bar = empty_bar + kStages;
  uint32_t* tmem_ptr_smem = (uint32_t*)(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * BLOCK_N;
  const int m0 = blockIdx.y * BLOCK_M;
  const int num_kb = (args.K + BLOCK_K - 1) / BLOCK_K;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tma_a) : "memory");
    if (!kMasked) asm volatile("prefetch.tensormap [%0];" ::"l"(&tma_b) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], kMasked ? 1 + kNumTransformWarps : 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)),
                 "r"((uint32_t)BLOCK_N));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % kStages;
        const uint32_t ph = (kb / kStages) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        uint8_t* sa = smem + s * L::kStageBytes;
        mbar_expect_tx(&full_bar[s], kMasked ? L::kABytes : L::kStageBytes);
        tma_load_2d(&tma_a, &full_bar[s], sa, kb * BLOCK_K, m0);
        if (!kMasked) tma_load_2d(&tma_b, &full_bar[s], sa + L::kABytes, kb * BLOCK_K, n0);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(BLOCK_M, BLOCK_N);
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % kStages;
        const uint32_t ph = (kb / kStages) & 1;
        mbar_wait(&full_bar[s], ph);
        tcgen05_fence_after();
        const uint32_t sa = smem_u32(smem + s * L::kStageBytes);
        const uint64_t da = make_smem_desc(sa);
        const uint64_t db = make_smem_desc(sa + L::kABytes);
#pragma unroll
        for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
          // advance 32 B (16 bf16) along K inside the 128B swizzle row: +2 in 16-byte units
          umma_bf16(tmem_base, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kb > 0 || k > 0) ? 1u : 0u);
        }
        tcgen05_commit(&empty_bar[s]);  // frees the smem slot once these MMAs have read it
      }
      tcgen05_commit(tmem_full_bar);  // accumulator complete
    }
  } else if (warp >= 4 && warp < 8) {
    // ===== epilogue: TMEM -> registers -> global =====
    mbar_wait(tmem_full_bar, 0);
    tcgen05_fence_after();
    const int q = warp & 3;  // TMEM lane quarter this warp may read
    const int row = m0 + q * 32 + lane;
    const bool row_ok = row < args.M;
    const bool vec_ok = (args.N % 8) == 0;
#pragma unroll 1
    for (int c = 0; c < BLOCK_N / 32; ++c) {
      uint32_t v[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32), v);
      const int col0 = n0 + c * 32;
      if (!row_ok || col0 >= args.N) continue;
      float f[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
      if (args.wgrad) {
        // K2: dW = dWm (.) m ; dS = dWm (.) W (.) sigmoid'(S) (+ sparsity term), mask regenerated from (seed, stream, element)
        const sc::Philox wph(args.seed);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int col = col0 + j;
          if (col >= args.N) continue;
          const size_t e = (size_t)row * args.N + col;
          const float sv = args.wg_s ? __ldg(args.wg_s + e) : 0.f;
          const float m = sc::mask_value(args.mask_mode, sv, args.wg_u ? __ldg(args.wg_u + e) : 0.f, wph, e, args.stream_id);
          float gw, gs;
          sc::mask_grad_elem(args.mask_mode, f[j], __ldg(args.wg_w + e), sv, m, args.bypass, args.sp_coeff, gw, gs);
          if (args.dw) args.dw[e] = (args.accumulate ? args.dw[e] : 0.f) + gw;
          if (args.ds) args.ds[e] = (args.accumulate ? args.ds[e] : 0.f) + gs;
        }
        continue;
      }
      if (args.dropout_p > 0.f) {
        // training forward epilogue with dropout: scalar path (bias, act, dropout, residual)
        const sc::Philox dph(args.drop_seed);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int col = col0 + j;
          if (col >= args.N) continue;
          const size_t e = (size_t)row * args.N + col;
          float x = f[j];
          if (args.bias) x += __ldg(args.bias + col);
          if (args.relu) x = fmaxf(x, 0.f);
          x *= sc::keep_scale(dph, e, args.drop_stream, args.dropout_p);
          if (args.residual) x += __ldg(args.residual + e);
          if (args.y_bf16) ((__nv_bfloat16*)args.y)[e] = __float2bfloat16_rn(x);
          else ((float*)args.y)[e] = x;
        }
        continue;
      }
      if (vec_ok && col0 + 32 <= args.N) {
        if (args.bias) {
          const float4* bp = (const float4*)(args.bias + col0);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float4 b = __ldg(bp + j);
            f[4 * j] += b.x; f[4 * j + 1] += b.y; f[4 * j + 2] += b.z; f[4 * j + 3] += b.w;
          }
        }
        if (args.relu) {
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
        }
        if (args.residual) {
          const float4* rp = (const float4*)(args.residual + (size_t)row * args.N + col0);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float4 r = __ldg(rp + j);
            f[4 * j] += r.x; f[4 * j + 1] += r.y; f[4 * j + 2] += r.z; f[4 * j + 3] += r.w;
          }
        }
        if (args.y_bf16) {
          uint4* yp = (uint4*)((__nv_bfloat16*)args.y + (size_t)row * args.N + col0);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            __nv_bfloat162 p0 = __floats2bfloat162_rn(f[8 * j], f[8 * j + 1]);
            __nv_bfloat162 p1 = __floats2bfloat162_rn(f[8 * j + 2], f[8 * j + 3]);
            __nv_bfloat162 p2 = __floats2bfloat162_rn(f[8 * j + 4], f[8 * j + 5]);
            __nv_bfloat162 p3 = __floats2bfloat162_rn(f[8 * j + 6], f[8 * j + 7]);
            uint4 o;
            o.x = *(uint32_t*)&p0; o.y = *(uint32_t*)&p1; o.z = *(uint32_t*)&p2; o.w = *(uint32_t*)&p3;
            yp[j] = o;
          }
        } else {
          float4* yp = (float4*)((float*)args.y + (size_t)row * args.N + col0);
#pragma unroll
          for (int j = 0; j < 8; ++j) yp[j] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int col = col0 + j;
          if (col >= args.N) continue;
          float x = f[j];
          if (args.bias) x += __ldg(args.bias + col);
          if (args.relu) x = fmaxf(x, 0.f);
          if (args.residual) x += __ldg(args.residual + (size_t)row * args.N + col);
          if (args.y_bf16) ((__nv_bfloat16*)args.y)[(size_t)row * args.N + col] = __float2bfloat16_rn(x);
          else ((float*)args.y)[(size_t)row * args.N + col] = x;
        }
      }
    }
  } else if (kMasked && warp >= 8) {
    // ===== B transform: fp32 W (+ mask logits) -> masked bf16 operand tile (swizzled K-major) =====
    const int t = threadIdx.x - 256;  // 0..127
    const int chunk = t & 15;         // float4 index inside the 64-wide k block
    const int rbase = t >> 4;         // 0..7
    const sc::Philox philox(args.seed);
    constexpr int kPasses = BLOCK_N / 8;
    for (int kb = 0; kb < num_kb; ++kb) {
      const int s = kb % kStages;
      const uint32_t ph = (kb / kStages) & 1;
      mbar_wait(&empty_bar[s], ph ^ 1);
      uint8_t* sb = smem + s * L::kStageBytes + L::kABytes;
      const int k = kb * BLOCK_K + chunk * 4;
#pragma unroll 4
      for (int p = 0; p < kPasses; ++p) {
        const int r = rbase + p * 8;
        const int n = n0 + r;
        float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
        if (n < args.N && k < args.K) {
          const size_t e = (size_t)n * args.K + k;
          w = __ldg((const float4*)(args.w32 + e));
          if (args.mask_mode != SC_MASK_NONE) {
            const float4 sv = __ldg((const float4*)(args.mask + e));
            float m[4];
            const float sa[4] = {sv.x, sv.y, sv.z, sv.w};
            if (args.mask_mode == SC_MASK_ROUND) {
#pragma unroll
              for (int i = 0; i < 4; ++i) m[i] = sc::mask_round(sa[i]);
            } else if (args.mask_mode == SC_MASK_BERNOULLI) {
              sc::bernoulli4(philox, e >> 2, args.stream_id, sa, m);
            } else if (args.mask_mode == SC_MASK_UNIFORM) {
              const float4 u = __ldg((const float4*)(args.uniforms + e));
              m[0] = u.x < sc::sigmoidf_(sa[0]) ? 1.f : 0.f;
              m[1] = u.y < sc::sigmoidf_(sa[1]) ? 1.f : 0.f;
              m[2] = u.z < sc::sigmoidf_(sa[2]) ? 1.f : 0.f;
              m[3] = u.w < sc::sigmoidf_(sa[3]) ? 1.f : 0.f;
            } else {
#pragma unroll
              for (int i = 0; i < 4; ++i) m[i] = sa[i];
            }
            w.x *= m[0]; w.y *= m[1]; w.z *= m[2]; w.w *= m[3];
          }
        }
        __nv_bfloat162 lo = __floats2bfloat162_rn(w.x, w.y);
        __nv_bfloat162 hi = __floats2bfloat162_rn(w.z, w.w);
        uint2 o;
        o.x = *(uint32_t*)&lo; o.y = *(uint32_t*)&hi;
        // element (r, kk = chunk*4): 16-byte chunk index kk/8 = chunk>>1, XOR-swizzled with r%8
        const uint32_t off = (uint32_t)r * 128u + ((((uint32_t)chunk >> 1) ^ ((uint32_t)r & 7u)) << 4) + (((uint32_t)chunk & 1u) << 3);
        *(uint2*)(sb + off) = o;
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to UMMA
      __syncwarp();
      if (lane == 0) mbar_arrive(&full_bar[s]);
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)BLOCK_N));
  }
}

// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = (EncodeTiledFn)p;
  return fn;
}

// bf16 [rows, cols] row-major, box = [box_rows, 64 cols], SWIZZLE_128B
int make_tmap(CUtensorMap* map, const void* ptr, int rows, int cols, int box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  SC_CHECK(fn != nullptr, SC_ERR_DRIVER, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
  cuuint32_t box[2] = {(cuuint32_t)BLOCK_K, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SC_CHECK(r == CUDA_SUCCESS, SC_ERR_DRIVER, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return SC_OK;
}

template <int BLOCK_N, bool kMasked, int kStages>
int launch(const CUtensorMap& ta, const CUtensorMap& tb, const GemmArgs& a, cudaStream_t stream) {
  auto kern = sc_gemm_bf16_kernel<BLOCK_N, kMasked, kStages>;
  constexpr int smem = Smem<BLOCK_N, kStages>::kTotal;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    SC_CHECK(e == cudaSuccess, (int)e, "cudaFuncSetAttribute(smem=%d): %s", smem, cudaGetErrorString(e));
    attr_set = true;
  }
  dim3 grid((a.N + BLOCK_N - 1) / BLOCK_N, (a.M + BLOCK_M - 1) / BLOCK_M);
  kern<<<grid, kMasked ? 384 : 256, smem, stream>>>(ta, tb, a);
  SC_LAUNCH_CHECK("sc_gemm_bf16_kernel");
  return SC_OK;
}

}  // namespace

// Internal entry used by sc_linear (sc_api.cu).
int sc_gemm_bf16_launch(const void* x, const void* w, int w_dtype, const float* mask, int mask_mode, const float* uniforms,
                        unsigned long long seed, unsigned long long stream_id, const float* bias, const float* residual,
                        void* y, int y_dtype, int M, int N, int K, int relu, int block_n, const ScGemmExtra* ex,
                        cudaStream_t stream) {
  SC_CHECK(M > 0 && N > 0 && K > 0, SC_ERR_SHAPE, "sc_linear: empty problem M=%d N=%d K=%d", M, N, K);
  SC_CHECK(K % 8 == 0, SC_ERR_SHAPE, "sc_linear(bf16): K=%d must be a multiple of 8 (16-byte TMA rows)", K);
  SC_CHECK(((uintptr_t)x & 15) == 0 && ((uintptr_t)w & 15) == 0 && ((uintptr_t)y & 15) == 0, SC_ERR_ALIGN,
           "sc_linear(bf16): x, w, y must be 16-byte aligned");
  SC_CHECK(y_dtype == SC_F32 || y_dtype == SC_BF16, SC_ERR_DTYPE, "sc_linear: bad y dtype %d", y_dtype);
  const bool wgrad = ex && ex->wgrad;
  const bool masked = (w_dtype == SC_F32);
  SC_CHECK(masked || mask_mode == SC_MASK_NONE || wgrad, SC_ERR_DTYPE,
           "sc_linear(bf16): mask modes need fp32 master weights (bf16 weights are expected pre-masked)");
  if (masked && mask_mode != SC_MASK_NONE) {
    SC_CHECK(mask != nullptr && ((uintptr_t)mask & 15) == 0, SC_ERR_ALIGN, "sc_linear: mask must be 16-byte aligned");
    SC_CHECK(mask_mode != SC_MASK_UNIFORM || uniforms != nullptr, SC_ERR_SHAPE, "sc_linear: uniforms missing");
    SC_CHECK(K % 4 == 0, SC_ERR_SHAPE, "K %% 4");
  }
  const long tiles128 = (long)((N + 127) / 128) * ((M + 127) / 128);
  int force_stages = 0;
  if (block_n >= 1000) {  // tuning hook: tile_n = 1000 * stages + block_n
    force_stages = block_n / 1000;
    block_n %= 1000;
  }
  if (block_n == 0) {
    // fill the 148 SMs: prefer the widest tile that still yields >= ~1 wave
    // measured with scripts/gemm_sweep.py (in-graph, L2-warm): 128-wide tiles win once they fill the 148 SMs
    block_n = (tiles128 >= 148) ? 128 : 64;
    if (N <= 64) block_n = 64;
  }
  CUtensorMap ta, tb;
  int rc = make_tmap(&ta, x, M, K, BLOCK_M);
  if (rc) return rc;
  if (!masked) {
    rc = make_tmap(&tb, w, N, K, block_n);
    if (rc) return rc;
  } else {
    tb = ta;
  }
  GemmArgs a;
  a.M = M; a.N = N; a.K = K;
  a.w32 = masked ? (const float*)w : nullptr;
  a.mask = mask; a.uniforms = uniforms; a.mask_mode = mask_mode;
  a.seed = seed; a.stream_id = stream_id;
  a.bias = bias; a.residual = residual; a.y = y; a.y_bf16 = (y_dtype == SC_BF16); a.relu = relu;
  a.dropout_p = 0.f; a.drop_seed = 0; a.drop_stream = 0; a.wgrad = 0; a.bypass = 0; a.sp_coeff = 0.f; a.accumulate = 0;
  a.wg_w = nullptr; a.wg_s = nullptr; a.wg_u = nullptr; a.dw = nullptr; a.ds = nullptr;
  if (ex) {
    a.dropout_p = ex->dropout_p; a.drop_seed = ex->drop_seed; a.drop_stream = ex->drop_stream;
    a.wgrad = ex->wgrad; a.bypass = ex->bypass; a.sp_coeff = ex->sp_coeff; a.accumulate = ex->accumulate;
    a.wg_w = ex->wg_w; a.wg_s = ex->wg_s; a.wg_u = ex->wg_u; a.dw = ex->dw; a.ds = ex->ds;
  }
#define SC_GEMM_CASE(BN, ST) \
  if (block_n == BN && stages == ST) return masked ? launch<BN, true, ST>(ta, tb, a, stream) : launch<BN, false, ST>(ta, tb, a, stream)
  int stages = force_stages;
  if (stages == 0) stages = block_n == 64 ? 4 : block_n == 256 ? 4 : 3;
  SC_GEMM_CASE(64, 4); SC_GEMM_CASE(64, 8);
  SC_GEMM_CASE(128, 3); SC_GEMM_CASE(128, 4); SC_GEMM_CASE(128, 6);
  SC_GEMM_CASE(256, 4);
#undef SC_GEMM_CASE
  SC_CHECK(false, SC_ERR_UNSUPPORTED, "sc_linear(bf16): tile %d x %d stages not instantiated", block_n, stages);
  return SC_OK;
}
