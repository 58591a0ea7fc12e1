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
// Warp roles (persistent CTA, one per SM): w0 TMA producer, w1 MMA issuer (one elected lane), w2 TMEM allocator,
// w4-11 epilogue (TMEM lane quarter = warp%4, two warps per quarter), w12-15 B-transform (kMasked only).
#include "sc_common.cuh"
#include <cuda.h>
#include <cstring>
#include <cstdlib>

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;  // 64 bf16 = 128 B = one swizzle row
constexpr int UMMA_K = 16;
constexpr int kNumTransformWarps = 4;
// Epilogue warps: 8 (two per TMEM lane quarter, splitting the tile's 32-column chunks) for the wide / deep-ring
// configurations that own an SM, 4 for the small-problem configurations, whose footprint (<= 113 KB smem, <= 128
// registers x 256 threads) lets two CTAs - e.g. of different streams - share an SM.
__host__ __device__ constexpr int num_epilogue_warps(int block_n, int stages) { return (block_n == 256 || stages >= 5) ? 8 : 4; }

struct GemmArgs {
  int M, N, K;
  int tiles_n, num_tiles, splits, kb_per_split;  // persistent schedule: unit u -> (tile = u / splits, split = u % splits)
  int tiles_per_cta;     // host only: > 1 caps the persistent grid at ceil(units / tiles_per_cta) CTAs
  size_t split_stride;   // split-K partial products: split s stores its fp32 tile at y + s * split_stride (0: single output)
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
  // folded LayerNorm (consumer) / residual-stream producer (ScGemmExtra)
  const float* ln_stats; const float* ln_c; float ln_eps;
  void* y2; float* stats_out;
  // cluster2: CTAs are launched as clusters of 2 that work on two M blocks of the same N tile in lockstep; each CTA loads its
  // own A tile and HALF of the shared B tile, multicast to both (L2 reads per CTA: 16 + 16 KB per k-block instead of 16 + 32)
  int cluster2;
  int pdl_early;  // trigger the dependent launch as soon as this CTA's loads are in flight (else: implicit, at exit)
  // kEpi == 3 (generator fused with the beam step's row pass): no output tile; per (row, N tile, epilogue-warp half) one
  // record {max, sum exp(x - max), kTopK largest values, their columns} -> topk_part[row][tiles_n * 2][kTopKRec]
  float* topk_part;
  // dX GEMM that prepares the NEXT linear's gradient operand (kEpi == 4, bf16 output, no residual): y = acc * hscale where
  // the saved post-ReLU/dropout activation hmask[row, col] != 0, else 0; colsum[col] += column sums of the bf16 values
  const __nv_bfloat16* hmask; float hscale; float* colsum;
  int dbg;  // diagnostics (SC_GEMM_DBG): 1 = top-k epilogue without the insertion, 2 = without the exp-sum as well
};

constexpr int kTopK = 5;                 // candidates kept per record (beam sizes up to 5 use the fused path)
constexpr int kTopKRec = 2 + 2 * kTopK;  // floats per record

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
__device__ __forceinline__ void tma_load_2d_mc(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void tcgen05_commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"(mask)
               : "memory");
}
// ---- CTA-pair mode (cluster2 == 2): one tcgen05.mma.cta_group::2 of M = 256 spans both CTAs' TMEM; every CTA stages its own
// 128 rows of A and its own HALF of the B tile (32 instead of 48 KB per k-block into each SM) ----
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
// TMA load whose completion bytes are counted on a barrier of the pair's leader CTA (shared::cluster address)
__device__ __forceinline__ void tma_load_2d_pair(const CUtensorMap* map, uint32_t bar_cluster, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(bar_cluster), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster) : "memory");
}
__device__ __forceinline__ void tcgen05_commit_pair(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"(mask)
               : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
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
// MN-major, SWIZZLE_128B operand tile (weight gradient: A[n, token] / B[k, token] read from row-major [tokens, n|k]
// activations WITHOUT a transpose): TMA boxes of 64 tokens x 64 MN-elements; one 128-byte row per token, 8 tokens per
// 1024-byte swizzle atom (stride byte offset), the next 64 MN-elements in the next 8 KB box (leading byte offset).
__device__ __forceinline__ uint64_t make_smem_desc_mn(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)(8192 >> 4) << 16;   // leading byte offset: next 64 elements along M/N
  d |= (uint64_t)(1024 >> 4) << 32;   // stride byte offset: next 8 elements along K
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// kind::f16, A=B=bf16, D=f32; mn = 0: both operands K-major, 1: both MN-major (bits 15 / 16)
__host__ __device__ constexpr uint32_t make_idesc(int m, int n, int mn = 0) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)mn << 15) | ((uint32_t)mn << 16) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(m >> 4) << 24);
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
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
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

// two chunks in flight before one wait
__device__ __forceinline__ void tmem_ld32x2(uint32_t ta, uint32_t (&v)[32], uint32_t tb, uint32_t (&w)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(ta));
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]), "=r"(w[8]),
        "=r"(w[9]), "=r"(w[10]), "=r"(w[11]), "=r"(w[12]), "=r"(w[13]), "=r"(w[14]), "=r"(w[15]), "=r"(w[16]),
        "=r"(w[17]), "=r"(w[18]), "=r"(w[19]), "=r"(w[20]), "=r"(w[21]), "=r"(w[22]), "=r"(w[23]), "=r"(w[24]),
        "=r"(w[25]), "=r"(w[26]), "=r"(w[27]), "=r"(w[28]), "=r"(w[29]), "=r"(w[30]), "=r"(w[31])
      : "r"(tb));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

template <int BLOCK_N, int kStages>
struct Smem {
  static constexpr int kNumEpilogueWarps = num_epilogue_warps(BLOCK_N, kStages);
  static constexpr int kABytes = BLOCK_M * BLOCK_K * 2;
  static constexpr int kBBytes = BLOCK_N * BLOCK_K * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStagingOffset = kStages * kStageBytes;  // epilogue warps x 4 KB transpose staging
  // direct (4-warp) configurations with 64-wide tiles prefetch the fp32 residual tile (128 rows x 64 columns) with
  // cp.async while the main loop runs; the area is shared with the weight-gradient staging (never used together)
  static constexpr int kResPrefetchBytes = (kNumEpilogueWarps == 4 && BLOCK_N == 64) ? BLOCK_M * BLOCK_N * 4 : 0;
  static constexpr int kEpiBytes = kNumEpilogueWarps * 4096 > kResPrefetchBytes ? kNumEpilogueWarps * 4096 : kResPrefetchBytes;
  // per-epilogue-warp bias / ln_c slices (BLOCK_N floats each): inside the (otherwise unused) transpose staging for the
  // direct configurations without residual prefetch, a dedicated area elsewhere
  static constexpr bool kSliceInStaging = (kNumEpilogueWarps == 4) && kResPrefetchBytes == 0;
  static constexpr int kSliceOffset = kStagingOffset + kEpiBytes;
  static constexpr int kSliceBytes = kSliceInStaging ? 0 : kNumEpilogueWarps * BLOCK_N * 8;
  static constexpr int kUsed = kStagingOffset + kEpiBytes + kSliceBytes;
  // 1 KB of slack: up to 768 B of alignment padding in front, the barriers + TMEM pointer in its last 256 B
  static constexpr int kTotal = kUsed + 1024;
};

// ---- epilogue ------------------------------------------------------------------------------------------------
// tcgen05.ld hands every thread one accumulator ROW (32 columns of a chunk); touching global memory in that mapping
// means 32 different 128-byte lines per warp instruction.  The chunk is therefore transposed through a 4 KB
// XOR-swizzled staging tile per warp, and all element-wise work (bias, ReLU, dropout, residual, folded LayerNorm,
// straight-through mask gradients) runs in the COALESCED mapping: 8 lanes x 16 bytes cover one row's 128 bytes, a
// warp instruction covers 4 full rows.
//
// Forward element-wise stage in the ROW mapping (thread = accumulator row, f = its 32 columns of the chunk, residual
// already added through the staging transpose): [folded LayerNorm] + bias, ReLU, dropout, + residual; kFull adds the
// dropout / LayerNorm / row-statistics code.
template <bool kFull>
__device__ __forceinline__ void epilogue_row(const GemmArgs& args, float (&f)[32], const float (&res)[32], int row, int col0,
                                             float ln_rstd, float ln_mr, const float* sb, const float* sc,
                                             const uint8_t* res_stage = nullptr, int res_sw = 0) {
  // res_stage: the thread's residual row piece still sits in the warp's (XOR-swizzled) staging tile and is added from there -
  // 32 registers less than a separate res[] array (the fast residual tile spilled its base pointers without this)
  // sb / sc: this chunk's 32 bias / ln_c values in shared memory (zero-filled beyond N; staged once per tile so that no
  // global latency sits between the accumulator load and the stores)
  if (kFull && args.ln_stats) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 c = *(const float4*)(sc + 4 * j);
      f[4 * j] = ln_rstd * f[4 * j] - ln_mr * c.x; f[4 * j + 1] = ln_rstd * f[4 * j + 1] - ln_mr * c.y;
      f[4 * j + 2] = ln_rstd * f[4 * j + 2] - ln_mr * c.z; f[4 * j + 3] = ln_rstd * f[4 * j + 3] - ln_mr * c.w;
    }
  }
  if (args.bias) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 b = *(const float4*)(sb + 4 * j);
      f[4 * j] += b.x; f[4 * j + 1] += b.y; f[4 * j + 2] += b.z; f[4 * j + 3] += b.w;
    }
  }
  if (args.relu) {
#pragma unroll
    for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
  }
  if (kFull && args.dropout_p > 0.f) {
    // inverted dropout, keep mask = Philox(drop_seed, drop_stream, element); one draw serves 4 consecutive elements
    const sc::Philox dph(args.drop_seed);
    const float keep = 1.f / (1.f - args.dropout_p);
    const size_t e0 = (size_t)row * args.N + col0;
    if ((e0 & 3) == 0) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const uint4 rr = dph((e0 >> 2) + j, args.drop_stream);
        f[4 * j] *= sc::u24(rr.x) >= args.dropout_p ? keep : 0.f; f[4 * j + 1] *= sc::u24(rr.y) >= args.dropout_p ? keep : 0.f;
        f[4 * j + 2] *= sc::u24(rr.z) >= args.dropout_p ? keep : 0.f; f[4 * j + 3] *= sc::u24(rr.w) >= args.dropout_p ? keep : 0.f;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) f[j] *= sc::keep_scale(dph, e0 + j, args.drop_stream, args.dropout_p);
    }
  }
  if (args.residual) {
    if (res_stage != nullptr) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 r4 = *(const float4*)(res_stage + ((j ^ res_sw) << 4));
        f[4 * j] += r4.x; f[4 * j + 1] += r4.y; f[4 * j + 2] += r4.z; f[4 * j + 3] += r4.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) f[j] += res[j];
    }
  }
  if (kFull && args.stats_out) {
    // (sum, M2 about the chunk mean) of this row's 32-column chunk; the LayerNorm consumer merges the N/32 chunks
    float sm = 0.f;
#pragma unroll
    for (int j = 0; j < 32; ++j) sm += f[j];
    const float mu = sm * (1.f / 32.f);
    float m2 = 0.f;
#pragma unroll
    for (int j = 0; j < 32; ++j) { const float dlt = f[j] - mu; m2 += dlt * dlt; }
    if (row < args.M) *(float2*)(args.stats_out + ((size_t)row * (args.N >> 5) + (col0 >> 5)) * 2) = make_float2(sm, m2);
  }
}

// Store stage in the ROW mapping (latency-oriented small-problem configurations): the thread writes its 32 columns.
template <bool kFull>
__device__ __forceinline__ void epilogue_store_row(const GemmArgs& args, const float (&f)[32], int row, int col0, size_t yoff) {
  const size_t e0 = (size_t)row * args.N + col0 + yoff;
  if ((args.N & 7) == 0 && col0 + 32 <= args.N) {
    if (args.y_bf16 || (kFull && args.y2)) {
      uint4* yp = (uint4*)((__nv_bfloat16*)(args.y_bf16 ? args.y : args.y2) + e0);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const __nv_bfloat162 p0 = __floats2bfloat162_rn(f[8 * j], f[8 * j + 1]), p1 = __floats2bfloat162_rn(f[8 * j + 2], f[8 * j + 3]);
        const __nv_bfloat162 p2 = __floats2bfloat162_rn(f[8 * j + 4], f[8 * j + 5]), p3 = __floats2bfloat162_rn(f[8 * j + 6], f[8 * j + 7]);
        uint4 o;
        o.x = *(const uint32_t*)&p0; o.y = *(const uint32_t*)&p1; o.z = *(const uint32_t*)&p2; o.w = *(const uint32_t*)&p3;
        yp[j] = o;
      }
    }
    if (!args.y_bf16) {
      float4* yp = (float4*)((float*)args.y + e0);
#pragma unroll
      for (int j = 0; j < 8; ++j) yp[j] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
    }
    return;
  }
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    if (col0 + j >= args.N) continue;
    if (args.y_bf16) ((__nv_bfloat16*)args.y)[e0 + j] = __float2bfloat16_rn(f[j]);
    else ((float*)args.y)[e0 + j] = f[j];
  }
}

// Store stage in the COALESCED mapping: 4 consecutive columns of one row (fp32 and/or bf16 copy).
template <bool kFull>
__device__ __forceinline__ void epilogue_store4(const GemmArgs& args, const float4& f, int row, int col, size_t yoff) {
  const size_t e = (size_t)row * args.N + col + yoff;
  if ((args.N & 3) == 0) {
    if (args.y_bf16 || (kFull && args.y2)) {
      const __nv_bfloat162 lo = __floats2bfloat162_rn(f.x, f.y), hi = __floats2bfloat162_rn(f.z, f.w);
      uint2 o;
      o.x = *(const uint32_t*)&lo; o.y = *(const uint32_t*)&hi;
      *(uint2*)((__nv_bfloat16*)(args.y_bf16 ? args.y : args.y2) + e) = o;
    }
    if (!args.y_bf16) *(float4*)((float*)args.y + e) = f;
    return;
  }
  const float x[4] = {f.x, f.y, f.z, f.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (col + i >= args.N) break;
    if (args.y_bf16) ((__nv_bfloat16*)args.y)[e + i] = __float2bfloat16_rn(x[i]);
    else ((float*)args.y)[e + i] = x[i];
  }
}

// One float4 of the weight-gradient tile (K2): dW = dWm (.) m ; dS = dWm (.) W (.) sigmoid'(S) (+ sparsity term); the
// mask is regenerated from (seed, stream, element).  Split-K partial sums go out as vector reductions into
// pre-zeroed buffers; the sparsity term is added by split 0 only.
__device__ __forceinline__ void epilogue_wgrad4(const GemmArgs& args, float4 f, int row, int col, bool atomic, bool first_split,
                                                const float4& w4, const float4& s4, const float4& u4) {
  const sc::Philox wph(args.seed);
  const float sp = first_split ? args.sp_coeff : 0.f;
  const size_t e = (size_t)row * args.N + col;
  const float g[4] = {f.x, f.y, f.z, f.w};
  if ((args.N & 3) == 0) {
    const float wv[4] = {w4.x, w4.y, w4.z, w4.w}, sv[4] = {s4.x, s4.y, s4.z, s4.w}, uv[4] = {u4.x, u4.y, u4.z, u4.w};
    float m[4];
    if (args.mask_mode == SC_MASK_BERNOULLI) {
      sc::bernoulli4(wph, e >> 2, args.stream_id, sv, m);
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) m[i] = sc::mask_value(args.mask_mode, sv[i], uv[i], wph, e + i, args.stream_id);
    }
    float gw[4], gs[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) sc::mask_grad_elem(args.mask_mode, g[i], wv[i], sv[i], m[i], args.bypass, sp, gw[i], gs[i]);
    if (atomic) {
      if (args.dw)
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(args.dw + e), "f"(gw[0]), "f"(gw[1]), "f"(gw[2]), "f"(gw[3]) : "memory");
      if (args.ds)
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(args.ds + e), "f"(gs[0]), "f"(gs[1]), "f"(gs[2]), "f"(gs[3]) : "memory");
    } else {
      if (args.dw) {
        float4 o = make_float4(gw[0], gw[1], gw[2], gw[3]);
        if (args.accumulate) { const float4 p = *(const float4*)(args.dw + e); o.x += p.x; o.y += p.y; o.z += p.z; o.w += p.w; }
        *(float4*)(args.dw + e) = o;
      }
      if (args.ds) {
        float4 o = make_float4(gs[0], gs[1], gs[2], gs[3]);
        if (args.accumulate) { const float4 p = *(const float4*)(args.ds + e); o.x += p.x; o.y += p.y; o.z += p.z; o.w += p.w; }
        *(float4*)(args.ds + e) = o;
      }
    }
    return;
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (col + i >= args.N) break;
    const float sv = args.wg_s ? __ldg(args.wg_s + e + i) : 0.f;
    const float m = sc::mask_value(args.mask_mode, sv, args.wg_u ? __ldg(args.wg_u + e + i) : 0.f, wph, e + i, args.stream_id);
    float gw, gs;
    sc::mask_grad_elem(args.mask_mode, g[i], __ldg(args.wg_w + e + i), sv, m, args.bypass, sp, gw, gs);
    if (atomic) {
      if (args.dw) atomicAdd(args.dw + e + i, gw);
      if (args.ds) atomicAdd(args.ds + e + i, gs);
    } else {
      if (args.dw) args.dw[e + i] = (args.accumulate ? args.dw[e + i] : 0.f) + gw;
      if (args.ds) args.ds[e + i] = (args.accumulate ? args.ds[e + i] : 0.f) + gs;
    }
  }
}

// ---- kEpi == 3 / 5: generator GEMM fused with the beam step's row pass ----------------------------------------------------
// Per thread (= accumulator row) and chunk of 32 columns: running max, sum exp(x - max) (log2 domain: FFMA + MUFU.EX2 + FADD per
// element) and the kTK largest (value, column) pairs by a branch-free sorted insertion (a divergent `if (candidate)` ran for
// nearly every element: some lane of the 32 independent rows always had one).  The insertion is 3 FSETP + 10 SEL per element on
// the half-rate ALU pipe and a serial chain through the kTK slots: the thread therefore runs TWO independent chains (chunk pairs)
// whose instructions interleave, and merges them once per tile.
template <int kTK>
struct TopkState {
  float m, s;
  float v[kTK];
  int i[kTK];
  __device__ __forceinline__ void init() {
    m = -INFINITY; s = 0.f;
#pragma unroll
    for (int k = 0; k < kTK; ++k) { v[k] = -INFINITY; i[k] = 0x7fffffff; }
  }
  // columns arrive in ascending order: strict '>' keeps the earlier (smaller) column on ties
  __device__ __forceinline__ void insert(float x, int col) {
    bool pgt[kTK];
#pragma unroll
    for (int k = 0; k < kTK; ++k) pgt[k] = x > v[k];
#pragma unroll
    for (int k = kTK - 1; k >= 1; --k) {
      v[k] = pgt[k - 1] ? v[k - 1] : (pgt[k] ? x : v[k]);
      i[k] = pgt[k - 1] ? i[k - 1] : (pgt[k] ? col : i[k]);
    }
    v[0] = pgt[0] ? x : v[0];
    i[0] = pgt[0] ? col : i[0];
  }
  // any column order (merging the second chain): equal values rank by the smaller column
  __device__ __forceinline__ void insert_any(float x, int col) {
    bool pgt[kTK];
#pragma unroll
    for (int k = 0; k < kTK; ++k) pgt[k] = x > v[k] || (x == v[k] && col < i[k]);
#pragma unroll
    for (int k = kTK - 1; k >= 1; --k) {
      v[k] = pgt[k - 1] ? v[k - 1] : (pgt[k] ? x : v[k]);
      i[k] = pgt[k - 1] ? i[k - 1] : (pgt[k] ? col : i[k]);
    }
    v[0] = pgt[0] ? x : v[0];
    i[0] = pgt[0] ? col : i[0];
  }
};

__device__ __forceinline__ float max32(const float (&f)[32]) {
  float c[4] = {f[0], f[1], f[2], f[3]};
#pragma unroll
  for (int j = 4; j < 32; ++j) c[j & 3] = fmaxf(c[j & 3], f[j]);
  return fmaxf(fmaxf(c[0], c[1]), fmaxf(c[2], c[3]));
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// one chunk (kTwo: two chunks, instruction streams interleaved) of 32 accumulator columns + bias into the running statistics
template <int kTK, bool kTwo>
__device__ __forceinline__ void topk_absorb(TopkState<kTK>& A, float (&fa)[32], int cola, const float* ba, TopkState<kTK>& B,
                                            float (&fb)[32], int colb, const float* bb, int N, int dbg) {
  constexpr float kL2E = 1.4426950408889634f;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float4 x = *(const float4*)(ba + 4 * j);
    fa[4 * j] += x.x; fa[4 * j + 1] += x.y; fa[4 * j + 2] += x.z; fa[4 * j + 3] += x.w;
    if (kTwo) {
      const float4 y = *(const float4*)(bb + 4 * j);
      fb[4 * j] += y.x; fb[4 * j + 1] += y.y; fb[4 * j + 2] += y.z; fb[4 * j + 3] += y.w;
    }
  }
  if (cola + 32 > N) {  // warp-uniform: only the tail chunk of the last N tile has columns to blank
#pragma unroll
    for (int j = 0; j < 32; ++j) if (cola + j >= N) fa[j] = -INFINITY;
  }
  if (kTwo && colb + 32 > N) {
#pragma unroll
    for (int j = 0; j < 32; ++j) if (colb + j >= N) fb[j] = -INFINITY;
  }
  const float cma = max32(fa);
  if (cma > A.m) { A.s *= __expf(A.m - cma); A.m = cma; }  // first chunk: 0 * exp(-inf) = 0
  const float nma = -A.m * kL2E;
  float nmb = 0.f;
  if (kTwo) {
    const float cmb = max32(fb);
    if (cmb > B.m) { B.s *= __expf(B.m - cmb); B.m = cmb; }
    nmb = -B.m * kL2E;
  }
  float pa[4] = {0.f, 0.f, 0.f, 0.f}, pb[4] = {0.f, 0.f, 0.f, 0.f};
  if (dbg >= 2) return;
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    pa[j & 3] += ex2_approx(fmaf(fa[j], kL2E, nma));
    if (kTwo) pb[j & 3] += ex2_approx(fmaf(fb[j], kL2E, nmb));
  }
  A.s += (pa[0] + pa[1]) + (pa[2] + pa[3]);
  if (kTwo) B.s += (pb[0] + pb[1]) + (pb[2] + pb[3]);
  if (dbg >= 1) return;
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    A.insert(fa[j], cola + j);
    if (kTwo) B.insert(fb[j], colb + j);
  }
}

// The fp32 residual chunk (32 columns x 32 rows of one epilogue warp) in the COALESCED mapping: lane = 16-byte piece (lane & 7) of
// rows (lane >> 3) + 4 i.  Requested one chunk ahead of its use (the first one before the accumulator is awaited): the global
// latency (an L2 / HBM round trip per chunk, 4-8 chunks per tile) used to sit between every tcgen05.ld and its stores.
__device__ __forceinline__ void load_residual8(const GemmArgs& args, int rbase, int lane, int col, float4 (&r)[8]) {
  const bool vec = (args.N & 3) == 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    r[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    const int row = rbase + i * 4 + (lane >> 3);
    if (row < args.M && col < args.N) {
      const float* rp = args.residual + (size_t)row * args.N + col;  // plain loads: written by the predecessor
      if (vec) r[i] = *(const float4*)rp;
      else {
        r[i].x = rp[0];
        if (col + 1 < args.N) r[i].y = rp[1];
        if (col + 2 < args.N) r[i].z = rp[2];
        if (col + 3 < args.N) r[i].w = rp[3];
      }
    }
  }
}

// Persistent kernel: CTA c processes work units c, c + gridDim.x, ...  (unit = output tile x K split; N-tiles of one
// M block are adjacent units so that concurrently running CTAs share the A tile through L2).  The accumulator is
// double-buffered in TMEM (2 x BLOCK_N columns): the epilogue of unit i overlaps the TMA/MMA main loop of unit i+1.
// kStages: depth of the TMA->MMA smem ring.
// kEpi: 0 = plain forward epilogue, 1 = full forward epilogue (dropout, folded LayerNorm, statistics), 2 = weight gradient
// kMN: both operands are given transposed ([K, M] and [K, N] row-major, i.e. MN-major tiles): y = x^T w.
// kPair: CTA-pair variant (launched as clusters of 2 with args.cluster2 == 2).  A separate instantiation: a kernel that contains
// cta_group::2 instructions cannot be launched without a cluster.
// kResCo: the coalesced fp32-residual epilogue of the 8-epilogue-warp configurations is compiled in (launched when there IS a
// residual).  A separate instantiation because its register footprint (a prefetched residual chunk + two base pointers live
// across the chunk loop) cost the residual-free epilogues of the same kernel 13-15 % (same-box A/B, decode GEMMs).
template <int BLOCK_N, bool kMasked, int kStages, int kEpi, bool kMN = false, bool kPair = false, bool kResCo = false>
__global__ void __launch_bounds__(32 * (4 + num_epilogue_warps(BLOCK_N, kStages) + (kMasked ? kNumTransformWarps : 0)),
                                  (!kMasked && kEpi != 2 && num_epilogue_warps(BLOCK_N, kStages) == 4) ? 2 : 1)
sc_gemm_bf16_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b, const GemmArgs args) {
  // cluster2 mode: tma_b describes boxes of BLOCK_N / 2 rows (the half this CTA multicasts)
  const bool c2 = args.cluster2 != 0;
  constexpr bool pair = kPair;  // one cta_group::2 MMA per CTA pair, issued by rank 0
  if (pair) cluster_sync_all();  // both CTAs are resident before the pair-wide TMEM allocation
  const uint32_t crank = c2 ? cluster_ctarank() : 0u;
  const int unit0 = c2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int ustride = c2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  using L = Smem<BLOCK_N, kStages>;
  constexpr int kNumEpilogueWarps = L::kNumEpilogueWarps;
  constexpr int kFirstTransformWarp = 4 + kNumEpilogueWarps;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  if (smem - smem_raw > 768) __trap();  // the barrier block below would overlap the tiles (never with a 1 KB-aligned base)
  uint64_t* full_bar = (uint64_t*)(smem_raw + L::kTotal - 256);
  // pair mode stages 32 instead of 48 KB per k-block: the same ring bytes hold 3/2 as many stages (the main loop is bound by
  // bytes in flight = ring bytes / TMA latency, so the deeper ring is where the pair mode gains)
  const int nst = pair ? kStages * 3 / 2 : kStages;
  const int stage_bytes = pair ? L::kABytes + L::kBBytes / 2 : L::kStageBytes;
  uint64_t* empty_bar = full_bar + nst;
  uint64_t* tmem_full_bar = empty_bar + nst;   // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;    // [2]
  uint32_t* tmem_ptr_smem = (uint32_t*)(tmem_empty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_kb_total = (args.K + BLOCK_K - 1) / BLOCK_K;
  const int num_units = args.num_tiles * args.splits;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tma_a) : "memory");
    if (!kMasked) asm volatile("prefetch.tensormap [%0];" ::"l"(&tma_b) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < nst; ++s) {
      mbar_init(&full_bar[s], kMasked ? 1 + kNumTransformWarps : 1);
      // multicast mode: both CTAs' MMAs must have read the slot before it is refilled; pair mode: one multicast commit
      mbar_init(&empty_bar[s], (c2 && !pair) ? 2 : 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tmem_full_bar[b], 1);
      mbar_init(&tmem_empty_bar[b], pair ? 2 * kNumEpilogueWarps : kNumEpilogueWarps);  // pair: the peer's epilogue arrives here too
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    if (pair) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)),
                   "r"((uint32_t)(2 * BLOCK_N)));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)),
                   "r"((uint32_t)(2 * BLOCK_N)));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (c2) cluster_sync_all();  // the peer's barriers are initialised before anything multicasts into / arrives on them
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      sc::pdl_wait();  // A (and B) may be written by the previous kernel in the stream
      int it = 0;
      for (int u = unit0; u < num_units; u += ustride) {
        const int tile = u / args.splits, split = u - tile * args.splits;
        const int m0 = (c2 ? (tile / args.tiles_n) * 2 + (int)crank : tile / args.tiles_n) * BLOCK_M, n0 = (tile % args.tiles_n) * BLOCK_N;
        const int kb0 = split * args.kb_per_split;
        const int kb1 = min(kb0 + args.kb_per_split, num_kb_total);
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const int s = pair ? it % nst : it % kStages;
          const uint32_t ph = (pair ? it / nst : it / kStages) & 1;
          mbar_wait(&empty_bar[s], ph ^ 1);
          uint8_t* sa = smem + s * stage_bytes;
          if (!kMasked && !kMN && pair) {
            // both CTAs' bytes (A block + half of the B tile each) are counted on the leader's barrier
            const uint32_t fb = mapa_u32(smem_u32(&full_bar[s]), 0u);
            if (crank == 0) mbar_expect_tx(&full_bar[s], 2 * (L::kABytes + L::kBBytes / 2));
            tma_load_2d_pair(&tma_a, fb, sa, kb * BLOCK_K, m0);
            tma_load_2d_pair(&tma_b, fb, sa + L::kABytes, kb * BLOCK_K, n0 + (int)crank * (BLOCK_N / 2));
            continue;
          }
          mbar_expect_tx(&full_bar[s], kMasked ? L::kABytes : L::kStageBytes);
          if (kMN) {
            // boxes of {64 MN-elements, 64 tokens}: coordinate 0 = position along M / N, coordinate 1 = token
#pragma unroll
            for (int hb = 0; hb < BLOCK_M / 64; ++hb) tma_load_2d(&tma_a, &full_bar[s], sa + hb * 8192, m0 + hb * 64, kb * BLOCK_K);
#pragma unroll
            for (int hb = 0; hb < BLOCK_N / 64; ++hb)
              tma_load_2d(&tma_b, &full_bar[s], sa + L::kABytes + hb * 8192, n0 + hb * 64, kb * BLOCK_K);
          } else {
            tma_load_2d(&tma_a, &full_bar[s], sa, kb * BLOCK_K, m0);
            if (!kMasked) {
              if (c2) tma_load_2d_mc(&tma_b, &full_bar[s], sa + L::kABytes + crank * (L::kBBytes / 2), kb * BLOCK_K,
                                     n0 + (int)crank * (BLOCK_N / 2), (uint16_t)3);
              else tma_load_2d(&tma_b, &full_bar[s], sa + L::kABytes, kb * BLOCK_K, n0);
            }
          }
        }
      }
      if (args.pdl_early) sc::pdl_launch();  // all loads of this CTA are in flight: let the next kernel's prologue start
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0 && !(pair && crank != 0)) {
      constexpr uint32_t idesc = make_idesc(BLOCK_M, BLOCK_N, kMN ? 1 : 0);
      constexpr uint32_t idesc_pair = make_idesc(2 * BLOCK_M, BLOCK_N, 0);
      int it = 0, lt = 0;
      for (int u = unit0; u < num_units; u += ustride, ++lt) {
        const int split = u % args.splits;
        const int kb0 = split * args.kb_per_split;
        const int kb1 = min(kb0 + args.kb_per_split, num_kb_total);
        const int buf = lt & 1;
        mbar_wait(&tmem_empty_bar[buf], ((lt >> 1) & 1) ^ 1);  // epilogue has drained this accumulator
        tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(buf * BLOCK_N);
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const int s = pair ? it % nst : it % kStages;
          const uint32_t ph = (pair ? it / nst : it / kStages) & 1;
          mbar_wait(&full_bar[s], ph);
          tcgen05_fence_after();
          const uint32_t sa = smem_u32(smem + s * stage_bytes);
          const uint64_t da = kMN ? make_smem_desc_mn(sa) : make_smem_desc(sa);
          const uint64_t db = kMN ? make_smem_desc_mn(sa + L::kABytes) : make_smem_desc(sa + L::kABytes);
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            // K-major: advance 32 B (16 bf16) along K inside the 128B swizzle row: +2 in 16-byte units;
            // MN-major: 16 tokens = two 1024-byte atoms: +128 units
            const uint64_t adv = kMN ? (uint64_t)(128 * k) : (uint64_t)(2 * k);
            if (pair) umma_bf16_pair(tmem_d, da + adv, db + adv, idesc_pair, (kb > kb0 || k > 0) ? 1u : 0u);
            else umma_bf16(tmem_d, da + adv, db + adv, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          if (pair) tcgen05_commit_pair(&empty_bar[s], (uint16_t)3);
          else if (c2) tcgen05_commit_mc(&empty_bar[s], (uint16_t)3);
          else tcgen05_commit(&empty_bar[s]);  // frees the smem slot once these MMAs have read it
        }
        if (pair) tcgen05_commit_pair(&tmem_full_bar[buf], (uint16_t)3);  // both CTAs' epilogues
        else tcgen05_commit(&tmem_full_bar[buf]);                          // accumulator complete
      }
    }
  } else if (warp >= 4 && warp < 4 + kNumEpilogueWarps) {
    // ===== epilogue: TMEM -> registers -> swizzled staging (transpose) -> coalesced global access =====
    sc::pdl_wait();  // residual / statistics come from, and y may still be read by, the previous kernel
    const int ew = warp - 4;
    const int q = ew & 3;      // TMEM lane quarter this warp may read
    const int half = ew >> 2;  // 8 warps: which of the tile's 32-column chunks (even / odd) this warp takes
    constexpr int kChunkStep = kNumEpilogueWarps / 4;
    // kDirect: single-wave, latency-bound problems.  The thread keeps its accumulator row, loads / prefetches the
    // residual in the same mapping and stores straight from registers (no transposes on the critical path).
    constexpr bool kDirect = (kNumEpilogueWarps == 4) && (kEpi != 2);
    constexpr bool kPrefetchRes = kDirect && L::kResPrefetchBytes > 0;
    uint8_t* stg = smem + L::kStagingOffset + ew * 4096;
    float4* pre = (float4*)(smem + L::kStagingOffset);  // residual prefetch: piece j of epilogue thread t at [j * 128 + t]
    const int et = threadIdx.x - 128;                    // 0..127 (4-warp configurations)
    const int jsw = lane & 7;  // swizzle key of this thread's own row (row-mapping: row = lane)
    // coalesced-path residual, software-pipelined TWO chunks ahead (two register buffers; see load_residual8) and across tiles:
    // the first two chunks of the NEXT tile are requested while the last two of this one are processed
    constexpr bool kPipeRes = !kDirect && kEpi != 2 && kResCo;
    float4 rnext[kPipeRes ? 8 : 1], rnext2[kPipeRes ? 8 : 1];
    bool pre_tile = false;  // rnext / rnext2 already hold the requests of this tile's first chunks
    int lt = 0;
    for (int u = unit0; u < num_units; u += ustride, ++lt) {
      const int tile = u / args.splits, split = u - tile * args.splits;
      const int m0 = (c2 ? (tile / args.tiles_n) * 2 + (int)crank : tile / args.tiles_n) * BLOCK_M, n0 = (tile % args.tiles_n) * BLOCK_N;
      const int buf = lt & 1;
      const int rbase = m0 + q * 32;
      float ln_rstd = 1.f, ln_mr = 0.f;
      if (kEpi == 1 && args.ln_stats && rbase + lane < args.M) {
        // merge the K/32 chunk statistics of row (rbase + lane) (Chan et al.): unbiased std, a*(x-mean)/(std+eps)+b
        const int parts = args.K >> 5;
        const float2* sp = (const float2*)(args.ln_stats + (size_t)(rbase + lane) * parts * 2);
        float tot = 0.f, m2 = 0.f, mean;
        if (parts == 16) {
          // d_model = 512: the 16 (sum, M2) pairs of the row are 128 contiguous bytes -> 8 independent 16-byte loads
          float4 sv[8];
#pragma unroll
          for (int p = 0; p < 8; ++p) sv[p] = ((const float4*)sp)[p];
#pragma unroll
          for (int p = 0; p < 8; ++p) tot += sv[p].x + sv[p].z;
          mean = tot / (float)args.K;
#pragma unroll
          for (int p = 0; p < 8; ++p) {
            const float d0 = sv[p].x * (1.f / 32.f) - mean, d1 = sv[p].z * (1.f / 32.f) - mean;
            m2 += sv[p].y + sv[p].w + 32.f * (d0 * d0 + d1 * d1);
          }
        } else {
          for (int p = 0; p < parts; ++p) tot += sp[p].x;
          mean = tot / (float)args.K;
          for (int p = 0; p < parts; ++p) {
            const float2 pv = sp[p];
            const float dlt = pv.x * (1.f / 32.f) - mean;
            m2 += pv.y + 32.f * dlt * dlt;
          }
        }
        ln_rstd = 1.f / (sqrtf(m2 / (float)(args.K - 1)) + args.ln_eps);
        ln_mr = ln_rstd * mean;
      }
      // this tile's bias / ln_c values -> the warp's shared-memory slice
      float* sbias = (float*)(L::kSliceInStaging ? (smem + L::kStagingOffset + ew * 4096) : (smem + L::kSliceOffset + ew * BLOCK_N * 8));
      float* slnc = sbias + BLOCK_N;
      if (kEpi != 2) {
        __syncwarp();
        for (int i = lane; i < BLOCK_N; i += 32) {
          const int cc = n0 + i;
          sbias[i] = (args.bias && cc < args.N) ? __ldg(args.bias + cc) : 0.f;
          if (kEpi == 1) slnc[i] = (args.ln_stats && cc < args.N) ? __ldg(args.ln_c + cc) : 0.f;
        }
        __syncwarp();
      }
      bool pre_ok = false;
      if (kPrefetchRes) {
        // fp32 residual row of this thread -> shared memory while the main loop runs (16-byte cp.async, no registers)
        pre_ok = args.residual != nullptr && (args.N & 3) == 0 && rbase + lane < args.M;
        if (pre_ok) {
          const float* rp = args.residual + (size_t)(rbase + lane) * args.N + n0;
#pragma unroll
          for (int j = 0; j < BLOCK_N / 4; ++j) {
            if (n0 + 4 * j < args.N)
              asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(pre + j * 128 + et)), "l"(rp + 4 * j) : "memory");
          }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
      }
      const bool pipe_res = kPipeRes && args.residual != nullptr;
      // fast tile: the warp's 32 rows and the whole N tile lie inside the matrix, rows are 16-byte aligned, plain fp32 output -
      // every chunk is straight-line code from two per-lane base pointers (the generic path spends ~700 instructions per 32 x 32
      // chunk on bounds checks, dtype branches and 64-bit address arithmetic for 64 useful FADDs: it made the fp32-residual
      // epilogue, not the MMAs, the long pole of the N = d_model GEMMs)
      const bool fast_tile = pipe_res && !args.y_bf16 && (args.N & 3) == 0 && rbase + 32 <= args.M && n0 + BLOCK_N <= args.N;
      const size_t lane_off = (size_t)(rbase + (lane >> 3)) * args.N + n0 + (lane & 7) * 4;
      const size_t row4 = (size_t)4 * args.N;  // this lane's next row (4 rows down)
      const float* res_lane = args.residual + lane_off;
      float* y_lane = (float*)args.y + (size_t)split * args.split_stride + lane_off;
      constexpr int kWarpChunks = BLOCK_N / 32 / kChunkStep;  // chunks of a tile per epilogue warp
      if constexpr (kPipeRes) {
        if (fast_tile) {
          if (!pre_tile) {
#pragma unroll
            for (int i = 0; i < 8; ++i) rnext[i] = *(const float4*)(res_lane + half * 32 + i * row4);
            if constexpr (kWarpChunks > 1) {
#pragma unroll
              for (int i = 0; i < 8; ++i) rnext2[i] = *(const float4*)(res_lane + (half + kChunkStep) * 32 + i * row4);
            }
          }
        } else if (pipe_res && rbase < args.M && n0 + half * 32 < args.N) {
          load_residual8(args, rbase, lane, n0 + half * 32 + (lane & 7) * 4, rnext);
        }
        pre_tile = false;
      }
      mbar_wait(&tmem_full_bar[buf], (lt >> 1) & 1);
      tcgen05_fence_after();
      if (kPrefetchRes) asm volatile("cp.async.wait_group 0;" ::: "memory");
      // kEpi == 3 / 5: running log-sum-exp statistics and top-kTK (5 / 3) of this thread's row over the warp's chunks of the
      // tile (beam <= 3 uses the top-3 variant: the depth of the insertion is what the epilogue costs)
      constexpr bool kTopkEpi = (kEpi == 3 || kEpi == 5);
      constexpr int kTK = (kEpi == 5) ? 3 : kTopK;
      if constexpr (kTopkEpi) {
        TopkState<kTK> sa, sb2;
        sa.init(); sb2.init();
        if (rbase < args.M) {
#pragma unroll 1
          for (int c = half; c < BLOCK_N / 32; c += 2 * kChunkStep) {
            const int cola = n0 + c * 32, colb = cola + kChunkStep * 32;
            if (cola >= args.N) break;  // warp-uniform
            const uint32_t ta_ = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BLOCK_N + c * 32);
            uint32_t va[32];
            float fa[32], fb[32];
            if (c + kChunkStep < BLOCK_N / 32 && colb < args.N) {
              uint32_t vb[32];
              tmem_ld32x2(ta_, va, ta_ + (uint32_t)(kChunkStep * 32), vb);
#pragma unroll
              for (int j = 0; j < 32; ++j) { fa[j] = __uint_as_float(va[j]); fb[j] = __uint_as_float(vb[j]); }
              topk_absorb<kTK, true>(sa, fa, cola, sbias + c * 32, sb2, fb, colb, sbias + (c + kChunkStep) * 32, args.N, args.dbg);
            } else {
              tmem_ld32(ta_, va);
#pragma unroll
              for (int j = 0; j < 32; ++j) fa[j] = __uint_as_float(va[j]);
              topk_absorb<kTK, false>(sa, fa, cola, sbias + c * 32, sb2, fb, colb, sbias, args.N, args.dbg);
            }
          }
          // merge the second chain (its columns interleave with the first one's: full (value, column) order)
          const float mm = fmaxf(sa.m, sb2.m);
          if (mm > -INFINITY) sa.s = sa.s * __expf(sa.m - mm) + sb2.s * __expf(sb2.m - mm);
          sa.m = mm;
#pragma unroll
          for (int k = 0; k < kTK; ++k) sa.insert_any(sb2.v[k], sb2.i[k]);
        }
        if (rbase + lane < args.M) {
          float* rec = args.topk_part + ((size_t)(rbase + lane) * (args.tiles_n * kChunkStep) + (size_t)(tile % args.tiles_n) * kChunkStep + half) * kTopKRec;
          rec[0] = sa.m; rec[1] = sa.s;
#pragma unroll
          for (int k = 0; k < kTopK; ++k) {
            rec[2 + k] = k < kTK ? sa.v[k < kTK ? k : 0] : -INFINITY;
            rec[2 + kTopK + k] = __int_as_float(k < kTK ? sa.i[k < kTK ? k : 0] : 0x7fffffff);
          }
        }
      }
      bool fast_done = false;
      if constexpr (kPipeRes) {
        if (fast_tile) {
          // staging offsets: row rl = 4 i + (lane >> 3), piece pj ^ (rl & 7); (rl & 7) = (lane >> 3) + 4 (i & 1)
          const int pj = lane & 7;
          uint8_t* stc = stg + (lane >> 3) * 128;
          const uint32_t sw0 = (uint32_t)((pj ^ (lane >> 3)) << 4);
          uint8_t* str_ = stg + lane * 128;
          // the next unit of this CTA: its first chunks are requested underneath the last chunks of this tile
          const float* res_nt = nullptr;
          {
            const int un = u + ustride;
            if (un < num_units) {
              const int tn = un / args.splits;
              const int m0n = (c2 ? (tn / args.tiles_n) * 2 + (int)crank : tn / args.tiles_n) * BLOCK_M, n0n = (tn % args.tiles_n) * BLOCK_N;
              if (m0n + q * 32 + 32 <= args.M && n0n + BLOCK_N <= args.N)
                res_nt = args.residual + (size_t)(m0n + q * 32 + (lane >> 3)) * args.N + n0n + pj * 4;
            }
          }
          // one chunk: residual (requested earlier into rb) -> staging; request `reload` into rb; accumulator; element-wise stage
          // with the residual added straight from the staging tile; results -> staging -> coalesced stores
          auto fast_chunk = [&](int c, float4 (&rb)[8], const float* reload) {
#pragma unroll
            for (int i = 0; i < 8; ++i) *(float4*)(stc + i * 512 + (sw0 ^ ((i & 1) << 6))) = rb[i];
            __syncwarp();
            if (reload != nullptr) {
#pragma unroll
              for (int i = 0; i < 8; ++i) rb[i] = *(const float4*)(reload + i * row4);
            }
            uint32_t v[32];
            tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BLOCK_N + c * 32), v);
            float f[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
            epilogue_row<kEpi == 1>(args, f, f, rbase + lane, n0 + c * 32, ln_rstd, ln_mr, sbias + c * 32, slnc + c * 32, str_, jsw);
            __syncwarp();  // every lane has read its residual piece: the tile may take the results
#pragma unroll
            for (int j = 0; j < 8; ++j)
              *(float4*)(str_ + ((j ^ jsw) << 4)) = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 o = *(const float4*)(stc + i * 512 + (sw0 ^ ((i & 1) << 6)));
              *(float4*)(y_lane + c * 32 + i * row4) = o;
              if constexpr (kEpi == 1) {
                if (args.y2) {  // bf16 copy of the residual stream (LayerNorm-folding consumers read it as their A operand)
                  const __nv_bfloat162 lo = __floats2bfloat162_rn(o.x, o.y), hi = __floats2bfloat162_rn(o.z, o.w);
                  uint2 pk;
                  pk.x = *(const uint32_t*)&lo; pk.y = *(const uint32_t*)&hi;
                  *(uint2*)((__nv_bfloat16*)args.y2 + lane_off + c * 32 + i * row4) = pk;
                }
              }
            }
            __syncwarp();  // staging is rewritten by the next chunk
          };
#pragma unroll
          for (int k = 0; k < kWarpChunks; k += 2) {
            const int c = half + k * kChunkStep;
            fast_chunk(c, rnext, k + 2 < kWarpChunks ? res_lane + (c + 2 * kChunkStep) * 32 : (res_nt ? res_nt + half * 32 : nullptr));
            if (k + 1 < kWarpChunks)
              fast_chunk(c + kChunkStep, rnext2,
                         k + 3 < kWarpChunks ? res_lane + (c + 3 * kChunkStep) * 32 : (res_nt ? res_nt + (half + kChunkStep) * 32 : nullptr));
          }
          pre_tile = res_nt != nullptr;
          fast_done = true;
        }
      }
#pragma unroll 1
      for (int c = half; c < ((kTopkEpi || fast_done) ? 0 : BLOCK_N / 32); c += kChunkStep) {
        const int col0 = n0 + c * 32;
        if (col0 >= args.N || rbase >= args.M) continue;  // warp-uniform
        if (kDirect || (kEpi != 2 && (!kResCo || args.residual == nullptr))) {
          // (the throughput configurations only take the transposed path below for fp32 residual streams: without a
          // residual to read, storing the row straight from registers measured faster)
          const int row = rbase + lane;
          float res[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) res[j] = 0.f;
          if (args.residual != nullptr && row < args.M) {
            if (kPrefetchRes && pre_ok) {
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                if (col0 + 4 * j < args.N) {
                  const float4 r4 = pre[(c * 8 + j) * 128 + et];
                  res[4 * j] = r4.x; res[4 * j + 1] = r4.y; res[4 * j + 2] = r4.z; res[4 * j + 3] = r4.w;
                }
              }
            } else if ((args.N & 3) == 0 && col0 + 32 <= args.N) {
              const float4* rp = (const float4*)(args.residual + (size_t)row * args.N + col0);  // plain loads (PDL)
#pragma unroll
              for (int j = 0; j < 8; ++j) { const float4 r4 = rp[j]; res[4 * j] = r4.x; res[4 * j + 1] = r4.y; res[4 * j + 2] = r4.z; res[4 * j + 3] = r4.w; }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) if (col0 + j < args.N) res[j] = args.residual[(size_t)row * args.N + col0 + j];
            }
          }
          uint32_t v[32];
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BLOCK_N + c * 32), v);
          float f[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
          // kEpi == 4 (hmask): the saved activation of this thread's row chunk is requested before the accumulator load
          uint4 hm[kEpi == 4 ? 4 : 1];
          if constexpr (kEpi == 4) {
            if (args.hmask != nullptr) {
#pragma unroll
              for (int j = 0; j < 4; ++j) hm[j] = make_uint4(0u, 0u, 0u, 0u);
              if (row < args.M && (args.N & 7) == 0 && col0 + 32 <= args.N) {
                const uint4* hp = (const uint4*)(args.hmask + (size_t)row * args.N + col0);
#pragma unroll
                for (int j = 0; j < 4; ++j) hm[j] = hp[j];
              } else if (row < args.M) {
                unsigned short* hs = (unsigned short*)hm;
                for (int j = 0; j < 32; ++j)
                  hs[j] = (col0 + j < args.N) ? __bfloat16_as_ushort(args.hmask[(size_t)row * args.N + col0 + j]) : (unsigned short)0;
              }
            }
          }
          epilogue_row<kEpi == 1>(args, f, res, row, col0, ln_rstd, ln_mr, sbias + c * 32, slnc + c * 32);
          if constexpr (kEpi == 4) {
            if (args.hmask != nullptr) {
              const unsigned short* hs = (const unsigned short*)hm;
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                // (0x0000 and 0x8000 are +-0: "activation was zero" = ReLU inactive or dropped)
                const float v = ((hs[j] & 0x7fffu) != 0 && row < args.M && col0 + j < args.N) ? f[j] * args.hscale : 0.f;
                f[j] = __bfloat162float(__float2bfloat16_rn(v));  // the value the consuming GEMMs (and the bias gradient) see
              }
              if (args.colsum != nullptr) {
                // column sums over the warp's 32 rows: recursive halving leaves lane j with the sum of column j
                float cs[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) cs[j] = f[j];
#pragma unroll
                for (int half_w = 16; half_w >= 1; half_w >>= 1) {
                  const bool upper = (lane & half_w) != 0;
#pragma unroll
                  for (int j = 0; j < half_w; ++j) {
                    // keep the half of the columns this lane stays responsible for, send the other half to the partner
                    const float keep = upper ? cs[j + half_w] : cs[j];
                    const float send = upper ? cs[j] : cs[j + half_w];
                    cs[j] = keep + __shfl_xor_sync(0xffffffffu, send, half_w);
                  }
                }
                // lane's column: bit-reversal-free mapping - lane l ends up owning column l (upper halves took the upper columns)
                if (col0 + lane < args.N) atomicAdd(args.colsum + col0 + lane, cs[0]);
              }
            }
          }
          if constexpr (!kDirect) {
            // bf16 output of the throughput configurations: the thread's 64-byte row piece goes through the warp's staging
            // tile (16-byte pieces XOR-swizzled by row pair) so that one store instruction writes 8 rows x 64 contiguous
            // bytes (full sectors, 8 requests) instead of 32 rows x 16 bytes - the row-mapped stores cost 30 % of the
            // GEMM's throughput by crowding the load path
            if (args.y_bf16 && (args.N & 7) == 0 && col0 + 32 <= args.N) {
              const int sw = (lane >> 1) & 3;
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const __nv_bfloat162 p0 = __floats2bfloat162_rn(f[8 * j], f[8 * j + 1]), p1 = __floats2bfloat162_rn(f[8 * j + 2], f[8 * j + 3]);
                const __nv_bfloat162 p2 = __floats2bfloat162_rn(f[8 * j + 4], f[8 * j + 5]), p3 = __floats2bfloat162_rn(f[8 * j + 6], f[8 * j + 7]);
                uint4 o;
                o.x = *(const uint32_t*)&p0; o.y = *(const uint32_t*)&p1; o.z = *(const uint32_t*)&p2; o.w = *(const uint32_t*)&p3;
                *(uint4*)(stg + lane * 64 + ((j ^ sw) << 4)) = o;
              }
              __syncwarp();
              __nv_bfloat16* yb = (__nv_bfloat16*)args.y + (size_t)split * args.split_stride + col0 + (lane & 3) * 8;
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const int rl = i * 8 + (lane >> 2);
                const uint4 o = *(const uint4*)(stg + rl * 64 + (((lane & 3) ^ ((rl >> 1) & 3)) << 4));
                if (rbase + rl < args.M) *(uint4*)(yb + (size_t)(rbase + rl) * args.N) = o;
              }
              __syncwarp();  // staging is rewritten by the next chunk
              continue;
            }
          }
          if constexpr (!kDirect && kEpi != 1) {
            // fp32 output (e.g. the training logits, 170 MB per step): same idea with the full 4 KB staging tile - one store
            // instruction writes 4 rows x 128 contiguous bytes
            if (!args.y_bf16 && (args.N & 3) == 0 && col0 + 32 <= args.N) {
#pragma unroll
              for (int j = 0; j < 8; ++j)
                *(float4*)(stg + lane * 128 + ((j ^ jsw) << 4)) = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
              __syncwarp();
              float* yf = (float*)args.y + (size_t)split * args.split_stride + col0 + (lane & 7) * 4;
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const int rl = i * 4 + (lane >> 3);
                const float4 o = *(const float4*)(stg + rl * 128 + (((lane & 7) ^ (rl & 7)) << 4));
                if (rbase + rl < args.M) *(float4*)(yf + (size_t)(rbase + rl) * args.N) = o;
              }
              __syncwarp();  // staging is rewritten by the next chunk
              continue;
            }
          }
          if (row < args.M) epilogue_store_row<kEpi == 1>(args, f, row, col0, (size_t)split * args.split_stride);
          continue;
        }
        // coalesced mapping: lane handles 16-byte piece (lane & 7) of rows (lane >> 3) + 4 i, i = 0..7
        const int pj = lane & 7;
        const int col = col0 + pj * 4;
        // the residual chunk was requested one chunk ago (the tile's first one before the accumulator was awaited)
        float4 rres[8];
        const bool has_res = kEpi != 2 && kResCo && args.residual != nullptr;
        if constexpr (kPipeRes) {
          if (has_res) {
#pragma unroll
            for (int i = 0; i < 8; ++i) rres[i] = rnext[i];
          }
        }
        // weight gradient: W, logits (and injected uniforms) of the 8 rows this lane serves, all in flight together
        float4 wg_w4[kEpi == 2 ? 8 : 1], wg_s4[kEpi == 2 ? 8 : 1], wg_u4[kEpi == 2 ? 8 : 1];
        if (kEpi == 2) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            wg_w4[i] = wg_s4[i] = wg_u4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            const int row = rbase + i * 4 + (lane >> 3);
            if ((args.N & 3) == 0 && row < args.M && col < args.N) {
              const size_t e = (size_t)row * args.N + col;
              wg_w4[i] = __ldg((const float4*)(args.wg_w + e));
              if (args.wg_s) wg_s4[i] = __ldg((const float4*)(args.wg_s + e));
              if (args.wg_u) wg_u4[i] = __ldg((const float4*)(args.wg_u + e));
            }
          }
        }
        // residual: coalesced mapping -> staging -> row mapping (16-byte piece p of row r sits at piece p ^ (r & 7)), done BEFORE the
        // accumulator load so that the next chunk's request is in flight underneath the tcgen05.ld, the element-wise stage and the stores
        // (the residual row piece is then added straight from the staging tile inside epilogue_row: no res[] registers)
        if constexpr (kEpi != 2) {
          if (has_res) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int rl = i * 4 + (lane >> 3);
              *(float4*)(stg + rl * 128 + ((pj ^ (rl & 7)) << 4)) = rres[i];
            }
            __syncwarp();
            if constexpr (kPipeRes) {
              const int cn = c + kChunkStep;  // next chunk of this warp
              if (cn < BLOCK_N / 32 && n0 + cn * 32 < args.N) load_residual8(args, rbase, lane, n0 + cn * 32 + pj * 4, rnext);
            }
          }
        }
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BLOCK_N + c * 32), v);
        if constexpr (kEpi != 2) {
          float f[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
          epilogue_row<kEpi == 1>(args, f, f, rbase + lane, col0, ln_rstd, ln_mr, sbias + c * 32, slnc + c * 32,
                                  has_res ? stg + lane * 128 : nullptr, jsw);
          __syncwarp();  // every lane has read its residual piece: the tile may take the results
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *(float4*)(stg + lane * 128 + ((j ^ jsw) << 4)) = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int rl = i * 4 + (lane >> 3);
            const float4 o = *(const float4*)(stg + rl * 128 + ((pj ^ (rl & 7)) << 4));
            if (rbase + rl < args.M && col < args.N) epilogue_store4<kEpi == 1>(args, o, rbase + rl, col, (size_t)split * args.split_stride);
          }
        } else {
          // weight gradient: transpose the raw accumulator, element-wise work in the coalesced mapping
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *(uint4*)(stg + lane * 128 + ((j ^ jsw) << 4)) = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int rl = i * 4 + (lane >> 3);
            const float4 g4 = *(const float4*)(stg + rl * 128 + ((pj ^ (rl & 7)) << 4));
            if (rbase + rl < args.M && col < args.N)
              epilogue_wgrad4(args, g4, rbase + rl, col, args.splits > 1, split == 0, wg_w4[i], wg_s4[i], wg_u4[i]);
          }
        }
        __syncwarp();  // staging is rewritten by the next chunk
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (pair) mbar_arrive_cluster(mapa_u32(smem_u32(&tmem_empty_bar[buf]), 0u));  // the leader's MMA warp owns both halves
        else mbar_arrive(&tmem_empty_bar[buf]);
      }
    }
  } else if (kMasked && warp >= kFirstTransformWarp) {
    // ===== B transform: fp32 W (+ mask logits) -> masked bf16 operand tile (swizzled K-major) =====
    sc::pdl_wait();
    const int t = threadIdx.x - 32 * kFirstTransformWarp;  // 0..127
    const int chunk = t & 15;         // float4 index inside the 64-wide k block
    const int rbase = t >> 4;         // 0..7
    const sc::Philox philox(args.seed);
    constexpr int kPasses = BLOCK_N / 8;
    int it = 0;
    for (int u = unit0; u < num_units; u += ustride) {
      const int tile = u / args.splits, split = u - tile * args.splits;
      const int n0 = (tile % args.tiles_n) * BLOCK_N;
      const int kb0 = split * args.kb_per_split;
      const int kb1 = min(kb0 + args.kb_per_split, num_kb_total);
      for (int kb = kb0; kb < kb1; ++kb, ++it) {
        const int s = it % kStages;
        const uint32_t ph = (it / kStages) & 1;
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
                const float4 uu = __ldg((const float4*)(args.uniforms + e));
                m[0] = uu.x < sc::sigmoidf_(sa[0]) ? 1.f : 0.f;
                m[1] = uu.y < sc::sigmoidf_(sa[1]) ? 1.f : 0.f;
                m[2] = uu.z < sc::sigmoidf_(sa[2]) ? 1.f : 0.f;
                m[3] = uu.w < sc::sigmoidf_(sa[3]) ? 1.f : 0.f;
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
  }

  tcgen05_fence_before();
  __syncthreads();
  if (c2) cluster_sync_all();  // the peer may still arrive on this CTA's barriers / multicast into its smem until it is done too
  if (warp == 2) {
    tcgen05_fence_after();
    if (pair) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)(2 * BLOCK_N)));
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)(2 * BLOCK_N)));
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

// bf16 [rows = K index, cols = M/N index] row-major, box = [64 rows, 64 cols], SWIZZLE_128B (MN-major operands)
int make_tmap_mn(CUtensorMap* map, const void* ptr, int rows, int cols) {
  EncodeTiledFn fn = get_encode_fn();
  SC_CHECK(fn != nullptr, SC_ERR_DRIVER, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
  cuuint32_t box[2] = {64u, (cuuint32_t)BLOCK_K};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SC_CHECK(r == CUDA_SUCCESS, SC_ERR_DRIVER, "cuTensorMapEncodeTiled (MN-major) failed with CUresult %d", (int)r);
  return SC_OK;
}

int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

template <int BLOCK_N, bool kMasked, int kStages, int kEpi, bool kMN = false, bool kPair = false, bool kResCo = false>
int launch(const CUtensorMap& ta, const CUtensorMap& tb, GemmArgs& a, int want_splits, cudaStream_t stream) {
  auto kern = sc_gemm_bf16_kernel<BLOCK_N, kMasked, kStages, kEpi, kMN, kPair, kResCo>;
  SC_CHECK(kResCo || a.residual == nullptr || num_epilogue_warps(BLOCK_N, kStages) == 4 || kEpi == 2, SC_ERR_UNSUPPORTED,
           "fp32 residual reached an 8-epilogue-warp instantiation without the coalesced residual epilogue");
  SC_CHECK(kPair == (a.cluster2 == 2), SC_ERR_UNSUPPORTED, "CTA-pair mode reached an instantiation without it (cluster2 = %d)", a.cluster2);
  constexpr int smem = Smem<BLOCK_N, kStages>::kTotal;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    SC_CHECK(e == cudaSuccess, (int)e, "cudaFuncSetAttribute(smem=%d): %s", smem, cudaGetErrorString(e));
    attr_set = true;
  }
  a.tiles_n = (a.N + BLOCK_N - 1) / BLOCK_N;
  a.num_tiles = a.tiles_n * ((a.M + BLOCK_M - 1) / BLOCK_M);
  if (a.cluster2) a.num_tiles = a.tiles_n * (((a.M + BLOCK_M - 1) / BLOCK_M + 1) / 2);  // pairs of M blocks
  const int num_kb = (a.K + BLOCK_K - 1) / BLOCK_K;
  // split K (weight gradients only: few output tiles, thousands of tokens to contract) until the SMs are covered.
  // want_splits > 0: the caller (sc_linear_wgrad with a workspace) stores one fp32 partial product per split and
  // reduces them in sc_mask_grad_reduce; otherwise the fused epilogue adds its partials with vector reductions.
  int splits = 1;
  if (a.wgrad || want_splits > 0) {
    splits = want_splits > 0 ? want_splits : (a.num_tiles >= sm_count() ? 1 : (sm_count() + a.num_tiles - 1) / a.num_tiles);
    splits = max(1, min(splits, num_kb / 4 > 0 ? num_kb / 4 : 1));
  }
  a.kb_per_split = (num_kb + splits - 1) / splits;
  a.splits = (num_kb + a.kb_per_split - 1) / a.kb_per_split;
  if (a.wgrad && a.splits > 1 && !a.accumulate) {
    if (a.dw) cudaMemsetAsync(a.dw, 0, (size_t)a.M * a.N * sizeof(float), stream);
    if (a.ds) cudaMemsetAsync(a.ds, 0, (size_t)a.M * a.N * sizeof(float), stream);
  }
  // persistent grid, never more CTAs than work units
  static int env_per_sm = -1;
  if (env_per_sm < 0) { const char* e = getenv("SC_GEMM_PER_SM"); env_per_sm = e ? atoi(e) : 0; }
  // one persistent CTA of a launch per SM: two co-resident CTAs of the SAME GEMM measured slower than one CTA looping
  // over two tiles (the small configurations still leave room for CTAs of other streams)
  const int per_sm = env_per_sm > 0 ? env_per_sm : 1;
  const int units = a.num_tiles * a.splits;
  // a.tiles_per_cta > 1 (throughput regime: several independent GEMM chains in flight): fewer persistent CTAs, each looping
  // over that many tiles - the barrier / TMEM / tensor-map prologue is paid once and the epilogue of a tile runs under the
  // main loop of the next (double-buffered accumulators) instead of holding an SM
  // SC_GEMM_MAX_CTAS (diagnostic): cap on the persistent grid - leaves SMs to the HBM-bound kernels of other streams
  static int env_cap_ctas = -1;
  if (env_cap_ctas < 0) { const char* e = getenv("SC_GEMM_MAX_CTAS"); env_cap_ctas = e ? atoi(e) : 0; }
  const int sm_budget = env_cap_ctas > 0 ? min(env_cap_ctas, sm_count()) : sm_count();
  int want = min(units, per_sm * sm_budget);
  if (a.tiles_per_cta > 1 && !a.cluster2) want = min(want, (units + a.tiles_per_cta - 1) / a.tiles_per_cta);
  dim3 grid(want);
  constexpr int threads = 32 * (4 + num_epilogue_warps(BLOCK_N, kStages) + (kMasked ? kNumTransformWarps : 0));
  cudaError_t e;
  if (a.cluster2) {
    int clusters = min(units, sm_budget / 2);
    if (a.tiles_per_cta > 1) clusters = min(clusters, max(1, (units + a.tiles_per_cta - 1) / a.tiles_per_cta));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * clusters); cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = (size_t)smem; cfg.stream = stream;
    cudaLaunchAttribute at[2];
    int na = 0;
    if (g_sc_pdl & 1) {
      at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      at[na].val.programmaticStreamSerializationAllowed = 1;
      ++na;
    }
    at[na].id = cudaLaunchAttributeClusterDimension;
    at[na].val.clusterDim.x = 2; at[na].val.clusterDim.y = 1; at[na].val.clusterDim.z = 1;
    ++na;
    cfg.attrs = at; cfg.numAttrs = na;
    e = cudaLaunchKernelEx(&cfg, kern, ta, tb, a);
  } else {
    e = sc::launch_pdl(kern, grid, dim3(threads), (size_t)smem, stream, ta, tb, a);
  }
  if (e != cudaSuccess) {
    sc_set_error("sc_gemm_bf16_kernel: launch failed: %s", cudaGetErrorString(e));
    return (int)e;
  }
  SC_LAUNCH_CHECK("sc_gemm_bf16_kernel");
  return SC_OK;
}

// (block_n, stages) -> instantiation: masked / plain weights, epilogue kind, CTA pairs, and - for the 8-epilogue-warp
// configurations only - the variant with the coalesced fp32-residual epilogue when the call has a residual
template <int BN, int ST>
int dispatch_case(const CUtensorMap& ta, const CUtensorMap& tb, GemmArgs& a, bool masked, int epi, int force_splits, cudaStream_t stream) {
  constexpr bool kWide = num_epilogue_warps(BN, ST) == 8;
  if constexpr (kWide) {
    if (a.residual != nullptr && epi != 2) {
      if (masked) return epi ? launch<BN, true, ST, 1, false, false, true>(ta, tb, a, force_splits, stream)
                             : launch<BN, true, ST, 0, false, false, true>(ta, tb, a, force_splits, stream);
      if constexpr (BN == 256) {
        if (a.cluster2 == 2)
          return epi == 1 ? launch<256, false, 3, 1, false, true, true>(ta, tb, a, force_splits, stream)
                          : launch<256, false, 3, 0, false, true, true>(ta, tb, a, force_splits, stream);
      }
      return epi == 1 ? launch<BN, false, ST, 1, false, false, true>(ta, tb, a, force_splits, stream)
                      : launch<BN, false, ST, 0, false, false, true>(ta, tb, a, force_splits, stream);
    }
  }
  if (masked) return epi ? launch<BN, true, ST, 1>(ta, tb, a, force_splits, stream)
                         : launch<BN, true, ST, 0>(ta, tb, a, force_splits, stream);
  if constexpr (BN == 256) {
    if (a.cluster2 == 2 && epi != 2)
      return epi == 1 ? launch<256, false, 3, 1, false, true>(ta, tb, a, force_splits, stream)
                      : launch<256, false, 3, 0, false, true>(ta, tb, a, force_splits, stream);
  }
  return epi == 2 ? launch<BN, false, ST, 2>(ta, tb, a, force_splits, stream)
       : epi == 1 ? launch<BN, false, ST, 1>(ta, tb, a, force_splits, stream)
                  : launch<BN, false, ST, 0>(ta, tb, a, force_splits, stream);
}

}  // namespace

// Split count the weight-gradient GEMM [N,K] = dyT [N,M] * xT[K,M]^T would use with `max_splits` partial buffers.
int sc_gemm_wgrad_splits(int N, int K, int M, int max_splits) {
  const long t128 = (long)((N + 127) / 128) * ((K + 127) / 128);
  const int bn = (t128 >= sm_count() / 4) ? 128 : 64;
  const long tiles = (long)((N + 127) / 128) * ((K + bn - 1) / bn);
  const int num_kb = (M + BLOCK_K - 1) / BLOCK_K;
  // as many splits as keep all work units in ONE wave, at most 2: alone, up to 4 splits are faster for the smallest weights
  // (scripts/wgrad_sweep.py), but the weight gradients run on a side stream BESIDE the backward's dX chain - fewer, longer CTAs
  // leave that chain its SMs (whole SMP step, SC_WGRAD_MAX_SPLITS = 4 / 3 / 2 / 1: 5.35 / 5.33 / 5.28 / 5.49 ms)
  static int env_cap = -1;
  if (env_cap < 0) { const char* e = getenv("SC_WGRAD_MAX_SPLITS"); env_cap = e ? atoi(e) : 0; }
  int splits = (int)(sm_count() / tiles);
  splits = max(1, min(splits, env_cap > 0 ? env_cap : 2));
  splits = max(1, min(splits, num_kb / 4 > 0 ? num_kb / 4 : 1));
  splits = min(splits, max(1, max_splits));
  const int per = (num_kb + splits - 1) / splits;
  return (num_kb + per - 1) / per;
}

// Internal entry used by sc_linear (sc_api.cu).
int sc_gemm_bf16_launch(const void* x, const void* w, int w_dtype, const float* mask, int mask_mode, const float* uniforms,
                        unsigned long long seed, unsigned long long stream_id, const float* bias, const float* residual,
                        void* y, int y_dtype, int M, int N, int K, int relu, int block_n, const ScGemmExtra* ex,
                        cudaStream_t stream) {
  SC_CHECK(M > 0 && N > 0 && K > 0, SC_ERR_SHAPE, "sc_linear: empty problem M=%d N=%d K=%d", M, N, K);
  SC_CHECK(K % 8 == 0 || (ex && ex->mn_major), SC_ERR_SHAPE, "sc_linear(bf16): K=%d must be a multiple of 8 (16-byte TMA rows)", K);
  SC_CHECK(((uintptr_t)x & 15) == 0 && ((uintptr_t)w & 15) == 0 && ((uintptr_t)y & 15) == 0, SC_ERR_ALIGN,
           "sc_linear(bf16): x, w, y must be 16-byte aligned");
  SC_CHECK(y_dtype == SC_F32 || y_dtype == SC_BF16, SC_ERR_DTYPE, "sc_linear: bad y dtype %d", y_dtype);
  const bool wgrad = ex && ex->wgrad;
  const bool masked = (w_dtype == SC_F32);
  SC_CHECK(!(masked && wgrad), SC_ERR_UNSUPPORTED, "sc_linear_wgrad: operands must be in the activation dtype");
  SC_CHECK(masked || mask_mode == SC_MASK_NONE || wgrad, SC_ERR_DTYPE,
           "sc_linear(bf16): mask modes need fp32 master weights (bf16 weights are expected pre-masked)");
  if (masked && mask_mode != SC_MASK_NONE) {
    SC_CHECK(mask != nullptr && ((uintptr_t)mask & 15) == 0, SC_ERR_ALIGN, "sc_linear: mask must be 16-byte aligned");
    SC_CHECK(mask_mode != SC_MASK_UNIFORM || uniforms != nullptr, SC_ERR_SHAPE, "sc_linear: uniforms missing");
    SC_CHECK(K % 4 == 0, SC_ERR_SHAPE, "K %% 4");
  }
  if (ex && (ex->ln_stats || ex->stats_out || ex->y2)) {
    SC_CHECK(!wgrad, SC_ERR_UNSUPPORTED, "sc_linear: LayerNorm folding is a forward feature");
    SC_CHECK(!ex->ln_stats || (K % 32 == 0 && ex->ln_c != nullptr && ((uintptr_t)ex->ln_c & 15) == 0), SC_ERR_SHAPE,
             "sc_linear_ln: K=%d must be a multiple of 32 and ln_c 16-byte aligned", K);
    SC_CHECK(!(ex->stats_out || ex->y2) || N % 32 == 0, SC_ERR_SHAPE, "sc_linear_ln: N=%d must be a multiple of 32 to emit row statistics", N);
    SC_CHECK(!ex->y2 || (y_dtype == SC_F32 && ((uintptr_t)ex->y2 & 15) == 0), SC_ERR_DTYPE, "sc_linear_ln: the bf16 copy needs an fp32 y");
  }
  const int sms = sm_count();
  const long mt = (M + 127) / 128;
  int force_stages = 0, force_splits = 0, tiles_per_cta = 0;
  {
    static int env_tpc = -1;
    if (env_tpc < 0) { const char* e = getenv("SC_GEMM_TPC"); env_tpc = e ? atoi(e) : 0; }
    if (block_n >= 10000000) { tiles_per_cta = block_n / 10000000; block_n %= 10000000; }  // hint: + 10^7 * tiles per CTA
    else if (env_tpc > 1 && mt <= 24) tiles_per_cta = env_tpc;                                // diagnostic: decode-sized problems
  }
  if (block_n >= 100000) {  // tuning hook: tile_n = 100000 * splits + 1000 * stages + block_n
    force_splits = block_n / 100000;
    block_n %= 100000;
  }
  if (block_n >= 1000) {
    force_stages = block_n / 1000;
    block_n %= 1000;
  }
  if (block_n == 0) {
    // widest tile that still gives most SMs a tile (scripts/gemm_sweep.py, in-graph, L2-warm, round 1 numbers in DESIGN.md)
    const long t256 = mt * ((N + 255) / 256), t128 = mt * ((N + 127) / 128);
    block_n = (t256 >= sms && !masked) ? 256 : (t128 >= 100 ? 128 : 64);
    if (wgrad || (ex && ex->partial_splits > 0)) block_n = (t128 >= sms / 4) ? 128 : 64;
    if (N <= 64) block_n = 64;
  }
  const bool topk = ex && ex->topk_part != nullptr;
  if (topk) {
    SC_CHECK(!masked && !wgrad && !residual && !relu && !(ex->mn_major) && !ex->ln_stats && !ex->y2 && !ex->stats_out && ex->dropout_p == 0.f,
             SC_ERR_UNSUPPORTED, "sc_linear_topk: plain bf16 x / bf16 w / bias only");
    block_n = 256;
  }
  const bool mn = ex && ex->mn_major;
  static int env_mc = -1;
  if (env_mc < 0) { const char* e = getenv("SC_GEMM_MULTICAST"); env_mc = e ? atoi(e) : 2; }
  // SC_GEMM_MULTICAST: 0 = no clusters, 1 = 2-CTA clusters with a multicast B tile (two cta_group::1 MMAs; measured: no gain),
  // 2 (default) / 3 = CTA pairs: one tcgen05.mma.cta_group::2 of M = 256, each CTA stages its A block and HALF of the B tile
  // (+5..20 % on the encoder GEMMs together with the coalesced staged stores)
  const bool c2 = env_mc > 0 && block_n == 256 && !masked && !mn && !wgrad && !(ex && ex->partial_splits > 0) && force_splits == 0 &&
                  mt >= 2 && !(ex && ex->hmask) && !(env_mc >= 2 && topk) &&
                  // default (2): pairs for the many-wave problems (the encoder's M = 18432 GEMMs), where the main loop is what
                  // bounds the launch; at M = 4250 / 1536 the whole step measured the same or slightly slower.  3 = every mt >= 2
                  !(env_mc == 2 && mt < 64);
  CUtensorMap ta, tb;
  int rc;
  if (mn) {
    SC_CHECK(!masked && !wgrad && M % 8 == 0 && N % 8 == 0, SC_ERR_SHAPE,
             "MN-major operands: bf16, M=%d and N=%d must be multiples of 8 (16-byte rows)", M, N);
    if (block_n > 128) block_n = 128;
    rc = make_tmap_mn(&ta, x, K, M);
    if (rc) return rc;
    rc = make_tmap_mn(&tb, w, K, N);
    if (rc) return rc;
  } else {
    rc = make_tmap(&ta, x, M, K, BLOCK_M);
    if (rc) return rc;
    if (!masked) {
      rc = make_tmap(&tb, w, N, K, c2 ? block_n / 2 : block_n);
      if (rc) return rc;
    } else {
      tb = ta;
    }
  }
  GemmArgs a;
  memset(&a, 0, sizeof(a));
  a.M = M; a.N = N; a.K = K;
  a.pdl_early = (g_sc_pdl & 4) ? 1 : 0;
  { static int env_dbg = -1; if (env_dbg < 0) { const char* e = getenv("SC_GEMM_DBG"); env_dbg = e ? atoi(e) : 0; } a.dbg = env_dbg; }
  a.tiles_per_cta = tiles_per_cta;
  a.cluster2 = c2 ? (env_mc >= 2 ? 2 : 1) : 0;  // 1: multicast B, two cta_group::1 MMAs; 2: one cta_group::2 MMA per pair (default)
  a.w32 = masked ? (const float*)w : nullptr;
  a.mask = mask; a.uniforms = uniforms; a.mask_mode = mask_mode;
  a.seed = seed; a.stream_id = stream_id;
  a.bias = bias; a.residual = residual; a.y = y; a.y_bf16 = (y_dtype == SC_BF16); a.relu = relu;
  if (ex) {
    a.dropout_p = ex->dropout_p; a.drop_seed = ex->drop_seed; a.drop_stream = ex->drop_stream;
    a.wgrad = ex->wgrad; a.bypass = ex->bypass; a.sp_coeff = ex->sp_coeff; a.accumulate = ex->accumulate;
    a.wg_w = ex->wg_w; a.wg_s = ex->wg_s; a.wg_u = ex->wg_u; a.dw = ex->dw; a.ds = ex->ds;
    a.ln_stats = ex->ln_stats; a.ln_c = ex->ln_c; a.ln_eps = ex->ln_eps; a.y2 = ex->y2; a.stats_out = ex->stats_out;
    a.topk_part = ex->topk_part;
    a.hmask = (const __nv_bfloat16*)ex->hmask; a.hscale = ex->hscale; a.colsum = ex->colsum;
    if (ex->partial_splits > 0) {
      SC_CHECK(!wgrad && y_dtype == SC_F32 && !bias && !residual && !relu, SC_ERR_UNSUPPORTED, "split-K partial products are plain fp32 tiles");
      force_splits = ex->partial_splits;
      a.split_stride = ex->split_stride;
    }
  }
  if (a.hmask) {
    SC_CHECK(!masked && !wgrad && !residual && !bias && !relu && y_dtype == SC_BF16 && !(ex->mn_major), SC_ERR_UNSUPPORTED,
             "sc_linear_hmask: bf16 operands, bf16 output, no bias / residual / ReLU");
  }
  const int epi = a.wgrad ? 2 : ((a.dropout_p > 0.f || a.ln_stats || a.y2 || a.stats_out) ? 1 : 0);
  if (a.hmask) {
    // own epilogue kind (4) so that the dropout / LayerNorm epilogue keeps its register budget; 3-stage ring
    if (block_n == 256) return launch<256, false, 3, 4>(ta, tb, a, force_splits, stream);
    if (block_n == 128) return launch<128, false, 3, 4>(ta, tb, a, force_splits, stream);
    return launch<64, false, 3, 4>(ta, tb, a, force_splits, stream);
  }
  if (topk) return ex->topk_n <= 3 ? launch<256, false, 3, 5>(ta, tb, a, force_splits, stream)
                                   : launch<256, false, 3, 3>(ta, tb, a, force_splits, stream);
  if (mn) {
    SC_CHECK(epi == 0, SC_ERR_UNSUPPORTED, "MN-major operands serve the plain epilogue only");
    return block_n == 128 ? launch<128, false, 3, 0, true>(ta, tb, a, force_splits, stream)
                          : launch<64, false, 3, 0, true>(ta, tb, a, force_splits, stream);
  }
#define SC_GEMM_CASE(BN, ST)                                                                                          \
  if (block_n == BN && stages == ST) return dispatch_case<BN, ST>(ta, tb, a, masked, epi, force_splits, stream);
  int stages = force_stages;
  if (stages == 0) stages = (block_n == 128 && (long)mt * ((N + 127) / 128) > 2L * sms) ? 5 : 3;  // 128x5: multi-wave throughput config
  SC_GEMM_CASE(64, 3); SC_GEMM_CASE(64, 4); SC_GEMM_CASE(64, 6);
  SC_GEMM_CASE(128, 3); SC_GEMM_CASE(128, 5);
  SC_GEMM_CASE(256, 3);
#undef SC_GEMM_CASE
  SC_CHECK(false, SC_ERR_UNSUPPORTED, "sc_linear(bf16): tile %d x %d stages not instantiated", block_n, stages);
  return SC_OK;
}
