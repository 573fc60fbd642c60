// fdm_gemm_bf16: persistent, warp-specialised tcgen05/TMEM GEMM fed by TMA (sm_100a).
//
//   C[M,N] = act(A[M,K] * W[N,K]^T + bias) + residual          bf16 operands, fp32 accumulate
//
// Both operands are K-major (activation rows and nn.Linear weights), so A and W tiles are loaded by
// TMA straight into the 128B-swizzled K-major layout tcgen05.mma consumes; no transposes anywhere.
// Implicit 1-D convolutions (VQ-decoder Conv1d k5, HuBERT conv stack / positional conv) are the same
// kernel: the producer shifts the TMA row coordinate per filter tap instead of materialising im2col.
//
// CTA layout (320 threads, 1 CTA/SM, persistent over a static tile schedule):
//   warp 0      TMA producer   (one elected lane)         smem ring: STAGES x {A 128x64, W BLOCK_Nx64}
//   warp 1      MMA issuer     (one elected lane)         tcgen05.mma 128 x BLOCK_N x 16, cta_group::1
//   warps 2-9   epilogue       (TMEM lane quadrant = warp & 3, column half = (warp-2)/4), double-buffered TMEM accumulator
// Pipelines: full/empty mbarriers (TMA <-> MMA), tmem_full/tmem_empty mbarriers (MMA <-> epilogue).
#include "tc_common.cuh"
#include <stdlib.h>
#include <string.h>

using namespace tc;

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;  // 64 bf16 = 128 bytes = one SWIZZLE_128B row
constexpr int UMMA_K = 16;
constexpr int NUM_THREADS = 320;  // warp 0 TMA, warp 1 MMA, warps 2-9 epilogue
constexpr int EPI_WARPS = 8;
constexpr int ACC_STAGES = 2;
constexpr int NB_STAGE = 2;  // epilogue staging tiles per warp
constexpr int IDENT_COPIES = 256;  // copies of the 64 x 64 identity in global memory (RESMMA): every CTA reads its own copy, so
                                   // the loads spread over the L2 slices instead of hammering the 64 lines of a single one

// CG = 1: one CTA per 128 x BLOCK_N tile (tcgen05 cta_group::1).
// CG = 2: a CTA pair (cluster of 2 SMs) per 256 x BLOCK_N tile (cta_group::2): each CTA stages its own 128 rows of A and
//         HALF of the W tile; one MMA issued by the leader CTA reads both CTAs' shared memory and writes both CTAs'
//         TMEM. Per output element that halves the W traffic from L2 and shared memory and deepens the ring
//         (6 stages instead of 4) - the GEMMs are power-capped, so fewer bytes moved is more FLOP/s.
template <int BLOCK_N, int CG = 1>
struct Cfg {
  static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;
  static constexpr int B_BYTES = (BLOCK_N / CG) * BLOCK_K * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  // ring depth: ~64 B/clk/SM of operand traffic x ~2.5k cycles of TMA latency wants ~160 KB in flight
  static constexpr int STAGES = (BLOCK_N / CG) == 256 ? 3 : ((BLOCK_N / CG) == 128 ? 5 : 6);
  static constexpr int TMEM_COLS = ACC_STAGES * BLOCK_N;  // 512 / 256 / 128: powers of two >= 32
  static constexpr int BAR_BYTES = 1024;                  // barriers + TMEM slot, keeps the staging tiles 1024-aligned
  static constexpr int STAGING_BYTES = EPI_WARPS * NB_STAGE * 4096;  // per warp: NB_STAGE tiles of 32 rows x 128 B
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + BAR_BYTES + STAGING_BYTES + 1024;  // +1024: manual alignment
};

struct Epilogue {
  const float* bias;
  const void* residual;
  void* C;
  int64_t ldr, ldc;
  int32_t res_dtype, out_dtype, act;
  int32_t tma_c, vec_r, vec_bias;  // C through TMA stores; 16-byte vector access legal for residual / bias
  int32_t tma_r;                   // bf16 residual tiles fetched by TMA into the staging tile (needs tma_c, bf16 out)
  int32_t lin_c;                   // fp32 C whose row pitch is not a multiple of 16 bytes (N = 15069, 70110: the vertex maps)
  int32_t M;
  // LayerNorm folding (see fdm_gemm_args): consumer correction, residual rebuilt from un-normalised rows, output statistics
  const float* a_ln;
  const float* w_colsum;
  const float* res_ln;
  const float* res_gamma;
  const float* res_beta;
  float* stats_out;
  int32_t stats_parts;
};

// tcgen05 / TMEM / TMA / mbarrier PTX wrappers: tc_common.cuh (shared by every tensor-core kernel of the library)
__device__ __forceinline__ uint64_t make_kmajor_sw128_desc(uint32_t smem_addr) { return make_desc_sw128(smem_addr, 16, 1024); }
__host__ __device__ constexpr uint32_t make_idesc_bf16_f32(int m, int n) { return make_idesc_bf16(m, n, 0); }

// ---- epilogue ----------------------------------------------------------------------------------------
// One thread owns one accumulator row (TMEM lane). A chunk is 128 bytes of output per row (64 bf16 or 32 fp32
// columns): bias / activation / residual are applied in registers, the chunk is written row-wise into a per-warp
// 32-row x 128-byte staging tile in the TMA 128B-swizzle layout (conflict-free 16-byte shared stores), and one
// elected lane hands the tile to the TMA store engine (cp.async.bulk.tensor, clipped at the M/N edges). Two staging
// tiles per warp let the store of chunk c overlap the math of chunk c+1. Row-per-thread global stores would cost
// 32 LSU wavefronts per instruction; this path costs none.
// bf16-path activations: the outputs are rounded to bf16 (or feed a bf16 GEMM), so the fast exponential / division /
// tanh units are accurate enough, and the epilogue - the scarce resource of these GEMMs - sheds ~20 instructions per
// element (latent-encoder GEMM with Mish: 79 -> 5x us, it ran at 335 TFLOP/s).
__device__ __forceinline__ float act_mish_fast(float x) {
  // x tanh(softplus(x)) = x n / (n + 2), n = e (e + 2), e = exp(x). Branch-free: beyond x = 20 the ratio is exactly 1.0f, so
  // the exponent is clamped there (a per-element `if (x > 20) return x` put a convergence barrier around every element of
  // the epilogue: the Mish GEMM ran at half the speed of the plain one)
  const float e = __expf(fminf(x, 20.f));
  const float n = e * (e + 2.f);
  return x * __fdividef(n, n + 2.f);
}
__device__ __forceinline__ float act_gelu_tanh_fast(float x) {
  const float inner = 0.7978845608028654f * (x + 0.044715f * x * x * x);
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(inner));
  return x * (0.5f * (1.f + t));
}
// erf by Abramowitz & Stegun 7.1.26 (|error| < 1.5e-7): one reciprocal, one exponential and a degree-5 Horner form instead
// of erff()'s ~30 branchy instructions (HuBERT's FFN GEMMs, N = 4096, run their epilogue on every element)
__device__ __forceinline__ float act_gelu_erf_fast(float x) {
  const float z = fabsf(x) * 0.70710678118654752440f;
  const float t = __fdividef(1.f, fmaf(0.3275911f, z, 1.f));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  const float erf_abs = 1.f - p * t * __expf(-z * z);
  return 0.5f * x * (1.f + copysignf(erf_abs, x));
}
// act_gelu_erf_poly (common.cuh): the MUFU-free erf-GELU for outputs that are rounded to bf16 (plain bf16 GEMMs only; fp32 /
// split-bf16 outputs keep the 1.5e-7 form above). HuBERT's FFN1 GEMMs (N = 4096, K = 1024) had their epilogue, not the MMAs, on
// the critical path (584 vs 813 TFLOP/s for the QKV GEMM).
constexpr int ACT_GELU_ERF_BF16 = 100;  // internal: FDM_ACT_GELU_ERF of a plain bf16-output GEMM
__device__ __forceinline__ void act_inplace(float (&v)[32], int act) {
  if (act == FDM_ACT_NONE) return;
  if (act == FDM_ACT_RELU) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
  } else if (act == FDM_ACT_MISH) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = act_mish_fast(v[j]);
  } else if (act == FDM_ACT_GELU_ERF) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = act_gelu_erf_fast(v[j]);
  } else if (act == ACT_GELU_ERF_BF16) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = act_gelu_erf_poly(v[j]);
  } else if (act == FDM_ACT_GELU_TANH) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = act_gelu_tanh_fast(v[j]);
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = v[j] > 0.f ? v[j] : 0.2f * v[j];
  }
}

// bias + activation + residual for 32 consecutive columns [col0, col0+32) of output row `row`
// FOLD = false compiles the LayerNorm-folding branches out: the plain kernels are instruction-for-instruction what they
// were before the feature (launch-bound problems such as MEAD's d = 512 GEMMs lost 10 % to the extra uniform branches).
template <bool FOLD>
__device__ __forceinline__ void epilogue_math(float (&v)[32], const Epilogue& ep, int64_t row, bool row_ok, int col0, int N,
                                              float a_mean = 0.f, float a_rstd = 1.f) {
  const int ncols = min(32, N - col0);
  const bool full = ncols == 32;
  if (FOLD && ep.a_ln) {  // A held un-normalised rows: C = rstd (acc - mean * colsum(W')) (+ bias' below)
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (j < ncols) v[j] = a_rstd * fmaf(-a_mean, __ldg(ep.w_colsum + col0 + j), v[j]);
  }
  if (ep.bias) {
    if (full && ep.vec_bias) {
      const float4* b4 = reinterpret_cast<const float4*>(ep.bias + col0);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 b = __ldg(b4 + j);
        v[4 * j] += b.x; v[4 * j + 1] += b.y; v[4 * j + 2] += b.z; v[4 * j + 3] += b.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < ncols) v[j] += __ldg(ep.bias + col0 + j);
    }
  }
  act_inplace(v, ep.act);
  if (ep.residual && row_ok) {
    if (ep.res_dtype == FDM_BF16) {
      const __nv_bfloat16* r = reinterpret_cast<const __nv_bfloat16*>(ep.residual) + row * ep.ldr + col0;
      if (full && ep.vec_r) {
        const uint4* r4 = reinterpret_cast<const uint4*>(r);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint4 u = __ldg(r4 + j);
          const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float2 f = __bfloat1622float2(h[q]);
            v[8 * j + 2 * q] += f.x; v[8 * j + 2 * q + 1] += f.y;
          }
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (j < ncols) v[j] += __bfloat162float(r[j]);
      }
    } else {
      const float* r = reinterpret_cast<const float*>(ep.residual) + row * ep.ldr + col0;
      if (full && ep.vec_r) {
        const float4* r4 = reinterpret_cast<const float4*>(r);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 b = __ldg(r4 + j);
          v[4 * j] += b.x; v[4 * j + 1] += b.y; v[4 * j + 2] += b.z; v[4 * j + 3] += b.w;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (j < ncols) v[j] += r[j];
      }
    }
  }
}

__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
// staging tile: 32 rows x 128 bytes, 16-byte chunk index XOR (row & 7) == CU_TENSOR_MAP_SWIZZLE_128B
__device__ __forceinline__ uint32_t stage_off(int row, int chunk) { return static_cast<uint32_t>(row * 128 + ((chunk ^ (row & 7)) << 4)); }


// ---- the kernel -----------------------------------------------------------------------------------
// Tensor maps of the low halves of split-bf16 operands. The plain kernels carry an empty struct: a producer that picks its
// tensor map through a run-time pointer select (even one that always resolves to the same map) issues TMA loads measurably
// slower - the QKV GEMM of the VOCASET step went from 111.6 to 124.2 us (A/B on one box, profiles/r02_b_gemm_ab.md) - so
// SPLIT is a template parameter and the plain instantiation is instruction-for-instruction the single-map kernel.
template <bool SPLIT> struct LoMaps {};
template <> struct LoMaps<true> { CUtensorMap a, b; };
// RESMMA: the bf16 residual is added BY THE TENSOR CORE. After the K loop of a tile the producer streams the residual
// tile through the operand ring as BLOCK_N / 64 extra k-blocks (A slot: residual[128 rows, 64 columns], B slot: a 64 x 64
// identity from a tiny L2-resident matrix) and the issuer accumulates each into its 64-column slice of the TMEM
// accumulator with N = 64 MMAs: D[:, 64 j + n] += sum_k R[:, 64 j + k] I[n, k]. bf16 x 1.0 is exact in the fp32
// accumulator, so the result equals the epilogue add up to fp32 summation order, the tile costs one k-block equivalent
// more of MMA time (+6 % at K = 1024), and the epilogue is the plain one: the TMA-residual epilogue took 11-16k cycles
// per 256 x 256 tile against 8k cycles of MMA (profiles/r01_j_gemm_tc_full.md: tensor pipe 60 % on the out-projection).
template <bool RESMMA> struct ResMaps {};
template <> struct ResMaps<true> { CUtensorMap res, ident; };
// TAILK: split-K for the tiles of a partially filled last wave. A static schedule of `tiles` full tiles on `slots` CTAs (or
// CTA pairs) leaves slots - rem of them idle during the last wave when rem = tiles mod slots is small (N = 1024 at
// M = 25344: 396 tiles = 5 waves + 26 on 74 pairs; BIWI at 16 clips per GPU: 76 tiles = 1 wave + 2). With TAILK the last
// rem tiles are cut along K into S = min(slots / rem, 8, k-blocks / 2) slices, one per CTA (pair), all running in the
// last wave. A slice writes its fp32 partial accumulator into its own slab of a caller-provided workspace (plain TMA
// stores: no zero-initialised state, no atomics on data) and bumps a counter per (tile, CTA, epilogue warp) region; the
// warp that arrives LAST at a region sums the S slabs in slice order (deterministic), applies the epilogue and stores C.
// No CTA ever waits for another one. Plain TMA-store epilogue only (no TMA residual, folding, split operands, RESMMA).
template <bool TAILK> struct TailK {};
template <> struct TailK<true> {
  CUtensorMap ws;      // fp32 [rem * S * 128 * CG, BLOCK_N] slabs, box 32 x 32
  const float* slabs;  // the same memory for the finaliser's loads
  int* counters;       // [rem][CG][EPI_WARPS], zero between launches
  int num_full, rem, S, kb_per_slice;
};

template <int BLOCK_N, int CG, bool FOLD, bool SPLIT, bool RESMMA, bool TAILK = false>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
               const __grid_constant__ CUtensorMap tmap_c, const __grid_constant__ CUtensorMap tmap_r,
               const __grid_constant__ LoMaps<SPLIT> lo, const __grid_constant__ ResMaps<RESMMA> rm,
               const __grid_constant__ TailK<TAILK> tk, const Epilogue ep,
               const int M, const int N, const int num_k_blocks, const int kb_per_tap, const int tap_row_shift, const int m_tiles,
               const int n_tiles, const int kb_per_seg, const int a_group_cols) {
  // num_k_blocks = segments x kb_per_seg. One segment: the plain bf16 GEMM. Three segments (SPLIT, split-bf16 operands): the
  // same K range three times into the same accumulator - A_lo W_hi, A_hi W_lo, A_hi W_hi (small terms first) - only the
  // producer's choice of tensor map differs, the MMA issuer and the epilogue see one long K loop.
  using C = Cfg<BLOCK_N, CG>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;  // SWIZZLE_128B tiles need 1024-byte alignment
  uint8_t* smem = smem_raw + (base - raw_addr);

  const uint32_t bar_base = base + C::STAGES * C::STAGE_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (C::STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * C::STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * C::STAGES + ACC_STAGES + a); };
  auto res_bar = [&](int w, int b) { return bar_base + 8u * (2 * C::STAGES + 2 * ACC_STAGES + 1 + NB_STAGE * w + b); };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + C::STAGES * C::STAGE_BYTES + 8 * (2 * C::STAGES + 2 * ACC_STAGES));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_tiles = m_tiles * n_tiles;                 // tiles of (128 * CG) x BLOCK_N
  const uint32_t cta_rank = CG == 2 ? cluster_ctarank() : 0u;  // 0 = leader of the pair
  const int tile_first = blockIdx.x / CG, tile_step = gridDim.x / CG;
  const int row_in_tile = static_cast<int>(cta_rank) * BLOCK_M;
  // TAILK: full tiles are [0, full_tiles); CTA (pair) p < rem * S then takes slice p % S of tile full_tiles + p / S
  int full_tiles = num_tiles, tail_tile = -1, tail_slab = 0, tail_kb0 = 0, tail_kb1 = 0;
  if constexpr (TAILK) {
    full_tiles = tk.num_full;
    if (tile_first < tk.rem * tk.S) {
      const int t = tile_first / tk.S, sl = tile_first - t * tk.S;
      tail_tile = tk.num_full + t;
      tail_slab = tile_first;  // = t * S + sl
      tail_kb0 = sl * tk.kb_per_slice;
      tail_kb1 = min(tail_kb0 + tk.kb_per_slice, num_k_blocks);
    }
  }

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    if (ep.tma_c) tma_prefetch_desc(&tmap_c);
    if (ep.tma_r) tma_prefetch_desc(&tmap_r);
    if constexpr (SPLIT) { tma_prefetch_desc(&lo.a); tma_prefetch_desc(&lo.b); }
    if constexpr (RESMMA) { tma_prefetch_desc(&rm.res); tma_prefetch_desc(&rm.ident); }
    if constexpr (TAILK) tma_prefetch_desc(&tk.ws);
  } else if (warp == 1 && lane == 0) {
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(full_bar(s), CG);  // CG = 2: leader's expect_tx arrival + the peer's remote arrival
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < ACC_STAGES; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), EPI_WARPS * 32 * CG);  // epilogue threads of both CTAs arrive on the leader's barrier
    }
    for (int w = 0; w < EPI_WARPS; ++w) {
      for (int b = 0; b < NB_STAGE; ++b) mbar_init(res_bar(w, b), 1);
    }
    fence_barrier_init();
  } else if (warp == 2) {
    if (CG == 2) tmem_alloc_2sm(smem_u32(const_cast<uint32_t*>(tmem_slot)), C::TMEM_COLS);
    else tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_slot)), C::TMEM_COLS);
  }
  tcgen05_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();  // both CTAs' barriers are initialised before any remote arrive / multicast commit
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();  // everything above overlapped the previous kernel's tail; operands are valid from here on

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer =====
      int stage = 0;
      uint32_t phase = 0;
      int tile_done_tail = 0;
      for (int tile = tile_first; tile < full_tiles || (TAILK && tile_done_tail == 0 && tail_tile >= 0); tile += tile_step) {
        int kb_begin = 0, kb_end = num_k_blocks;
        if (TAILK && tile >= full_tiles) { tile = tail_tile; kb_begin = tail_kb0; kb_end = tail_kb1; tile_done_tail = 1; }
        const int m_blk = tile / n_tiles, n_blk = tile - m_blk * n_tiles;
        const int a_row = m_blk * (BLOCK_M * CG) + row_in_tile;
        const int b_row = n_blk * BLOCK_N + static_cast<int>(cta_rank) * (BLOCK_N / CG);
        const int a_col0 = n_blk * a_group_cols;  // grouped convolution: output tile n_blk reads its own channel group of A
        for (int kb = kb_begin; kb < kb_end; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          const CUtensorMap* ma = &tmap_a;
          const CUtensorMap* mb = &tmap_b;
          int ks = kb;  // k-block inside the segment
          if constexpr (SPLIT) {
            const int seg = kb / kb_per_seg;
            ks = kb - seg * kb_per_seg;
            if (seg == 0) ma = &lo.a;
            if (seg == 1) mb = &lo.b;
          }
          const int tap = ks / kb_per_tap;
          const int kc = ks - tap * kb_per_tap;
          const uint32_t sa = base + stage * C::STAGE_BYTES;
          if (CG == 1) {
            mbar_expect_tx(full_bar(stage), C::STAGE_BYTES);
            tma_load_2d(sa, ma, full_bar(stage), a_col0 + kc * BLOCK_K, a_row + tap * tap_row_shift);
            tma_load_2d(sa + C::A_BYTES, mb, full_bar(stage), ks * BLOCK_K, b_row);
          } else {
            // both CTAs' bytes complete on the leader's barrier; the leader posts the whole transaction count
            if (cta_rank == 0) mbar_expect_tx(full_bar(stage), 2 * C::STAGE_BYTES);
            tma_load_2d_2sm(sa, ma, full_bar(stage), a_col0 + kc * BLOCK_K, a_row + tap * tap_row_shift);
            tma_load_2d_2sm(sa + C::A_BYTES, mb, full_bar(stage), ks * BLOCK_K, b_row);
            if (cta_rank != 0) mbar_arrive_remote(full_bar(stage), 0);
          }
          if (++stage == C::STAGES) { stage = 0; phase ^= 1u; }
        }
        if constexpr (RESMMA) {  // residual tile as BLOCK_N / 64 identity k-blocks
          constexpr uint32_t RES_TX = C::A_BYTES + (64 / CG) * 128;
          for (int j = 0; j < BLOCK_N / 64; ++j) {
            const int col = n_blk * BLOCK_N + 64 * j;
            if (col >= N) break;
            mbar_wait(empty_bar(stage), phase ^ 1u);
            const uint32_t sa = base + stage * C::STAGE_BYTES;
            if (CG == 1) {
              mbar_expect_tx(full_bar(stage), RES_TX);
              tma_load_2d(sa, &rm.res, full_bar(stage), col, a_row);
              tma_load_2d(sa + C::A_BYTES, &rm.ident, full_bar(stage), 0, static_cast<int>(blockIdx.x % IDENT_COPIES) * 64);
            } else {
              if (cta_rank == 0) mbar_expect_tx(full_bar(stage), 2 * RES_TX);
              tma_load_2d_2sm(sa, &rm.res, full_bar(stage), col, a_row);
              tma_load_2d_2sm(sa + C::A_BYTES, &rm.ident, full_bar(stage), 0,
                              static_cast<int>(blockIdx.x % IDENT_COPIES) * 64 + static_cast<int>(cta_rank) * 32);
              if (cta_rank != 0) mbar_arrive_remote(full_bar(stage), 0);
            }
            if (++stage == C::STAGES) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && cta_rank == 0) {
      // ===== MMA issuer (leader CTA only when CG = 2) =====
      constexpr uint32_t idesc = make_idesc_bf16_f32(BLOCK_M * CG, BLOCK_N);
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      int tail_done = 0;
      for (int tile = tile_first; tile < full_tiles || (TAILK && tail_done == 0 && tail_tile >= 0); tile += tile_step) {
        int kb_begin = 0, kb_end = num_k_blocks;
        if (TAILK && tile >= full_tiles) { tile = tail_tile; kb_begin = tail_kb0; kb_end = tail_kb1; tail_done = 1; }
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);  // epilogue has drained this accumulator
        tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BLOCK_N;
        for (int kb = kb_begin; kb < kb_end; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tcgen05_fence_after();
          const uint32_t sa = base + stage * C::STAGE_BYTES;
          const uint64_t adesc = make_kmajor_sw128_desc(sa);
          const uint64_t bdesc = make_kmajor_sw128_desc(sa + C::A_BYTES);
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            // advance 16 elements (32 bytes) along K inside the 128-byte swizzle atom: +2 in (addr >> 4) units
            if (CG == 2) umma_bf16_2sm(tmem_d, adesc + 2u * k, bdesc + 2u * k, idesc, ((kb - kb_begin) | k) != 0 ? 1u : 0u);
            else umma_bf16(tmem_d, adesc + 2u * k, bdesc + 2u * k, idesc, ((kb - kb_begin) | k) != 0 ? 1u : 0u);
          }
          // frees the smem slot (in both CTAs when CG = 2) once these MMAs retire
          if (CG == 2) umma_commit_2sm(empty_bar(stage));
          else umma_commit(empty_bar(stage));
          if (++stage == C::STAGES) { stage = 0; phase ^= 1u; }
        }
        if constexpr (RESMMA) {
          constexpr uint32_t idesc64 = make_idesc_bf16_f32(BLOCK_M * CG, 64);
          const int n_blk = tile % n_tiles;
          for (int j = 0; j < BLOCK_N / 64; ++j) {
            if (n_blk * BLOCK_N + 64 * j >= N) break;
            mbar_wait(full_bar(stage), phase);
            tcgen05_fence_after();
            const uint32_t sa = base + stage * C::STAGE_BYTES;
            const uint64_t adesc = make_kmajor_sw128_desc(sa);
            const uint64_t bdesc = make_kmajor_sw128_desc(sa + C::A_BYTES);
#pragma unroll
            for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
              if (CG == 2) umma_bf16_2sm(tmem_d + 64u * j, adesc + 2u * k, bdesc + 2u * k, idesc64, 1u);
              else umma_bf16(tmem_d + 64u * j, adesc + 2u * k, bdesc + 2u * k, idesc64, 1u);
            }
            if (CG == 2) umma_commit_2sm(empty_bar(stage));
            else umma_commit(empty_bar(stage));
            if (++stage == C::STAGES) { stage = 0; phase ^= 1u; }
          }
        }
        // accumulator complete -> epilogue (of both CTAs when CG = 2)
        if (CG == 2) umma_commit_2sm(tfull_bar(acc));
        else umma_commit(tfull_bar(acc));
        if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1u; }
      }
    }
  } else {
    // ===== epilogue warps: TMEM -> registers -> swizzled smem staging -> TMA store =====
    // Each warp owns NB_STAGE staging tiles (32 rows x 128 B) used round-robin over a running chunk counter g, so the
    // TMA store of chunk g (read latency ~2-3k cycles) has NB_STAGE - 1 chunks of math to finish before its tile is
    // reused; with two tiles the epilogue, not the MMA, set the pace (11-16k cycles per 128 x 256 tile vs 8k of MMA).
    // Eight warps: warps w and w+4 share a TMEM lane quadrant (rows) and split the tile's columns in two halves.
    const int quad = warp & 3;  // TMEM lanes [32*quad, 32*quad+32) are accessible to this warp
    const int ew = warp - 2;
    const int col_half = ew >> 2;
    const int c_begin = BLOCK_N >= 128 ? col_half * (BLOCK_N / 2) : 0;                       // this warp's columns of the tile
    const int c_end = BLOCK_N >= 128 ? c_begin + BLOCK_N / 2 : (col_half == 0 ? BLOCK_N : 0);
    const uint32_t stage_base = bar_base + C::BAR_BYTES + static_cast<uint32_t>(ew) * (NB_STAGE * 4096u);
    const bool out_bf16 = ep.out_dtype == FDM_BF16;
    const int CW = out_bf16 ? 64 : 32;  // output columns per 128-byte staging row
    int acc = 0;
    uint32_t acc_phase = 0;
    uint32_t g = 0;  // chunks processed by this warp so far
    if (ep.tma_r) {
      // ---- bf16 out + bf16 residual, both through TMA: the residual chunk is fetched into the staging tile two chunks
      //      ahead, each thread adds its own row in place, and the same tile is handed to the TMA store engine ----
      auto n_chunks_of = [&](int tile) {  // 64-column chunks of this warp's half that start inside the matrix
        const int n_blk = tile % n_tiles;
        const int avail = N - n_blk * BLOCK_N - c_begin;
        return avail <= 0 ? 0 : min((c_end - c_begin) / 64, (avail + 63) / 64);
      };
      auto issue_residual = [&](int tile, int c, uint32_t gg) {  // lane 0 only
        const int m_blk = tile / n_tiles, n_blk = tile - m_blk * n_tiles;
        const uint32_t b = gg % NB_STAGE;
        bulk_wait_read<0>();  // the store that last used this staging tile (issued one chunk ago) has released it
        mbar_expect_tx(res_bar(ew, b), 4096);
        tma_load_2d(stage_base + b * 4096u, &tmap_r, res_bar(ew, b), n_blk * BLOCK_N + c_begin + c * 64,
                    m_blk * (BLOCK_M * CG) + row_in_tile + quad * 32);
      };
      // look-ahead cursor: position of chunk g + 2
      int la_tile = tile_first, la_c = 0;
      uint32_t la_g = 0;
      auto skip_empty = [&]() {  // tiles whose columns of this half lie outside the matrix have no chunks
        while (la_tile < num_tiles && n_chunks_of(la_tile) == 0) la_tile += tile_step;
      };
      auto advance_la = [&]() {
        ++la_g;
        if (++la_c >= n_chunks_of(la_tile)) { la_c = 0; la_tile += tile_step; skip_empty(); }
      };
      if (lane == 0) {
        skip_empty();
        if (la_tile < num_tiles) { issue_residual(la_tile, la_c, la_g); advance_la(); }  // one chunk ahead
      }
      for (int tile = tile_first; tile < num_tiles; tile += tile_step) {
        const int m_blk = tile / n_tiles, n_blk = tile - m_blk * n_tiles;
        const int row_w = m_blk * (BLOCK_M * CG) + row_in_tile + quad * 32;
        const int n_base = n_blk * BLOCK_N + c_begin;
        const int n_chunks = n_chunks_of(tile);
        mbar_wait(tfull_bar(acc), acc_phase);
        tcgen05_fence_after();
        const uint32_t tacc = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * BLOCK_N + c_begin;
        const int64_t row = static_cast<int64_t>(row_w) + lane;
        const bool row_ok = row < M;
        float r_mean = 0.f, r_rstd = 1.f;
        if (FOLD && ep.res_ln && row_ok) {
          const float2 mr = __ldg(reinterpret_cast<const float2*>(ep.res_ln) + row);
          r_mean = mr.x;
          r_rstd = mr.y;
        }
#pragma unroll 1
        for (int c = 0; c < n_chunks; ++c, ++g) {
          const int col0 = n_base + c * 64;
          const uint32_t b = g % NB_STAGE;
          const uint32_t sbuf = stage_base + b * 4096u;
          float st_s = 0.f, st_q = 0.f;  // sum / sum of squares of this thread's 64 output values (stats_out)
          mbar_wait(res_bar(ew, b), (g / NB_STAGE) & 1u);
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            if (col0 + half * 32 >= N) break;  // warp-uniform
            uint32_t r[32];
            float v[32];
            tmem_ld_32x32b_x32(tacc + c * 64 + half * 32, r);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
            {  // bias + activation (no residual here)
              Epilogue e2 = ep;
              e2.residual = nullptr;
              epilogue_math<false>(v, e2, 0, false, col0 + half * 32, N);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint32_t addr = sbuf + stage_off(lane, half * 4 + j);
              uint32_t q0, q1, q2, q3;
              asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(q0), "=r"(q1), "=r"(q2), "=r"(q3) : "r"(addr));
              const uint32_t qq[4] = {q0, q1, q2, q3};
              uint32_t oo[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&qq[e]));
                if (FOLD && ep.res_ln) {  // the residual tile holds un-normalised rows: rebuild LN(u) element-wise
                  const int col = col0 + half * 32 + 8 * j + 2 * e;
                  const float2 gm = __ldg(reinterpret_cast<const float2*>(ep.res_gamma + col));
                  const float2 bt = __ldg(reinterpret_cast<const float2*>(ep.res_beta + col));
                  f.x = fmaf((f.x - r_mean) * r_rstd, gm.x, bt.x);
                  f.y = fmaf((f.y - r_mean) * r_rstd, gm.y, bt.y);
                }
                const float o0 = v[8 * j + 2 * e] + f.x, o1 = v[8 * j + 2 * e + 1] + f.y;
                if (FOLD && ep.stats_out) {  // (warp-uniform; the plain path must not pay for the statistics)
                  st_s += o0 + o1;
                  st_q = fmaf(o0, o0, fmaf(o1, o1, st_q));
                }
                oo[e] = pack2(o0, o1);
              }
              st_shared_v4(addr, oo[0], oo[1], oo[2], oo[3]);
            }
          }
          if (FOLD && ep.stats_out && row_ok)
            *reinterpret_cast<float2*>(ep.stats_out + (row * ep.stats_parts + (col0 >> 6)) * 2) = make_float2(st_s, st_q);
          fence_async_smem();
          __syncwarp();
          if (lane == 0) {
            // next chunk's residual into the other staging tile: its last store was issued a whole chunk ago
            if (la_tile < num_tiles) { issue_residual(la_tile, la_c, la_g); advance_la(); }
            tma_store_2d(&tmap_c, sbuf, col0, row_w);
            bulk_commit();
          }
        }
        tcgen05_fence_before();
        if (CG == 2) mbar_arrive_remote(tempty_bar(acc), 0);
        else mbar_arrive(tempty_bar(acc));
        if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1u; }
      }
    } else
    for (int tile = tile_first; tile < full_tiles; tile += tile_step) {
      const int m_blk = tile / n_tiles, n_blk = tile - m_blk * n_tiles;
      mbar_wait(tfull_bar(acc), acc_phase);
      tcgen05_fence_after();
      const int row_w = m_blk * (BLOCK_M * CG) + row_in_tile + quad * 32;  // first row of this warp's slab
      const int64_t row = static_cast<int64_t>(row_w) + lane;
      const bool row_ok = row < M;
      float a_mean = 0.f, a_rstd = 1.f;
      if (FOLD && ep.a_ln && row_ok) {
        const float2 mr = __ldg(reinterpret_cast<const float2*>(ep.a_ln) + row);
        a_mean = mr.x;
        a_rstd = mr.y;
      }
      const uint32_t tacc = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * BLOCK_N;
#pragma unroll 1
      for (int c0 = c_begin; c0 < c_end; c0 += CW, ++g) {
        const int col0 = n_blk * BLOCK_N + c0;
        if (col0 >= N) break;  // warp-uniform
        const uint32_t sbuf = stage_base + (g % NB_STAGE) * 4096u;
        if (lane == 0) bulk_wait_read<NB_STAGE - 1>();  // the store issued NB_STAGE chunks ago has finished reading this tile
        __syncwarp();
        {
          uint32_t r[32];
          float v[32];
          tmem_ld_32x32b_x32(tacc + c0, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
          epilogue_math<FOLD>(v, ep, row, row_ok, col0, N, a_mean, a_rstd);
          if (out_bf16) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
              st_shared_v4(sbuf + stage_off(lane, j), pack2(v[8 * j], v[8 * j + 1]), pack2(v[8 * j + 2], v[8 * j + 3]),
                           pack2(v[8 * j + 4], v[8 * j + 5]), pack2(v[8 * j + 6], v[8 * j + 7]));
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j)
              st_shared_v4(sbuf + stage_off(lane, j), __float_as_uint(v[4 * j]), __float_as_uint(v[4 * j + 1]),
                           __float_as_uint(v[4 * j + 2]), __float_as_uint(v[4 * j + 3]));
          }
        }
        if (out_bf16 && col0 + 32 < N) {  // second half of the 64-column bf16 chunk
          uint32_t r[32];
          float v[32];
          tmem_ld_32x32b_x32(tacc + c0 + 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
          epilogue_math<FOLD>(v, ep, row, row_ok, col0 + 32, N, a_mean, a_rstd);
#pragma unroll
          for (int j = 0; j < 4; ++j)
            st_shared_v4(sbuf + stage_off(lane, 4 + j), pack2(v[8 * j], v[8 * j + 1]), pack2(v[8 * j + 2], v[8 * j + 3]),
                         pack2(v[8 * j + 4], v[8 * j + 5]), pack2(v[8 * j + 6], v[8 * j + 7]));
        }
        if (ep.tma_c) {
          fence_async_smem();  // generic-proxy writes -> visible to the TMA engine
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&tmap_c, sbuf, col0, row_w);  // rows >= M and columns >= N are clipped by the tensor map
            bulk_commit();
          }
        } else {
          // C rows are not 16-byte aligned (N = 15069 / 70110 fp32: the vertex maps; TMA needs 16-byte aligned box starts, which
          // no tensor map over such a pitch can give): coalesced element stores, one row per iteration. fp32 fast path: the
          // warp stores a full 32-column row per instruction, eight rows' shared-memory reads in flight at a time.
          __syncwarp();
          const int esz = out_bf16 ? 2 : 4;
          if (!out_bf16 && col0 + 32 <= N) {
            const int nr = min(32, M - row_w);
            float* cp = reinterpret_cast<float*>(ep.C) + static_cast<int64_t>(row_w) * ep.ldc + col0 + lane;
            const uint32_t sw = static_cast<uint32_t>(lane >> 2), lo = static_cast<uint32_t>(lane & 3) * 4u;
#pragma unroll
            for (int r8 = 0; r8 < 32; r8 += 8) {
              float fv[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const int rr = r8 + i;
                asm("ld.shared.f32 %0, [%1];" : "=f"(fv[i]) : "r"(sbuf + static_cast<uint32_t>(rr) * 128u + ((sw ^ (rr & 7)) << 4) + lo));
              }
#pragma unroll
              for (int i = 0; i < 8; ++i)
                if (r8 + i < nr) cp[static_cast<int64_t>(r8 + i) * ep.ldc] = fv[i];
            }
          } else
          for (int rr = 0; rr < 32; ++rr) {
            const int64_t orow = static_cast<int64_t>(row_w) + rr;
            if (orow >= M) break;
            for (int cc = lane; cc < CW; cc += 32) {
              if (col0 + cc >= N) break;
              const int byte = cc * esz;
              const uint32_t src = sbuf + stage_off(rr, byte >> 4) + (byte & 15);
              if (out_bf16) {
                uint16_t hv;
                asm volatile("ld.shared.u16 %0, [%1];" : "=h"(hv) : "r"(src));
                reinterpret_cast<uint16_t*>(ep.C)[orow * ep.ldc + col0 + cc] = hv;
              } else {
                float fv;
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(fv) : "r"(src));
                reinterpret_cast<float*>(ep.C)[orow * ep.ldc + col0 + cc] = fv;
              }
            }
          }
          __syncwarp();
        }
      }
      tcgen05_fence_before();
      if (CG == 2) mbar_arrive_remote(tempty_bar(acc), 0);
      else mbar_arrive(tempty_bar(acc));
      if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1u; }
    }
    if constexpr (TAILK) {
      if (tail_tile >= 0 && c_end > c_begin) {
        // ===== K-slice of a tail tile: fp32 partial -> this slice's workspace slab; the last slice to arrive at this warp's
        //       region (32 rows x this warp's columns) sums the slabs in slice order, applies the epilogue and stores C =====
        const int m_blk = tail_tile / n_tiles, n_blk = tail_tile - m_blk * n_tiles;
        mbar_wait(tfull_bar(acc), acc_phase);
        tcgen05_fence_after();
        const uint32_t tacc = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * BLOCK_N;
        const int slab_row = row_in_tile + quad * 32;  // this warp's first row inside a (128 * CG)-row slab
#pragma unroll 1
        for (int c0 = c_begin; c0 < c_end; c0 += 32, ++g) {
          if (n_blk * BLOCK_N + c0 >= N) break;  // warp-uniform (nobody reads those columns)
          const uint32_t sbuf = stage_base + (g % NB_STAGE) * 4096u;
          if (lane == 0) bulk_wait_read<NB_STAGE - 1>();
          __syncwarp();
          uint32_t r[32];
          tmem_ld_32x32b_x32(tacc + c0, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 8; ++j) st_shared_v4(sbuf + stage_off(lane, j), r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
          fence_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&tk.ws, sbuf, c0, tail_slab * (BLOCK_M * CG) + slab_row);
            bulk_commit();
          }
        }
        tcgen05_fence_before();
        if (CG == 2) mbar_arrive_remote(tempty_bar(acc), 0);
        else mbar_arrive(tempty_bar(acc));
        const int t_idx = tail_tile - tk.num_full;
        int* counter = tk.counters + (t_idx * CG + static_cast<int>(cta_rank)) * EPI_WARPS + ew;
        int arrived = 0;
        if (lane == 0) {
          bulk_wait_all();   // this warp's slab stores have COMPLETED (written, not merely read from shared memory)
          fence_async_all();  // async-proxy writes -> ordered before the generic-proxy release below
          __threadfence();
          arrived = atomicAdd(counter, 1);
        }
        arrived = __shfl_sync(0xffffffffu, arrived, 0);
        if (arrived == tk.S - 1) {
          __threadfence();  // acquire: every slice's slab rows of this region are visible (read through L2 below)
          const int row_w = m_blk * (BLOCK_M * CG) + slab_row;
          const int64_t row = static_cast<int64_t>(row_w) + lane;
          const bool row_ok = row < M;
          const uint32_t buf_t = stage_base, buf_o = stage_base + 4096u;  // fp32 transpose scratch, output tile
          const float* slab0 = tk.slabs + (static_cast<int64_t>(t_idx) * tk.S * (BLOCK_M * CG) + slab_row) * BLOCK_N;
          const int64_t slab_stride = static_cast<int64_t>(BLOCK_M * CG) * BLOCK_N;
#pragma unroll 1
          for (int c0 = c_begin; c0 < c_end; c0 += CW) {
            const int col0 = n_blk * BLOCK_N + c0;
            if (col0 >= N) break;  // warp-uniform
#pragma unroll 1
            for (int half = 0; half < CW / 32; ++half) {
              if (col0 + half * 32 >= N) break;  // warp-uniform
              // coalesced sum of the S slabs: lane -> (row 4 i + lane / 8, columns 4 (lane % 8) .. + 3)
              float4 acc4[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) acc4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
              const float* src = slab0 + static_cast<int64_t>(lane >> 3) * BLOCK_N + c0 + half * 32 + 4 * (lane & 7);
#pragma unroll 1
              for (int sl = 0; sl < tk.S; ++sl) {
                float4 x[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) x[i] = __ldcg(reinterpret_cast<const float4*>(src + sl * slab_stride + static_cast<int64_t>(4 * i) * BLOCK_N));
#pragma unroll
                for (int i = 0; i < 8; ++i) { acc4[i].x += x[i].x; acc4[i].y += x[i].y; acc4[i].z += x[i].z; acc4[i].w += x[i].w; }
              }
              __syncwarp();  // every lane has finished reading its row of the scratch tile (previous half)
#pragma unroll
              for (int i = 0; i < 8; ++i)
                st_shared_v4(buf_t + stage_off(4 * i + (lane >> 3), lane & 7), __float_as_uint(acc4[i].x), __float_as_uint(acc4[i].y),
                             __float_as_uint(acc4[i].z), __float_as_uint(acc4[i].w));
              __syncwarp();
              float v[32];
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                uint32_t q0, q1, q2, q3;
                asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(q0), "=r"(q1), "=r"(q2), "=r"(q3) : "r"(buf_t + stage_off(lane, j)));
                v[4 * j] = __uint_as_float(q0); v[4 * j + 1] = __uint_as_float(q1);
                v[4 * j + 2] = __uint_as_float(q2); v[4 * j + 3] = __uint_as_float(q3);
              }
              epilogue_math<false>(v, ep, row, row_ok, col0 + half * 32, N);
              if (half == 0) {  // the previous chunk's TMA store has finished reading the output tile
                if (lane == 0) bulk_wait_read<0>();
                __syncwarp();
              }
              if (out_bf16) {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                  st_shared_v4(buf_o + stage_off(lane, half * 4 + j), pack2(v[8 * j], v[8 * j + 1]), pack2(v[8 * j + 2], v[8 * j + 3]),
                               pack2(v[8 * j + 4], v[8 * j + 5]), pack2(v[8 * j + 6], v[8 * j + 7]));
              } else {
#pragma unroll
                for (int j = 0; j < 8; ++j)
                  st_shared_v4(buf_o + stage_off(lane, j), __float_as_uint(v[4 * j]), __float_as_uint(v[4 * j + 1]),
                               __float_as_uint(v[4 * j + 2]), __float_as_uint(v[4 * j + 3]));
              }
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) {
              tma_store_2d(&tmap_c, buf_o, col0, row_w);  // rows >= M and columns >= N are clipped by the tensor map
              bulk_commit();
            }
          }
          if (lane == 0) atomicExch(counter, 0);  // all S slices have arrived: ready for the next launch
        }
      }
    }
    if (lane == 0) bulk_wait_all();  // all TMA stores of this warp have completed before the CTA retires
  }

  pdl_trigger();  // this CTA's tiles are done: the next kernel's CTAs may start their prologue on freed SMs
  tcgen05_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();  // no CTA of the pair exits while its partner may still signal it
  if (warp == 2) {
    tcgen05_fence_after();
    if (CG == 2) tmem_dealloc_2sm(tmem_base, C::TMEM_COLS);
    else tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

// ---- host side -------------------------------------------------------------------------------------
// 2-D bf16 tensor map: inner extent `cols` (contiguous), `rows` rows with `ld` elements between rows,
// box = 64 x box_rows, 128-byte swizzle, out-of-bounds reads return zero.
int make_tmap(CUtensorMap* out, const void* ptr, int64_t cols, int64_t rows, int64_t ld, int box_rows, int box_cols = BLOCK_K,
              bool f32 = false) {
  PFN_tmapEncodeTiled enc = tmap_encode_fn();
  FDM_CHECK_ARG(enc != nullptr, "cuTensorMapEncodeTiled is unavailable (no CUDA driver?)");
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * (f32 ? 4 : 2)};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(box_cols), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  FDM_CHECK_ARG(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d (cols=%lld rows=%lld ld=%lld)",
                static_cast<int>(r), (long long)cols, (long long)rows, (long long)ld);
  return 0;
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// IDENT_COPIES x (64 x 64 bf16 identity) (the B operand of the residual k-blocks), per device, created by fdm_device_info()
void* g_identity[64] = {nullptr};
int g_resmma = -1;  // -1: read FDM_B200_GEMM_RESMMA on first use; fdm_gemm_set_option overrides

// Tail split-K schedule (see TailK): returns S >= 2 when the last wave of `tiles` on `slots` is at most half full
struct TailPlan { int num_full, rem, S, kb_per_slice; };
inline bool plan_tail(int64_t tiles, int64_t slots, int num_k_blocks, TailPlan* p) {
  const int64_t full = tiles / slots * slots, rem = tiles - full;
  if (rem == 0 || rem * 2 > slots) return false;
  int S = static_cast<int>(slots / rem);
  if (S > 8) S = 8;
  if (S > num_k_blocks / 2) S = num_k_blocks / 2;
  if (S < 2) return false;
  const int kbs = (num_k_blocks + S - 1) / S;
  S = (num_k_blocks + kbs - 1) / kbs;  // no empty slice
  if (S < 2) return false;
  p->num_full = static_cast<int>(full); p->rem = static_cast<int>(rem); p->S = S; p->kb_per_slice = kbs;
  return true;
}
constexpr int64_t TAILK_COUNTER_BYTES = 4096;  // counters at the start of the workspace, slabs behind them

template <int BLOCK_N, int CG, bool FOLD, bool SPLIT, bool RESMMA, bool TAILK = false>
int launch_impl(const fdm_gemm_args& a, const Epilogue& ep_in, cudaStream_t stream) {
  using C = Cfg<BLOCK_N, CG>;
  static bool attr_set = false;
  if (!attr_set) {
    FDM_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BLOCK_N, CG, FOLD, SPLIT, RESMMA, TAILK>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    attr_set = true;
  }
  Epilogue ep = ep_in;
  ResMaps<RESMMA> rm;
  if constexpr (RESMMA) {  // the residual goes through the tensor core: the epilogue sees a GEMM without one
    if (!fdm_gemm_identity()) {  // a device fdm_device_info() was never called on (allocates: not during stream capture)
      if (int rc = fdm_gemm_init_device()) return rc;
    }
    const void* ident = fdm_gemm_identity();
    if (int rc = make_tmap(&rm.res, a.residual, a.N, a.M, a.ldr, BLOCK_M)) return rc;
    if (int rc = make_tmap(&rm.ident, ident, 64, 64 * IDENT_COPIES, 64, 64 / CG)) return rc;
    ep.residual = nullptr;
    ep.tma_r = 0;
    ep.vec_r = 0;
  }
  const int taps = a.taps > 1 ? a.taps : 1;
  // grouped mode: the A map spans every group's columns, the producer moves the window with the N tile
  const int64_t a_cols = (taps > 1 ? a.tap_k : a.K) * (a.a_group_cols > 0 ? ceil_div64(a.N, BLOCK_N) : 1);
  CUtensorMap tm_a, tm_b, tm_c;
  if (int rc = make_tmap(&tm_a, a.A, a_cols, a.a_rows, a.lda, BLOCK_M)) return rc;
  if (int rc = make_tmap(&tm_b, a.W, a.K, a.N, a.ldw, BLOCK_N / CG)) return rc;
  if (ep.tma_c) {
    const bool f32 = a.out_dtype == FDM_F32;
    if (int rc = make_tmap(&tm_c, a.C, a.N, a.M, a.ldc, 32, f32 ? 32 : 64, f32)) return rc;
  } else {
    tm_c = tm_a;  // unused placeholder (must still be a valid parameter)
  }
  CUtensorMap tm_r = tm_a;
  if (ep.tma_r) {
    if (int rc = make_tmap(&tm_r, a.residual, a.N, a.M, a.ldr, 32, 64, false)) return rc;
  }
  LoMaps<SPLIT> lo;
  if constexpr (SPLIT) {
    if (int rc = make_tmap(&lo.a, a.A_lo, a_cols, a.a_rows, a.lda, BLOCK_M)) return rc;
    if (int rc = make_tmap(&lo.b, a.W_lo, a.K, a.N, a.ldw, BLOCK_N / CG)) return rc;
  }
  const int m_tiles = static_cast<int>(ceil_div64(a.M, BLOCK_M * CG));
  const int n_tiles = static_cast<int>(ceil_div64(a.N, BLOCK_N));
  const int kb_per_seg = static_cast<int>(ceil_div64(a.K, BLOCK_K));
  const int num_k_blocks = kb_per_seg * (SPLIT ? 3 : 1);
  const int kb_per_tap = taps > 1 ? static_cast<int>(a.tap_k / BLOCK_K) : kb_per_seg;
  const int64_t tiles = static_cast<int64_t>(m_tiles) * n_tiles;
  static const int sm_limit = [] { const char* e = getenv("FDM_B200_GEMM_SMS"); return e ? atoi(e) : 0; }();  // experiments only
  const int64_t slots = (sm_limit > 0 ? sm_limit : fdm_sm_count()) / CG;  // CTAs (CG = 1) or CTA pairs (CG = 2) the device holds
  int grid = static_cast<int>((tiles < slots ? tiles : slots) * CG);
  TailK<TAILK> tk;
  if constexpr (TAILK) {
    TailPlan tp;
    const bool ok = plan_tail(tiles, slots, num_k_blocks, &tp);
    const int64_t need = TAILK_COUNTER_BYTES + static_cast<int64_t>(tp.rem) * tp.S * (BLOCK_M * CG) * BLOCK_N * 4;
    FDM_CHECK_ARG(ok && a.splitk_ws && a.splitk_ws_bytes >= need && tp.rem * CG * EPI_WARPS * 4 <= TAILK_COUNTER_BYTES,
                  "fdm_gemm_bf16: internal: tail split-K launched without a plan / workspace");
    tk.counters = reinterpret_cast<int*>(a.splitk_ws);
    tk.slabs = reinterpret_cast<const float*>(reinterpret_cast<const char*>(a.splitk_ws) + TAILK_COUNTER_BYTES);
    tk.num_full = tp.num_full; tk.rem = tp.rem; tk.S = tp.S; tk.kb_per_slice = tp.kb_per_slice;
    if (int rc = make_tmap(&tk.ws, tk.slabs, BLOCK_N, static_cast<int64_t>(tp.rem) * tp.S * (BLOCK_M * CG), BLOCK_N, 32, 32, true)) return rc;
    if (tp.num_full == 0) grid = tp.rem * tp.S * CG;  // fewer tiles than CTAs: only the slices run
  }
  FDM_CHECK_CUDA(fdm_launch_pdl(gemm_tc_kernel<BLOCK_N, CG, FOLD, SPLIT, RESMMA, TAILK>, dim3(grid), dim3(NUM_THREADS), C::SMEM_BYTES, stream, CG, tm_a, tm_b,
                                tm_c, tm_r, lo, rm, tk, ep, static_cast<int>(a.M), static_cast<int>(a.N), num_k_blocks,
                                kb_per_tap, taps > 1 ? static_cast<int>(a.tap_row_shift) : 0, m_tiles, n_tiles, kb_per_seg,
                                static_cast<int>(a.a_group_cols)));
  return 0;
}

template <int BLOCK_N, int CG = 1>
int launch(const fdm_gemm_args& a, const Epilogue& ep, cudaStream_t stream) {
  if (a.A_lo) {  // split-bf16 operands (never combined with LayerNorm folding)
    FDM_CHECK_ARG(!(ep.a_ln || ep.res_ln || ep.stats_out), "fdm_gemm_bf16: split-bf16 operands cannot be combined with LayerNorm folding");
    return launch_impl<BLOCK_N, CG, false, true, false>(a, ep, stream);
  }
  if (ep.a_ln || ep.res_ln || ep.stats_out) return launch_impl<BLOCK_N, CG, true, false, false>(a, ep, stream);
  // bf16 residual without an activation: optionally added by the tensor core (FDM_B200_GEMM_RESMMA=1; default: the
  // TMA-residual epilogue)
  if (g_resmma < 0) {
    const char* e = getenv("FDM_B200_GEMM_RESMMA");
    g_resmma = (e && e[0] == '1') ? 1 : 0;
  }
  if (g_resmma == 1 && a.residual && a.res_dtype == FDM_BF16 && a.act == FDM_ACT_NONE && aligned16(a.residual) && (a.ldr * 2) % 16 == 0)
    return launch_impl<BLOCK_N, CG, false, false, true>(a, ep, stream);
  if constexpr (BLOCK_N == 256 && CG == 2) {
    // tail split-K (see TailK): plain TMA-store epilogue, workspace given, and a last wave that is at most half full
    static const bool tailk_on = [] { const char* e = getenv("FDM_B200_GEMM_TAILK"); return !(e && e[0] == '0'); }();
    if (tailk_on && a.splitk_ws && ep.tma_c && !ep.tma_r && a.a_group_cols == 0) {
      static const int sm_limit = [] { const char* e = getenv("FDM_B200_GEMM_SMS"); return e ? atoi(e) : 0; }();
      const int64_t slots = (sm_limit > 0 ? sm_limit : fdm_sm_count()) / CG;
      const int64_t tiles = ceil_div64(a.M, BLOCK_M * CG) * ceil_div64(a.N, BLOCK_N);
      TailPlan tp;
      if (plan_tail(tiles, slots, static_cast<int>(ceil_div64(a.K, BLOCK_K)), &tp) &&
          a.splitk_ws_bytes >= TAILK_COUNTER_BYTES + static_cast<int64_t>(tp.rem) * tp.S * (BLOCK_M * CG) * BLOCK_N * 4 &&
          tp.rem * CG * EPI_WARPS * 4 <= TAILK_COUNTER_BYTES)
        return launch_impl<BLOCK_N, CG, false, false, false, true>(a, ep, stream);
    }
  }
  return launch_impl<BLOCK_N, CG, false, false, false>(a, ep, stream);
}

}  // namespace

const void* fdm_gemm_identity() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  return g_identity[dev];
}
int fdm_gemm_init_device() {  // not during stream capture: allocates
  int dev = 0;
  FDM_CHECK_CUDA(cudaGetDevice(&dev));
  FDM_CHECK_ARG(dev >= 0 && dev < 64, "fdm_gemm_init_device: device index %d out of range", dev);
  if (g_identity[dev]) return 0;
  static uint16_t host[64 * 64];
  for (int i = 0; i < 64; ++i)
    for (int j = 0; j < 64; ++j) host[i * 64 + j] = i == j ? 0x3F80 : 0;  // bf16(1.0)
  void* p = nullptr;
  FDM_CHECK_CUDA(cudaMalloc(&p, sizeof(host) * IDENT_COPIES));
  for (int c = 0; c < IDENT_COPIES; ++c)
    FDM_CHECK_CUDA(cudaMemcpy(static_cast<uint8_t*>(p) + sizeof(host) * c, host, sizeof(host), cudaMemcpyHostToDevice));
  g_identity[dev] = p;
  return 0;
}

extern "C" int fdm_gemm_set_option(const char* name, int32_t value) {
  FDM_CHECK_ARG(name != nullptr, "fdm_gemm_set_option: null name");
  if (strcmp(name, "resmma") == 0) {
    g_resmma = value ? 1 : 0;
    return 0;
  }
  FDM_CHECK_ARG(false, "fdm_gemm_set_option: unknown option '%s'", name);
  return 1;
}

extern "C" int fdm_gemm_bf16(const fdm_gemm_args* args, void* stream) {
  FDM_CHECK_ARG(args != nullptr, "fdm_gemm_bf16: null args");
  const fdm_gemm_args& a = *args;
  FDM_CHECK_ARG(a.A && a.W && a.C, "fdm_gemm_bf16: null operand");
  FDM_CHECK_ARG(a.M > 0 && a.N > 0 && a.K > 0, "fdm_gemm_bf16: empty problem M=%lld N=%lld K=%lld", (long long)a.M, (long long)a.N, (long long)a.K);
  FDM_CHECK_ARG(a.M < (1ll << 31) && a.N < (1ll << 31) && a.K < (1ll << 31), "fdm_gemm_bf16: dimension too large");
  FDM_CHECK_ARG(aligned16(a.A) && aligned16(a.W), "fdm_gemm_bf16: A and W must be 16-byte aligned");
  FDM_CHECK_ARG(a.lda % 8 == 0 && a.ldw % 8 == 0, "fdm_gemm_bf16: lda/ldw must be multiples of 8 elements (TMA 16-byte strides)");
  FDM_CHECK_ARG(a.a_rows >= a.M, "fdm_gemm_bf16: a_rows < M");
  if (a.taps > 1) {
    FDM_CHECK_ARG(a.tap_k > 0 && a.tap_k % BLOCK_K == 0 && a.K == a.taps * a.tap_k,
                  "fdm_gemm_bf16: implicit conv needs K == taps*tap_k and tap_k %% 64 == 0");
  }
  FDM_CHECK_ARG(a.out_dtype == FDM_F32 || a.out_dtype == FDM_BF16, "fdm_gemm_bf16: bad out_dtype");
  FDM_CHECK_ARG((a.A_lo == nullptr) == (a.W_lo == nullptr), "fdm_gemm_bf16: split-bf16 operands need both A_lo and W_lo");
  FDM_CHECK_ARG(!a.A_lo || (aligned16(a.A_lo) && aligned16(a.W_lo)), "fdm_gemm_bf16: A_lo and W_lo must be 16-byte aligned");
  FDM_CHECK_ARG(!a.splitk_ws || ((reinterpret_cast<uintptr_t>(a.splitk_ws) & 255u) == 0 && a.splitk_ws_bytes > 0),
                "fdm_gemm_bf16: splitk_ws must be 256-byte aligned with splitk_ws_bytes > 0");
  Epilogue ep;
  ep.bias = a.bias;
  ep.residual = a.residual;
  ep.C = a.C;
  ep.ldr = a.ldr;
  ep.ldc = a.ldc;
  ep.res_dtype = a.res_dtype;
  ep.out_dtype = a.out_dtype;
  ep.act = a.act;
  static const bool gelu_poly = [] { const char* e = getenv("FDM_B200_GELU_POLY"); return !(e && e[0] == '0'); }();
  if (gelu_poly && a.act == FDM_ACT_GELU_ERF && a.out_dtype == FDM_BF16 && !a.A_lo) ep.act = ACT_GELU_ERF_BF16;
  const int64_t csz = a.out_dtype == FDM_BF16 ? 2 : 4, rsz = a.res_dtype == FDM_BF16 ? 2 : 4;
  ep.tma_c = aligned16(a.C) && (a.ldc * csz) % 16 == 0;
  ep.M = static_cast<int32_t>(a.M);
  ep.vec_r = a.residual && aligned16(a.residual) && (a.ldr * rsz) % 16 == 0;
  ep.vec_bias = a.bias && aligned16(a.bias);
  // fp32 output whose row pitch is not a multiple of 16 bytes (the vertex maps): element stores, but still the big tiles
  ep.lin_c = !ep.tma_c && a.out_dtype == FDM_F32;
  ep.tma_r = ep.tma_c && a.residual && a.res_dtype == FDM_BF16 && a.out_dtype == FDM_BF16 && aligned16(a.residual) &&
             (a.ldr * 2) % 16 == 0;
  ep.a_ln = a.a_ln;
  ep.w_colsum = a.w_colsum;
  ep.res_ln = a.res_ln;
  ep.res_gamma = a.res_gamma;
  ep.res_beta = a.res_beta;
  ep.stats_out = a.stats_out;
  ep.stats_parts = static_cast<int32_t>(a.N / 64);
  FDM_CHECK_ARG(!a.a_ln || (a.w_colsum && !ep.tma_r), "fdm_gemm_bf16: a_ln needs w_colsum and a GEMM without the bf16 residual path");
  FDM_CHECK_ARG(!a.res_ln || (ep.tma_r && a.res_gamma && a.res_beta && a.N % 64 == 0 && aligned16(a.res_gamma) && aligned16(a.res_beta)),
                "fdm_gemm_bf16: res_ln needs a bf16 residual / bf16 output on the TMA path, res_gamma, res_beta and N %% 64 == 0");
  FDM_CHECK_ARG(!a.stats_out || (ep.tma_r && a.N % 64 == 0),
                "fdm_gemm_bf16: stats_out needs the bf16 residual TMA path and N %% 64 == 0");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);

  // Tile width: the widest tile that still yields at least one tile per SM; narrow problems fall to 64.
  const int64_t m_tiles = ceil_div64(a.M, BLOCK_M);
  const int sms = fdm_sm_count();
  if (a.a_group_cols != 0) {  // grouped convolution: one 64-column tile per group
    FDM_CHECK_ARG(a.a_group_cols > 0 && a.a_group_cols % BLOCK_K == 0 && a.N % 64 == 0 &&
                      a.a_group_cols == (a.taps > 1 ? a.tap_k : a.K) && a.lda >= a.a_group_cols * (a.N / 64),
                  "fdm_gemm_bf16: grouped mode needs N %% 64 == 0, a_group_cols == tap_k (or K) %% 64 == 0 and lda covering every group");
    FDM_CHECK_ARG(!a.a_ln && !a.res_ln && !a.stats_out, "fdm_gemm_bf16: grouped mode has no LayerNorm folding");
    return launch<64>(a, ep, s);
  }
  {  // experiments: FDM_B200_GEMM_FORCE = 2562 | 1282 | 1281 | 641 forces <BLOCK_N, CG> wherever it is legal
    static const int force = [] { const char* e = getenv("FDM_B200_GEMM_FORCE"); return e ? atoi(e) : 0; }();
    if (force == 2562 && ep.tma_c && a.N >= 256 && a.M >= 512) return launch<256, 2>(a, ep, s);
    if (force == 1282 && ep.tma_c && a.N >= 128 && a.M >= 512) return launch<128, 2>(a, ep, s);
    if (force == 1281 && a.N >= 128) return launch<128>(a, ep, s);
    if (force == 641) return launch<64>(a, ep, s);
  }
  // CTA-pair kernel (256 x 256 tiles) whenever there is at least one tile per SM pair and C can go through TMA
  static const bool two_cta = [] { const char* e = getenv("FDM_B200_GEMM_2CTA"); return !(e && e[0] == '0'); }();
  if (two_cta && (ep.tma_c || ep.lin_c) && a.N >= 256 && a.M >= 512 && ceil_div64(a.M, 256) * ceil_div64(a.N, 256) >= sms / 2) {
    // 256 x 128 tiles (FDM_B200_GEMM_BN128=1) for schedules whose last 256 x 256 wave is mostly empty (N = 1024 at M = 25344:
    // 396 tiles on 74 CTA pairs = 5.35 waves, 89 % of six; 792 half-width tiles fill 97 % of eleven)
    static const int bn128 = [] { const char* e = getenv("FDM_B200_GEMM_BN128"); return e ? atoi(e) : 0; }();
    const int64_t pairs = sms / 2, t256 = ceil_div64(a.M, 256) * ceil_div64(a.N, 256), t128 = ceil_div64(a.M, 256) * ceil_div64(a.N, 128);
    const double e256 = static_cast<double>(t256) / (pairs * ceil_div64(t256, pairs));
    const double e128 = static_cast<double>(t128) / (pairs * ceil_div64(t128, pairs));
    (void)e256; (void)e128;
    // measured (profiles/README.md, r02): the half-width tiles lose more in the main loop (each A tile feeds half the MMA
    // work per shared-memory fetch) than the fuller last wave returns - out-projection 55.4 -> 71.4 us - so this is opt-in
    if (bn128 == 1) return launch<128, 2>(a, ep, s);
    return launch<256, 2>(a, ep, s);
  }
  if (a.N >= 128 && m_tiles * ceil_div64(a.N, 128) >= sms) return launch<128>(a, ep, s);
  if (a.N > 64 && a.N % 64 != 0 && a.N >= 128) return launch<128>(a, ep, s);
  return launch<64>(a, ep, s);
}
