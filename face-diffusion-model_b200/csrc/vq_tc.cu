// fdm_vq_quantize, tensor-core path (D = 64): EVQ-VAE nearest-code search as a distance GEMM on tcgen05 with an exact
// recheck, so that the indices stay BIT-EXACT against the defined fp32 expression of the oracle (oracle/vq_ref.c,
// restating models/lib/quantizer.py:35-64 and models/vq_vae_emotion.py:221-252):
//     d_j = fl(fl(zz + ee_j) - fl(2 * dot_j)),  zz / ee_j / dot_j sequential fmaf chains over k = 0..63 from +0.0f,
//     index = argmin_j d_j, lowest j on ties.
// The FFMA kernel (vq.cu) evaluates all 256 chains per row: 32.8 kFLOP per 520 bytes of traffic, compute-bound at 5 % of
// the HBM roofline. Here the tensor cores only FILTER:
//   * z and the codebook are split into bf16 pairs (x = hi + lo + r, |r| <= 2^-18 |x|) and
//         acc_j = z_hi.e_hi + z_lo.e_hi + z_hi.e_lo            (bf16 products are exact in fp32, fp32 accumulate in TMEM)
//     is one K = 192 tcgen05.mma chain per 128-row tile against the smem-resident codebook (N = 256 codes).
//   * v_j = fl(ee_j - 2 acc_j) ranks the codes. With DOT = 2^-12 |z| |e|_max >= |acc_j - dot_j| (expected error is ~2^-16:
//     3 * 2^-18 dropped terms, 2^-18 for the reference's own chain rounding, ~2^-18 accumulation; tests/test_kernels_gpu.py
//     measures it on the device and asserts < 2^-14) and RND = 2^-20 (zz + ee_max) >= every rounding of d_j and v_j,
//         |(zz + v_j) - d_j| <= delta := 2 DOT + RND                                   for every code j,
//     so a code whose v_j exceeds the smallest v by more than 2 delta cannot be the arg-min of d. A row with exactly one
//     code inside that window is decided; otherwise (ties, near-ties, NaNs: a fraction of a percent of the rows) the
//     thread evaluates the exact fmaf chains of the candidate codes only, in ascending j with strict '<' like the oracle.
// Pipeline of one persistent CTA (448 threads, 1 CTA / SM, contiguous range of 128-row tiles):
//   warp 5      cp.async.bulk (1-D TMA) of fp32 z half-tiles (64 rows = 16 KB) into a 5-deep ring
//   warps 0-3   fp32 -> (hi, lo) bf16 in the 128B-swizzled K-major A-operand layout, |z|^2 estimate per row
//   warp 4      12 x tcgen05.mma (128 x 256 x 16) per tile into one of two 256-column TMEM accumulators
//   warps 6-13  two epilogue groups (one per accumulator): TMEM -> registers, running (min, second min, index), recheck,
//               int64 index, gather of the winning code row into z_q (B, D, L) and / or (B, L, D)
// Per-clip codebook slices (MEAD emotions) are handled as segments: the CTA drains, reloads the slice, continues.
#include "common.cuh"
#include <cuda.h>
#include <stdlib.h>

namespace {

constexpr int D = 64;
constexpr int TILE_ROWS = 128;
constexpr int HALF_ROWS = 64;
constexpr int STG_STAGES = 5;
constexpr int STG_BYTES = HALF_ROWS * D * 4;  // 16 KB of fp32 rows
constexpr int A_STAGES = 2;
constexpr int A_OP_BYTES = TILE_ROWS * 128;   // 128 rows x 64 bf16
constexpr int A_STAGE_BYTES = 2 * A_OP_BYTES; // hi, lo
constexpr int B_OP_BYTES = 256 * 128;         // 256 codes x 64 bf16
constexpr int ZZ_SLOTS = 4;
constexpr int NUM_THREADS = 448;
constexpr int CONV_THREADS = 128;

constexpr int OFF_B = 0;                                   // e_hi, e_lo
constexpr int OFF_A = OFF_B + 2 * B_OP_BYTES;              // 2 stages x (z_hi, z_lo)
constexpr int OFF_STG = OFF_A + A_STAGES * A_STAGE_BYTES;  // fp32 staging ring
constexpr int OFF_EE = OFF_STG + STG_STAGES * STG_BYTES;   // ee[256]
constexpr int OFF_ZZ = OFF_EE + 256 * 4;                   // zz estimate [ZZ_SLOTS][128]
constexpr int OFF_BAR = OFF_ZZ + ZZ_SLOTS * TILE_ROWS * 4;
constexpr int NUM_BARS = 2 * STG_STAGES + 2 * A_STAGES + 2 * 2;
constexpr int OFF_MISC = OFF_BAR + NUM_BARS * 8;           // tmem slot, ee_max bits
constexpr int SMEM_BYTES = OFF_MISC + 16 + 1024;           // + manual 1024-byte alignment

constexpr float C_DOT = 1.0f / 2048.0f;      // 2 * 2^-12
constexpr float C_RND = 1.0f / 1048576.0f;   // 2^-20

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0, spins = 0;
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (done) break;
    if (++spins > (1u << 26)) {  // watchdog: a protocol bug must trap, never hang the GPU
      printf("fdm vq_tc: mbarrier wait timed out (block %d thread %d bar %u)\n", blockIdx.x, threadIdx.x, bar);
      __trap();
    }
  }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
// K-major SWIZZLE_128B operand descriptor: 8-row x 128-byte atoms, SBO = 1024 B, version 1 (sm_100), layout 2
__device__ __forceinline__ uint64_t make_kmajor_sw128_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
__device__ __forceinline__ uint32_t make_idesc_bf16_f32(int m, int n) {  // bf16 x bf16 -> f32, both K-major
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(m >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// x[0..7] -> bf16 hi (round to nearest) and bf16 lo = rn(x - hi); element k sits at byte 2k of the 16-byte chunk
__device__ __forceinline__ void split8(const float4& a, const float4& b, uint4& hi, uint4& lo) {
  const float x[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
  uint32_t h[4], l[4];
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    const __nv_bfloat162 hh = __floats2bfloat162_rn(x[2 * p], x[2 * p + 1]);
    const float2 hf = __bfloat1622float2(hh);
    const __nv_bfloat162 ll = __floats2bfloat162_rn(x[2 * p] - hf.x, x[2 * p + 1] - hf.y);
    h[p] = *reinterpret_cast<const uint32_t*>(&hh);
    l[p] = *reinterpret_cast<const uint32_t*>(&ll);
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}

struct Params {
  const float* z;
  const float* codebook;
  const int64_t* code_offset;
  int64_t B, L;
  int n_codes;
  int64_t* indices;
  float* zq_bdl;
  float* zq_rows;
  unsigned long long* recheck_rows;  // optional counter of rows that took the exact path
  float* dbg_acc;                    // optional [rows, n_codes] dump of the tensor-core dot products (tests)
  int64_t tiles_per_clip, num_tiles;
};

// exact distance of the oracle for one (row, code): sequential fmaf chains, every operation rounded to fp32
__device__ __noinline__ float exact_dot(const float* __restrict__ zr, const float* __restrict__ e) {
  float dot = 0.f;
  const float4* z4 = reinterpret_cast<const float4*>(zr);
  const float4* e4 = reinterpret_cast<const float4*>(e);
#pragma unroll 4
  for (int k = 0; k < D / 4; ++k) {
    const float4 a = z4[k], b = __ldg(e4 + k);
    dot = fmaf(a.x, b.x, dot);
    dot = fmaf(a.y, b.y, dot);
    dot = fmaf(a.z, b.z, dot);
    dot = fmaf(a.w, b.w, dot);
  }
  return dot;
}

#define VQ_UPD(t, val, j)                         \
  {                                               \
    const float _v = (val);                       \
    m2[t] = fminf(m2[t], fmaxf(m1[t], _v));       \
    if (_v < m1[t]) id[t] = (j);                  \
    m1[t] = fminf(m1[t], _v);                     \
  }

__global__ void __launch_bounds__(NUM_THREADS, 1) vq_tc_kernel(const Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw_addr);
  float* ee = reinterpret_cast<float*>(smem + OFF_EE);
  float* zzs = reinterpret_cast<float*>(smem + OFF_ZZ);
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + OFF_MISC);
  uint32_t* ee_max_bits = reinterpret_cast<uint32_t*>(smem + OFF_MISC + 4);
  auto stg_full = [&](int s) { return base + OFF_BAR + 8u * s; };
  auto stg_empty = [&](int s) { return base + OFF_BAR + 8u * (STG_STAGES + s); };
  auto a_full = [&](int s) { return base + OFF_BAR + 8u * (2 * STG_STAGES + s); };
  auto a_empty = [&](int s) { return base + OFF_BAR + 8u * (2 * STG_STAGES + A_STAGES + s); };
  auto t_full = [&](int s) { return base + OFF_BAR + 8u * (2 * STG_STAGES + 2 * A_STAGES + s); };
  auto t_empty = [&](int s) { return base + OFF_BAR + 8u * (2 * STG_STAGES + 2 * A_STAGES + 2 + s); };

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_codes = p.n_codes;
  const int nchunks = n_codes >> 5;
  const int64_t L = p.L, tpc = p.tiles_per_clip;

  if (tid == 0) {
    for (int s = 0; s < STG_STAGES; ++s) { mbar_init(stg_full(s), 1); mbar_init(stg_empty(s), CONV_THREADS / 32); }
    for (int s = 0; s < A_STAGES; ++s) { mbar_init(a_full(s), CONV_THREADS / 32); mbar_init(a_empty(s), 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(t_full(s), 1); mbar_init(t_empty(s), 4); }
    fence_barrier_init();
  } else if (warp == 4) {
    tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_slot)), 512);
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int64_t t_begin = static_cast<int64_t>(blockIdx.x) * p.num_tiles / gridDim.x;
  const int64_t t_end = static_cast<int64_t>(blockIdx.x + 1) * p.num_tiles / gridDim.x;
  uint32_t it = 0;  // tiles processed so far (every role counts the same sequence)
  uint32_t hc = 0;  // staged half-tiles so far
  int64_t tile = t_begin;

  while (tile < t_end) {
    // ---- segment = maximal run of tiles that share one codebook slice ------------------------------------------
    const int64_t b0 = tile / tpc;
    const int64_t off = p.code_offset ? p.code_offset[b0] : 0;
    int64_t seg_end = t_end;
    if (p.code_offset) {
      for (int64_t bb = b0 + 1; bb * tpc < t_end; ++bb)
        if (p.code_offset[bb] != off) { seg_end = bb * tpc; break; }
    }
    const float* cbg = p.codebook + off * D;

    __syncthreads();  // every role is done with the previous slice (epilogue waits imply its MMAs have retired)
    if (tid == 0) *ee_max_bits = 0u;
    for (int i = tid; i < n_codes * 8; i += NUM_THREADS) {
      const int c = i >> 3, c8 = i & 7;
      const float4* src = reinterpret_cast<const float4*>(cbg + c * D + c8 * 8);
      uint4 hi, lo;
      split8(__ldg(src), __ldg(src + 1), hi, lo);
      const uint32_t o = static_cast<uint32_t>(c * 128 + ((c8 ^ (c & 7)) << 4));
      st_shared_v4(base + OFF_B + o, hi);
      st_shared_v4(base + OFF_B + B_OP_BYTES + o, lo);
    }
    __syncthreads();
    for (int c = tid; c < n_codes; c += NUM_THREADS) {
      float s = 0.f;
      for (int k = 0; k < D; ++k) { const float e = __ldg(cbg + c * D + k); s = fmaf(e, e, s); }
      ee[c] = s;
      if (s == s) atomicMax(ee_max_bits, __float_as_uint(fabsf(s)));  // non-negative floats order like their bit patterns
    }
    fence_async_smem();
    __syncthreads();
    const float ee_max = __uint_as_float(*ee_max_bits);

    if (warp < 4) {
      // ===== converters: staged fp32 rows -> bf16 (hi, lo) A operand + |z|^2 estimate =====
      for (int64_t tl = tile; tl < seg_end; ++tl, ++it) {
        const int64_t b = tl / tpc;
        const int64_t l0 = (tl - b * tpc) * TILE_ROWS;
        const int nrows = static_cast<int>(min(static_cast<int64_t>(TILE_ROWS), L - l0));
        const int as = it & 1;
        mbar_wait(a_empty(as), ((it >> 1) & 1u) ^ 1u);
        const uint32_t a_hi = base + OFF_A + as * A_STAGE_BYTES, a_lo = a_hi + A_OP_BYTES;
        float* zz = zzs + (it & (ZZ_SLOTS - 1)) * TILE_ROWS;
        for (int half = 0; half < 2; ++half) {
          if (nrows - half * HALF_ROWS <= 0) break;
          const int s = hc % STG_STAGES;
          mbar_wait(stg_full(s), (hc / STG_STAGES) & 1u);
          const uint8_t* stg = smem + OFF_STG + s * STG_BYTES;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int u = tid + CONV_THREADS * q;
            const int row = u >> 3, c8 = u & 7;
            const float4* src = reinterpret_cast<const float4*>(stg + row * (D * 4) + c8 * 32);
            const float4 x0 = src[0], x1 = src[1];
            uint4 hi, lo;
            split8(x0, x1, hi, lo);
            const int arow = half * HALF_ROWS + row;
            const uint32_t o = static_cast<uint32_t>(arow * 128 + ((c8 ^ (row & 7)) << 4));
            st_shared_v4(a_hi + o, hi);
            st_shared_v4(a_lo + o, lo);
            float sq = x0.x * x0.x;
            sq = fmaf(x0.y, x0.y, sq); sq = fmaf(x0.z, x0.z, sq); sq = fmaf(x0.w, x0.w, sq);
            sq = fmaf(x1.x, x1.x, sq); sq = fmaf(x1.y, x1.y, sq); sq = fmaf(x1.z, x1.z, sq); sq = fmaf(x1.w, x1.w, sq);
            sq += __shfl_xor_sync(0xffffffffu, sq, 1);
            sq += __shfl_xor_sync(0xffffffffu, sq, 2);
            sq += __shfl_xor_sync(0xffffffffu, sq, 4);
            if (c8 == 0) zz[arow] = sq;
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(stg_empty(s));  // this warp has read its part of the staging slot
          ++hc;
        }
        fence_async_smem();  // generic-proxy stores -> visible to tcgen05.mma (async proxy)
        __syncwarp();
        if (lane == 0) mbar_arrive(a_full(as));
      }
    } else if (warp == 4) {
      // ===== MMA issuer =====
      const uint32_t idesc = make_idesc_bf16_f32(TILE_ROWS, n_codes);
      const uint64_t b_hi = make_kmajor_sw128_desc(base + OFF_B), b_lo = make_kmajor_sw128_desc(base + OFF_B + B_OP_BYTES);
      for (int64_t tl = tile; tl < seg_end; ++tl, ++it) {
        if (lane == 0) {
          const int as = it & 1;
          const uint32_t ph = (it >> 1) & 1u;
          mbar_wait(t_empty(as), ph ^ 1u);  // the epilogue group has drained this accumulator
          mbar_wait(a_full(as), ph);
          tcgen05_fence_after();
          const uint32_t tmem_d = tmem_base + as * 256;
          const uint64_t a_hi = make_kmajor_sw128_desc(base + OFF_A + as * A_STAGE_BYTES);
          const uint64_t a_lo = make_kmajor_sw128_desc(base + OFF_A + as * A_STAGE_BYTES + A_OP_BYTES);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(tmem_d, a_hi + 2u * k, b_hi + 2u * k, idesc, k != 0 ? 1u : 0u);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(tmem_d, a_lo + 2u * k, b_hi + 2u * k, idesc, 1u);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(tmem_d, a_hi + 2u * k, b_lo + 2u * k, idesc, 1u);
          umma_commit(a_empty(as));
          umma_commit(t_full(as));
        }
        __syncwarp();
      }
    } else if (warp == 5) {
      // ===== loader: 1-D bulk copies of contiguous fp32 rows =====
      for (int64_t tl = tile; tl < seg_end; ++tl, ++it) {
        const int64_t b = tl / tpc;
        const int64_t l0 = (tl - b * tpc) * TILE_ROWS;
        const int nrows = static_cast<int>(min(static_cast<int64_t>(TILE_ROWS), L - l0));
        for (int half = 0; half < 2; ++half) {
          const int hr = min(HALF_ROWS, nrows - half * HALF_ROWS);
          if (hr <= 0) break;
          if (lane == 0) {
            const int s = hc % STG_STAGES;
            mbar_wait(stg_empty(s), ((hc / STG_STAGES) & 1u) ^ 1u);
            mbar_expect_tx(stg_full(s), static_cast<uint32_t>(hr) * D * 4);
            bulk_load_1d(base + OFF_STG + s * STG_BYTES, p.z + (b * L + l0 + half * HALF_ROWS) * D,
                         static_cast<uint32_t>(hr) * D * 4, stg_full(s));
          }
          ++hc;
        }
        __syncwarp();
      }
    } else {
      // ===== epilogue: group g owns accumulator g =====
      const int grp = (warp - 6) >> 2;
      const int quad = warp & 3;
      const int row = quad * 32 + lane;
      const float se_max = sqrtf(ee_max) * 1.001f;
      const float4* ee4 = reinterpret_cast<const float4*>(ee);
      for (int64_t tl = tile; tl < seg_end; ++tl, ++it) {
        if ((it & 1) != static_cast<uint32_t>(grp)) continue;
        const int64_t b = tl / tpc;
        const int64_t l0 = (tl - b * tpc) * TILE_ROWS;
        const int nrows = static_cast<int>(min(static_cast<int64_t>(TILE_ROWS), L - l0));
        const bool row_ok = row < nrows;
        const int64_t grow = b * L + l0 + row;
        mbar_wait(t_full(grp), (it >> 1) & 1u);
        tcgen05_fence_after();
        const uint32_t tacc = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + grp * 256;
        const float zzr = zzs[(it & (ZZ_SLOTS - 1)) * TILE_ROWS + row];

        float m1[4], m2[4];
        int id[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) { m1[t] = INFINITY; m2[t] = INFINITY; id[t] = 0; }
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          if (c < nchunks) {
            uint32_t r[32];
            tmem_ld_32x32b_x32(tacc + c * 32, r);
            tmem_ld_wait();
            if (p.dbg_acc && row_ok) {
#pragma unroll
              for (int j = 0; j < 32; ++j) p.dbg_acc[grow * n_codes + c * 32 + j] = __uint_as_float(r[j]);
            }
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const float4 e = ee4[c * 8 + q];
              VQ_UPD(0, fmaf(-2.f, __uint_as_float(r[4 * q + 0]), e.x), c * 32 + 4 * q + 0);
              VQ_UPD(1, fmaf(-2.f, __uint_as_float(r[4 * q + 1]), e.y), c * 32 + 4 * q + 1);
              VQ_UPD(2, fmaf(-2.f, __uint_as_float(r[4 * q + 2]), e.z), c * 32 + 4 * q + 2);
              VQ_UPD(3, fmaf(-2.f, __uint_as_float(r[4 * q + 3]), e.w), c * 32 + 4 * q + 3);
            }
          }
        }
        float M1 = m1[0], M2 = m2[0];
        int idx = id[0];
#pragma unroll
        for (int t = 1; t < 4; ++t) {
          M2 = fminf(fminf(M2, m2[t]), fmaxf(M1, m1[t]));
          if (m1[t] < M1) idx = id[t];
          M1 = fminf(M1, m1[t]);
        }
        const float delta = C_DOT * sqrtf(zzr * 1.001f) * se_max + C_RND * (zzr + ee_max) + 1e-37f;
        const float window = 2.f * delta;
        const bool flagged = row_ok && !(M2 - M1 > window);
        if (__any_sync(0xffffffffu, flagged)) {
          // second pass: candidate mask of the flagged rows, then exact chains for the candidates only
          const float thr = M1 + window;
          uint32_t mask[8];
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            mask[c] = 0u;
            if (c < nchunks) {
              uint32_t r[32];
              tmem_ld_32x32b_x32(tacc + c * 32, r);
              tmem_ld_wait();
              uint32_t m = 0u;
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const float v = fmaf(-2.f, __uint_as_float(r[j]), ee[c * 32 + j]);
                m |= (v <= thr ? 1u : 0u) << j;
              }
              mask[c] = flagged ? m : 0u;
            }
          }
          if (flagged) {
            const float* zr = p.z + grow * D;
            float zz = 0.f;
            for (int k = 0; k < D; ++k) zz = fmaf(zr[k], zr[k], zz);
            float best = INFINITY;
            int bi = 0;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              uint32_t m = mask[c];
              while (m) {
                const int j = c * 32 + __ffs(m) - 1;
                m &= m - 1u;
                const float dot = exact_dot(zr, cbg + j * D);
                const float dist = __fsub_rn(__fadd_rn(zz, ee[j]), __fmul_rn(2.f, dot));
                if (dist < best) { best = dist; bi = j; }
              }
            }
            idx = bi;
            if (p.recheck_rows) atomicAdd(p.recheck_rows, 1ull);
          }
        }
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(t_empty(grp));  // accumulator free: the MMAs of tile it + 2 may start

        // ---- outputs: index, gathered code rows ----
        if (row_ok && p.indices) p.indices[grow] = idx;
        if (p.zq_bdl && row_ok) {  // (B, D, L): for each k the warp writes 32 consecutive floats
          const float4* src = reinterpret_cast<const float4*>(cbg + idx * D);
          float* dst = p.zq_bdl + (b * D) * L + l0 + row;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            float4 v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = __ldg(src + h * 8 + i);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int k = h * 32 + i * 4;
              dst[(k + 0) * L] = v[i].x;
              dst[(k + 1) * L] = v[i].y;
              dst[(k + 2) * L] = v[i].z;
              dst[(k + 3) * L] = v[i].w;
            }
          }
        }
        if (p.zq_rows) {  // (B, L, D): the warp copies one 256-byte code row per iteration
          const int wrows = min(32, nrows - quad * 32);
          for (int rr = 0; rr < wrows; ++rr) {
            const int ci = __shfl_sync(0xffffffffu, idx, rr);
            const float2 v = __ldg(reinterpret_cast<const float2*>(cbg + ci * D) + lane);
            reinterpret_cast<float2*>(p.zq_rows + (b * L + l0 + quad * 32 + rr) * D)[lane] = v;
          }
        }
      }
    }
    // `it` / `hc` are private per thread; each role advances the ones it uses over the same tile sequence
    tile = seg_end;
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem_base, 512);
}

}  // namespace

int fdm_vq_tc_launch(const float* z, const float* codebook, const int64_t* code_offset, int64_t B, int64_t L, int n_codes,
                     int64_t* indices, float* zq_bdl, float* zq_rows, unsigned long long* recheck_rows, float* dbg_acc,
                     cudaStream_t stream) {
  static bool attr = false;
  if (!attr) {
    FDM_CHECK_CUDA(cudaFuncSetAttribute(vq_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    attr = true;
  }
  Params p;
  p.z = z; p.codebook = codebook; p.code_offset = code_offset; p.B = B; p.L = L; p.n_codes = n_codes;
  p.indices = indices; p.zq_bdl = zq_bdl; p.zq_rows = zq_rows; p.recheck_rows = recheck_rows; p.dbg_acc = dbg_acc;
  p.tiles_per_clip = ceil_div64(L, TILE_ROWS);
  p.num_tiles = B * p.tiles_per_clip;
  const int64_t sms = fdm_sm_count();
  const int grid = static_cast<int>(p.num_tiles < sms ? p.num_tiles : sms);
  vq_tc_kernel<<<grid, NUM_THREADS, SMEM_BYTES, stream>>>(p);
  FDM_CHECK_LAUNCH();
  return 0;
}
