// Tensor-core self-attention for bf16 activations (head dim 64 / 128): flash-style online softmax with
// mma.sync.m16n8k16 (bf16 in, fp32 accumulate), K/V streamed through a double-buffered, XOR-swizzled shared
// memory ring with cp.async, Q fragments and the O accumulator held in registers.
//
// The FaceFormer-style bias (-slope_h * floor((t-j)/period), -inf above the diagonal; reference
// models/fdm_vocaset.py:94-115) depends only on (head, t-j): each CTA builds a (T+64)-entry table of it in shared
// memory once and every score costs one FFMA against a table entry (fetched two at a time) instead of an
// integer divide + compare + select; the causal mask is the table's -inf region, and causal key blocks beyond the
// diagonal are skipped. Softmax runs in the log2 domain on ex2.approx. CTA = 4 warps x 16 query rows; the Q tile
// is read into registers once and its buffer is recycled as the second K stage, so a CTA needs 66 KB and three
// fit on an SM (head dim 64: 41 KB and 128 registers, four per SM: 239 -> 234 us at 64 x 16 x 498).
// grid = (ceil(T/64), heads, sequences).
// (T <= 600 here and one head's K/V is at most 150 KB: the GEMMs around this kernel use tcgen05; attention is ~5 %
//  of the FLOPs and stays on the warp-level MMA path.)
#include "common.cuh"

namespace {

constexpr int QB = 64, KB = 64, THREADS = 128;
constexpr int TAB_MAX = 1024 + 64 + 8;  // T <= 1024

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
  const int sz = valid ? 16 : 0;  // src-size 0 -> 16 bytes of zeros
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float fast_exp2(float x) {  // ex2.approx.ftz: exp2(-inf) = +0, no range fix-ups
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// tile [rows][DH] bf16, 16-byte chunks XOR-swizzled by (row & 7)
template <int DH>
__device__ __forceinline__ uint32_t tile_off(int row, int chunk) {
  return static_cast<uint32_t>(row * (DH * 2) + ((chunk ^ (row & 7)) << 4));
}

// Cooperative 64-row tile load. Thread -> (row r_base + ROWS_PER_IT * it, chunk c): the swizzled chunk and the
// per-iteration strides are loop-invariant, so a tile costs one compare + one cp.async per 16 bytes.
template <int DH, int NTHR = THREADS>
struct TileLoader {
  static constexpr int CH = DH / 8;                 // 16-byte chunks per row
  static constexpr int ROWS_PER_IT = NTHR / CH;     // 8 (DH = 128) or 16 (DH = 64); 8 for DH = 256 with 256 threads
  static constexpr int ITERS = 64 / ROWS_PER_IT;
  int r_base;
  uint32_t soff;   // smem byte offset of (r_base, c) with the swizzle applied
  int goff;        // c * 8 elements
  __device__ __forceinline__ TileLoader() {
    r_base = threadIdx.x / CH;
    const int c = threadIdx.x - r_base * CH;
    soff = tile_off<DH>(r_base, c);
    goff = c * 8;
  }
  __device__ __forceinline__ void load(uint32_t smem_base, const __nv_bfloat16* g, int64_t ld, int row0, int n_rows_valid) const {
    const __nv_bfloat16* p = g + static_cast<int64_t>(row0 + r_base) * ld + goff;
    const int64_t step = static_cast<int64_t>(ROWS_PER_IT) * ld;
    uint32_t d = smem_base + soff;
    int r = row0 + r_base;
#pragma unroll
    for (int it = 0; it < ITERS; ++it) {
      const bool ok = r < n_rows_valid;
      cp_async16(d, ok ? p : g, ok);
      p += step;
      d += ROWS_PER_IT * DH * 2;  // ROWS_PER_IT is a multiple of 8: the swizzle phase (row & 7) is unchanged
      r += ROWS_PER_IT;
    }
  }
};

template <int DH, bool CAUSAL>
__global__ void __launch_bounds__(THREADS, DH == 64 ? 4 : 3) attn_mma_kernel(const fdm_attn_args a) {
  constexpr int KS = DH / 16;   // k-steps of Q.K^T
  constexpr int NT = DH / 8;    // n-tiles of P.V
  constexpr int TILE = 64 * DH * 2;
  extern __shared__ __align__(128) uint8_t smem[];
  // slots: [K0][V0][K1][V1]; the Q tile first lands in slot K1 and is consumed into registers before K1 is filled
  const uint32_t s0 = smem_u32(smem);
  float* tab = reinterpret_cast<float*>(smem + 4 * TILE);  // two copies: tab[0..TAB) and a copy shifted by one entry

  const int T = static_cast<int>(a.T);
  const int q0 = blockIdx.x * QB, h = blockIdx.y;
  const int64_t row0 = static_cast<int64_t>(blockIdx.z) * a.t_stride;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, tq = lane & 3;
  constexpr float LOG2E = 1.4426950408889634f;
  const float scale2 = a.scale * LOG2E;

  const __nv_bfloat16* Qg = reinterpret_cast<const __nv_bfloat16*>(a.Q) + row0 * a.ldq + static_cast<int64_t>(h) * DH;
  const __nv_bfloat16* Kg = reinterpret_cast<const __nv_bfloat16*>(a.K) + row0 * a.ldk + static_cast<int64_t>(h) * DH;
  const __nv_bfloat16* Vg = reinterpret_cast<const __nv_bfloat16*>(a.V) + row0 * a.ldv + static_cast<int64_t>(h) * DH;

  const int k_end = CAUSAL ? min(T, q0 + QB) : T;
  const int n_kb = (k_end + KB - 1) / KB;
  const int tabn = T + 64;  // entry k <-> delta = (T - 1) - k, delta in [-64, T-1]

  const TileLoader<DH> tl;
  pdl_trigger();
  pdl_wait();
  tl.load(s0 + 2 * TILE, Qg, a.ldq, q0, T);
  tl.load(s0, Kg, a.ldk, 0, T);
  tl.load(s0 + TILE, Vg, a.ldv, 0, T);
  cp_async_commit();
  if (CAUSAL) {
    const float slope2 = a.slopes[h] * LOG2E;
    const int period = a.period;
    for (int k = threadIdx.x; k < tabn; k += THREADS) {
      const int delta = (T - 1) - k;
      const float v = delta < 0 ? -INFINITY : -slope2 * static_cast<float>(delta / period);
      tab[k] = v;
      tab[TAB_MAX + k + 1] = v;
    }
  }
  cp_async_wait<0>();
  __syncthreads();

  // Q fragments for this warp's 16 rows
  uint32_t qf[KS][4];
#pragma unroll
  for (int ks = 0; ks < KS; ++ks) {
    const int r = warp * 16 + (lane & 15);
    const int c = ks * 2 + (lane >> 4);
    ldsm_x4(s0 + 2 * TILE + tile_off<DH>(r, c), qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3]);
  }
  __syncthreads();  // everyone has its Q fragments: slot K1 may be overwritten
  if (n_kb > 1) {
    tl.load(s0 + 2 * TILE, Kg, a.ldk, KB, T);
    tl.load(s0 + 3 * TILE, Vg, a.ldv, KB, T);
  }
  cp_async_commit();

  float o[NT][4];
#pragma unroll
  for (int i = 0; i < NT; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
  const int t_row0 = q0 + warp * 16 + g;  // second row of this thread: +8
  // bias-table cursors: entry for (row r, key j) is k = (T-1) - (t_r - j); LDS.64 needs an even index, so rows whose
  // index is odd read the shifted copy
  const float* tabp[2];
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int kb0 = (T - 1) - (t_row0 + 8 * r) + 2 * tq;  // + j0 + 8n (+e) later; may be negative only for garbage rows
    tabp[r] = (kb0 & 1) ? tab + TAB_MAX + kb0 + 1 : tab + kb0;
  }
  const int warp_last_row = q0 + warp * 16 + 15;

  for (int kb = 0; kb < n_kb; ++kb) {
    const uint32_t kbuf = s0 + (kb & 1) * 2 * TILE, vbuf = kbuf + TILE;
    const int j0 = kb * KB;
    if (!CAUSAL || j0 <= warp_last_row) {  // warp-uniform: this warp has at least one visible key in the block
      // ---- S = Q K^T (16 x 64 per warp) ----
      float s[8][4];
#pragma unroll
      for (int n = 0; n < 8; ++n) s[n][0] = s[n][1] = s[n][2] = s[n][3] = 0.f;
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
        for (int np = 0; np < 4; ++np) {  // pairs of 8-key n-tiles
          uint32_t b0, b1, b2, b3;
          const int key = np * 16 + ((lane >> 4) << 3) + (lane & 7);
          const int c = ks * 2 + ((lane >> 3) & 1);
          ldsm_x4(kbuf + tile_off<DH>(key, c), b0, b1, b2, b3);
          mma_bf16(s[2 * np], qf[ks], b0, b1);
          mma_bf16(s[2 * np + 1], qf[ks], b2, b3);
        }
      }
      // ---- scale + bias/mask (log2 domain) ----
      if (CAUSAL) {
#pragma unroll
        for (int n = 0; n < 8; ++n) {
          const float2 b0 = *reinterpret_cast<const float2*>(tabp[0] + j0 + 8 * n);
          const float2 b1 = *reinterpret_cast<const float2*>(tabp[1] + j0 + 8 * n);
          s[n][0] = fmaf(s[n][0], scale2, b0.x);
          s[n][1] = fmaf(s[n][1], scale2, b0.y);
          s[n][2] = fmaf(s[n][2], scale2, b1.x);
          s[n][3] = fmaf(s[n][3], scale2, b1.y);
        }
      } else {
        const bool edge = j0 + KB > T;  // block-uniform: keys beyond the sequence end must be masked
#pragma unroll
        for (int n = 0; n < 8; ++n)
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float x = s[n][e] * scale2;
            s[n][e] = (edge && j0 + n * 8 + tq * 2 + (e & 1) >= T) ? -INFINITY : x;
          }
      }
      // ---- online softmax ----
      float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
      for (int n = 0; n < 8; ++n) {
        mx0 = fmaxf(mx0, fmaxf(s[n][0], s[n][1]));
        mx1 = fmaxf(mx1, fmaxf(s[n][2], s[n][3]));
      }
      mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
      mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
      mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
      mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
      const float mn0 = fmaxf(m_run[0], mx0), mn1 = fmaxf(m_run[1], mx1);
      // a row with no visible key so far keeps m = -inf; use 0 as the reference so that exp2(-inf - 0) = 0
      const float ms0 = mn0 == -INFINITY ? 0.f : mn0, ms1 = mn1 == -INFINITY ? 0.f : mn1;
      const float corr0 = fast_exp2(m_run[0] - ms0), corr1 = fast_exp2(m_run[1] - ms1);
      m_run[0] = mn0;
      m_run[1] = mn1;
      float rs0 = 0.f, rs1 = 0.f;
      uint32_t pf[4][4];  // P as A fragments: 4 k-steps of 16 keys
#pragma unroll
      for (int n = 0; n < 8; ++n) {
        const float p0 = fast_exp2(s[n][0] - ms0), p1 = fast_exp2(s[n][1] - ms0);
        const float p2 = fast_exp2(s[n][2] - ms1), p3 = fast_exp2(s[n][3] - ms1);
        rs0 += p0 + p1;
        rs1 += p2 + p3;
        pf[n >> 1][(n & 1) * 2 + 0] = pack_bf16(p0, p1);
        pf[n >> 1][(n & 1) * 2 + 1] = pack_bf16(p2, p3);
      }
      rs0 += __shfl_xor_sync(0xffffffffu, rs0, 1);
      rs0 += __shfl_xor_sync(0xffffffffu, rs0, 2);
      rs1 += __shfl_xor_sync(0xffffffffu, rs1, 1);
      rs1 += __shfl_xor_sync(0xffffffffu, rs1, 2);
      l_run[0] = l_run[0] * corr0 + rs0;
      l_run[1] = l_run[1] * corr1 + rs1;
#pragma unroll
      for (int i = 0; i < NT; ++i) {
        o[i][0] *= corr0; o[i][1] *= corr0;
        o[i][2] *= corr1; o[i][3] *= corr1;
      }
      // ---- O += P V ----
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
        for (int np = 0; np < NT / 2; ++np) {
          uint32_t b0, b1, b2, b3;
          const int key = ks * 16 + (((lane >> 3) & 1) << 3) + (lane & 7);
          const int c = np * 2 + (lane >> 4);
          ldsm_x4_t(vbuf + tile_off<DH>(key, c), b0, b1, b2, b3);
          mma_bf16(o[2 * np], pf[ks], b0, b1);
          mma_bf16(o[2 * np + 1], pf[ks], b2, b3);
        }
      }
    }
    // ---- pipeline: refill this stage with block kb+2, make block kb+1 visible ----
    __syncthreads();
    if (kb + 2 < n_kb) {
      tl.load(kbuf, Kg, a.ldk, (kb + 2) * KB, T);
      tl.load(vbuf, Vg, a.ldv, (kb + 2) * KB, T);
    }
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
  }

  // ---- normalise, stage through shared memory (slot K0 is free), coalesced 16-byte stores ----
  const float inv0 = 1.f / l_run[0], inv1 = 1.f / l_run[1];
#pragma unroll
  for (int i = 0; i < NT; ++i) {
    const int r0 = warp * 16 + g, r1 = r0 + 8;
    const uint32_t w0 = s0 + tile_off<DH>(r0, i) + tq * 4;
    const uint32_t w1 = s0 + tile_off<DH>(r1, i) + tq * 4;
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(w0), "r"(pack_bf16(o[i][0] * inv0, o[i][1] * inv0)) : "memory");
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(w1), "r"(pack_bf16(o[i][2] * inv1, o[i][3] * inv1)) : "memory");
  }
  __syncthreads();
  __nv_bfloat16* Og = reinterpret_cast<__nv_bfloat16*>(a.O) + row0 * a.ldo + static_cast<int64_t>(h) * DH;
  constexpr int CH = DH / 8;
  for (int i = threadIdx.x; i < QB * CH; i += THREADS) {
    const int r = i / CH, c = i - r * CH;
    if (q0 + r < T) {
      uint4 v;
      asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(s0 + tile_off<DH>(r, c)));
      *reinterpret_cast<uint4*>(Og + static_cast<int64_t>(q0 + r) * a.ldo + c * 8) = v;
    }
  }
}

// ---- head dim 256 (BIWI FDM: 4 heads x 256) ------------------------------------------------------------------
// Same algorithm with 8 warps per CTA: warps w and w+4 own the same 16 query rows, both compute the full 16 x 64
// score tile (Q.K^T over 256 dims, Q fragments re-read from smem per key block instead of pinned in registers) and
// each accumulates one 128-column half of O, so the per-thread accumulator stays at 64 registers.
template <bool CAUSAL>
__global__ void __launch_bounds__(256, 1) attn_mma256_kernel(const fdm_attn_args a) {
  constexpr int DH = 256, KS = DH / 16, NTV = 16;  // NTV: n-tiles of this warp's 128-column half of V
  constexpr int TILE = 64 * DH * 2;                // 32 KB
  extern __shared__ __align__(128) uint8_t smem[];
  // slots: [Q][K0][V0][K1][V1]
  const uint32_t sQ = smem_u32(smem), s0 = sQ + TILE;
  float* tab = reinterpret_cast<float*>(smem + 5 * TILE);

  const int T = static_cast<int>(a.T);
  const int q0 = blockIdx.x * QB, h = blockIdx.y;
  const int64_t row0 = static_cast<int64_t>(blockIdx.z) * a.t_stride;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wq = warp & 3, vhalf = warp >> 2;  // query-row group, V column half
  const int g = lane >> 2, tq = lane & 3;
  constexpr float LOG2E = 1.4426950408889634f;
  const float scale2 = a.scale * LOG2E;

  const __nv_bfloat16* Qg = reinterpret_cast<const __nv_bfloat16*>(a.Q) + row0 * a.ldq + static_cast<int64_t>(h) * DH;
  const __nv_bfloat16* Kg = reinterpret_cast<const __nv_bfloat16*>(a.K) + row0 * a.ldk + static_cast<int64_t>(h) * DH;
  const __nv_bfloat16* Vg = reinterpret_cast<const __nv_bfloat16*>(a.V) + row0 * a.ldv + static_cast<int64_t>(h) * DH;

  const int k_end = CAUSAL ? min(T, q0 + QB) : T;
  const int n_kb = (k_end + KB - 1) / KB;
  const int tabn = T + 64;

  const TileLoader<DH, 256> tl;
  pdl_trigger();
  pdl_wait();
  tl.load(sQ, Qg, a.ldq, q0, T);
  tl.load(s0, Kg, a.ldk, 0, T);
  tl.load(s0 + TILE, Vg, a.ldv, 0, T);
  cp_async_commit();
  if (n_kb > 1) {
    tl.load(s0 + 2 * TILE, Kg, a.ldk, KB, T);
    tl.load(s0 + 3 * TILE, Vg, a.ldv, KB, T);
  }
  cp_async_commit();
  if (CAUSAL) {
    const float slope2 = a.slopes[h] * LOG2E;
    const int period = a.period;
    for (int k = threadIdx.x; k < tabn; k += 256) {
      const int delta = (T - 1) - k;
      const float v = delta < 0 ? -INFINITY : -slope2 * static_cast<float>(delta / period);
      tab[k] = v;
      tab[TAB_MAX + k + 1] = v;
    }
  }
  cp_async_wait<1>();
  __syncthreads();

  float o[NTV][4];
#pragma unroll
  for (int i = 0; i < NTV; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
  const int t_row0 = q0 + wq * 16 + g;
  const float* tabp[2];
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int kb0 = (T - 1) - (t_row0 + 8 * r) + 2 * tq;
    tabp[r] = (kb0 & 1) ? tab + TAB_MAX + kb0 + 1 : tab + kb0;
  }
  const int warp_last_row = q0 + wq * 16 + 15;

  for (int kb = 0; kb < n_kb; ++kb) {
    const uint32_t kbuf = s0 + (kb & 1) * 2 * TILE, vbuf = kbuf + TILE;
    const int j0 = kb * KB;
    if (!CAUSAL || j0 <= warp_last_row) {
      float s[8][4];
#pragma unroll
      for (int n = 0; n < 8; ++n) s[n][0] = s[n][1] = s[n][2] = s[n][3] = 0.f;
#pragma unroll 4
      for (int ks = 0; ks < KS; ++ks) {
        uint32_t qf[4];
        ldsm_x4(sQ + tile_off<DH>(wq * 16 + (lane & 15), ks * 2 + (lane >> 4)), qf[0], qf[1], qf[2], qf[3]);
#pragma unroll
        for (int np = 0; np < 4; ++np) {
          uint32_t b0, b1, b2, b3;
          const int key = np * 16 + ((lane >> 4) << 3) + (lane & 7);
          const int c = ks * 2 + ((lane >> 3) & 1);
          ldsm_x4(kbuf + tile_off<DH>(key, c), b0, b1, b2, b3);
          mma_bf16(s[2 * np], qf, b0, b1);
          mma_bf16(s[2 * np + 1], qf, b2, b3);
        }
      }
      if (CAUSAL) {
#pragma unroll
        for (int n = 0; n < 8; ++n) {
          const float2 b0 = *reinterpret_cast<const float2*>(tabp[0] + j0 + 8 * n);
          const float2 b1 = *reinterpret_cast<const float2*>(tabp[1] + j0 + 8 * n);
          s[n][0] = fmaf(s[n][0], scale2, b0.x);
          s[n][1] = fmaf(s[n][1], scale2, b0.y);
          s[n][2] = fmaf(s[n][2], scale2, b1.x);
          s[n][3] = fmaf(s[n][3], scale2, b1.y);
        }
      } else {
        const bool edge = j0 + KB > T;
#pragma unroll
        for (int n = 0; n < 8; ++n)
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float x = s[n][e] * scale2;
            s[n][e] = (edge && j0 + n * 8 + tq * 2 + (e & 1) >= T) ? -INFINITY : x;
          }
      }
      float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
      for (int n = 0; n < 8; ++n) {
        mx0 = fmaxf(mx0, fmaxf(s[n][0], s[n][1]));
        mx1 = fmaxf(mx1, fmaxf(s[n][2], s[n][3]));
      }
      mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
      mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
      mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
      mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
      const float mn0 = fmaxf(m_run[0], mx0), mn1 = fmaxf(m_run[1], mx1);
      const float ms0 = mn0 == -INFINITY ? 0.f : mn0, ms1 = mn1 == -INFINITY ? 0.f : mn1;
      const float corr0 = fast_exp2(m_run[0] - ms0), corr1 = fast_exp2(m_run[1] - ms1);
      m_run[0] = mn0;
      m_run[1] = mn1;
      float rs0 = 0.f, rs1 = 0.f;
      uint32_t pf[4][4];
#pragma unroll
      for (int n = 0; n < 8; ++n) {
        const float p0 = fast_exp2(s[n][0] - ms0), p1 = fast_exp2(s[n][1] - ms0);
        const float p2 = fast_exp2(s[n][2] - ms1), p3 = fast_exp2(s[n][3] - ms1);
        rs0 += p0 + p1;
        rs1 += p2 + p3;
        pf[n >> 1][(n & 1) * 2 + 0] = pack_bf16(p0, p1);
        pf[n >> 1][(n & 1) * 2 + 1] = pack_bf16(p2, p3);
      }
      rs0 += __shfl_xor_sync(0xffffffffu, rs0, 1);
      rs0 += __shfl_xor_sync(0xffffffffu, rs0, 2);
      rs1 += __shfl_xor_sync(0xffffffffu, rs1, 1);
      rs1 += __shfl_xor_sync(0xffffffffu, rs1, 2);
      l_run[0] = l_run[0] * corr0 + rs0;
      l_run[1] = l_run[1] * corr1 + rs1;
#pragma unroll
      for (int i = 0; i < NTV; ++i) {
        o[i][0] *= corr0; o[i][1] *= corr0;
        o[i][2] *= corr1; o[i][3] *= corr1;
      }
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
        for (int np = 0; np < NTV / 2; ++np) {
          uint32_t b0, b1, b2, b3;
          const int key = ks * 16 + (((lane >> 3) & 1) << 3) + (lane & 7);
          const int c = vhalf * 16 + np * 2 + (lane >> 4);
          ldsm_x4_t(vbuf + tile_off<DH>(key, c), b0, b1, b2, b3);
          mma_bf16(o[2 * np], pf[ks], b0, b1);
          mma_bf16(o[2 * np + 1], pf[ks], b2, b3);
        }
      }
    }
    __syncthreads();
    if (kb + 2 < n_kb) {
      tl.load(kbuf, Kg, a.ldk, (kb + 2) * KB, T);
      tl.load(vbuf, Vg, a.ldv, (kb + 2) * KB, T);
    }
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
  }

  // normalise and stage through the Q tile (no longer needed), then coalesced 16-byte stores
  const float inv0 = 1.f / l_run[0], inv1 = 1.f / l_run[1];
#pragma unroll
  for (int i = 0; i < NTV; ++i) {
    const int r0 = wq * 16 + g, r1 = r0 + 8;
    const uint32_t w0 = sQ + tile_off<DH>(r0, vhalf * 16 + i) + tq * 4;
    const uint32_t w1 = sQ + tile_off<DH>(r1, vhalf * 16 + i) + tq * 4;
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(w0), "r"(pack_bf16(o[i][0] * inv0, o[i][1] * inv0)) : "memory");
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(w1), "r"(pack_bf16(o[i][2] * inv1, o[i][3] * inv1)) : "memory");
  }
  __syncthreads();
  __nv_bfloat16* Og = reinterpret_cast<__nv_bfloat16*>(a.O) + row0 * a.ldo + static_cast<int64_t>(h) * DH;
  constexpr int CH = DH / 8;
  for (int i = threadIdx.x; i < QB * CH; i += 256) {
    const int r = i / CH, c = i - r * CH;
    if (q0 + r < T) {
      uint4 v;
      asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(sQ + tile_off<DH>(r, c)));
      *reinterpret_cast<uint4*>(Og + static_cast<int64_t>(q0 + r) * a.ldo + c * 8) = v;
    }
  }
}

template <bool CAUSAL>
int launch256(const fdm_attn_args& a, cudaStream_t stream) {
  const size_t smem = 5 * 64 * 256 * 2 + 2 * TAB_MAX * sizeof(float);
  static bool attr = false;
  if (!attr) {
    FDM_CHECK_CUDA(cudaFuncSetAttribute(attn_mma256_kernel<CAUSAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    attr = true;
  }
  dim3 grid(static_cast<unsigned>(ceil_div64(a.T, QB)), static_cast<unsigned>(a.H), static_cast<unsigned>(a.B));
  FDM_CHECK_CUDA(fdm_launch_pdl(attn_mma256_kernel<CAUSAL>, grid, dim3(256), smem, stream, 1, a));
  return 0;
}

template <int DH, bool CAUSAL>
int launch(const fdm_attn_args& a, cudaStream_t stream) {
  const size_t smem = 4 * 64 * DH * 2 + 2 * TAB_MAX * sizeof(float);
  static bool attr = false;
  if (!attr) {
    FDM_CHECK_CUDA(cudaFuncSetAttribute(attn_mma_kernel<DH, CAUSAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    attr = true;
  }
  dim3 grid(static_cast<unsigned>(ceil_div64(a.T, QB)), static_cast<unsigned>(a.H), static_cast<unsigned>(a.B));
  FDM_CHECK_CUDA(fdm_launch_pdl(attn_mma_kernel<DH, CAUSAL>, grid, dim3(THREADS), smem, stream, 1, a));
  return 0;
}

}  // namespace

int fdm_attention_mma_try(const fdm_attn_args& a, cudaStream_t stream, bool* handled) {
  *handled = false;
  if (a.dtype != FDM_BF16 || (a.dh != 64 && a.dh != 128 && a.dh != 256)) return 0;
  // 16-byte cp.async / vector stores need aligned rows; the bias table covers T <= 1024
  const uintptr_t al = reinterpret_cast<uintptr_t>(a.Q) | reinterpret_cast<uintptr_t>(a.K) | reinterpret_cast<uintptr_t>(a.V) |
                       reinterpret_cast<uintptr_t>(a.O);
  if ((al & 15u) != 0 || a.ldq % 8 || a.ldk % 8 || a.ldv % 8 || a.ldo % 8 || a.T > 1024) return 0;
  *handled = true;
  if (a.dh == 256) return a.bias_mode == 1 ? launch256<true>(a, stream) : launch256<false>(a, stream);
  if (a.bias_mode == 1) return a.dh == 64 ? launch<64, true>(a, stream) : launch<128, true>(a, stream);
  return a.dh == 64 ? launch<64, false>(a, stream) : launch<128, false>(a, stream);
}
