// tcgen05/TMEM self-attention for the FDM denoiser (bf16, head dim 128, causal + periodic ALiBi, T <= 208).
//
// Persistent CTAs, one (sequence, head) item at a time. The whole head lives on chip: K and V (T x 128 each) are TMA-loaded once, the two
// 128-row query tiles are processed back to back:
//     S_i = Q_i K^T          tcgen05.mma, both operands K-major from 128B-swizzled smem, fp32 scores in TMEM
//     P_i = softmax(S_i)     one THREAD per query row (TMEM lane): no shuffles, no online rescaling; bias/mask from a
//                            per-head table in smem; P written as bf16 into the K-major swizzled A-operand layout
//     O_i = P_i V            tcgen05.mma with V as an MN-major (N = head dim contiguous) B operand, fp32 in TMEM
//     epilogue               O_i / rowsum -> bf16 -> swizzled staging -> TMA store (rows >= T clipped by a 3-D map)
// Warp roles: warp 0 TMA, warp 1 MMA issue, warps 2-9 softmax + epilogue (two threads per query row). Every mbarrier
// completes exactly once per item, so the parity is the item counter's low bit. S_1's MMA overlaps softmax_0, O_0's MMA
// overlaps the max pass of softmax_1, and the next item's Q_0/K/V are prefetched as soon as their buffers are dead. TMEM: S_1 [0,208) | S_0 [256,384) | O [384,512).
// The mma.sync kernel (attention_mma.cu) stays the fallback for other shapes; it needs ~7 ALU/LDSM instructions per
// HMMA and is power-capped at ~220 TFLOP/s, this kernel moves the matrix work to the tcgen05 pipe.
#include "common.cuh"
#include <cuda.h>
#include <stdlib.h>

namespace {

constexpr int TK_MAX = 208;                 // keys padded to a multiple of 16
constexpr int DH = 128;
constexpr int NUM_THREADS = 320;            // warp 0 TMA, warp 1 MMA, warps 2-9 softmax/epilogue
constexpr int H_MAX = 8;
constexpr int SQ_BYTES = 128 * DH * 2;      // 32 KB: one query tile (2 k-blocks of 128 x 64)
constexpr int SKV_BYTES = TK_MAX * DH * 2;  // 52 KB
constexpr int SP_BYTES = 4 * 128 * 128;     // 64 KB: P as 4 k-blocks of 128 rows x 64 keys
constexpr int TAB_FLOATS = 512;             // bias table per head: delta in [-127, 384]
constexpr int SMEM_BYTES = SQ_BYTES + 2 * SKV_BYTES + SP_BYTES + (8 * TAB_FLOATS + 1024) * 4 + 256 + 1024;

enum { B_Q0K = 0, B_V, B_Q1, B_S0, B_S1, B_P0, B_O0, B_P1, B_OE0, B_O1, B_OE1, B_QKFREE, NUM_BARS };

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
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {  // every barrier completes once per item
  uint32_t done = 0, spins = 0;
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) break;
    if (++spins > (1u << 24)) {
      printf("fdm attention_tc: mbarrier %u timed out (block %d,%d thread %d)\n", bar, blockIdx.x, blockIdx.y, threadIdx.x);
      __trap();
    }
  }
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, uint32_t src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(src), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// smem matrix descriptor, SWIZZLE_128B. K-major operands: LBO ignored, SBO = 1024 B (8 rows x 128 B).
// MN-major operand (V): LBO = byte distance between 64-element MN blocks, SBO = 1024 B (8 k-rows x 128 B).
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= static_cast<uint64_t>(1) << 46;  // descriptor version (sm_100)
  d |= static_cast<uint64_t>(2) << 61;  // SWIZZLE_128B
  return d;
}
__device__ __forceinline__ uint32_t make_idesc(int m, int n, int b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(b_mn_major) << 16) | (static_cast<uint32_t>(n >> 3) << 17) |
         (static_cast<uint32_t>(m >> 4) << 24);
}
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

__device__ __forceinline__ void named_bar_sync(int id, int threads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory"); }

__global__ void __launch_bounds__(NUM_THREADS, 1)
attn_tc_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_kv,
               const __grid_constant__ CUtensorMap tm_o, const int T, const int Tk, const int d_model, const int H,
               const int n_items, const float scale2, const float* __restrict__ slopes, const int period) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const uint32_t sQ = base, sK = sQ + SQ_BYTES, sV = sK + SKV_BYTES, sP = sV + SKV_BYTES;
  float* tab = reinterpret_cast<float*>(smem + SQ_BYTES + 2 * SKV_BYTES + SP_BYTES);  // [H_MAX][TAB_FLOATS]
  float* xch = tab + H_MAX * TAB_FLOATS;                                              // 4 x [2][128] partial max / sum exchange
  const uint32_t bars = sP + SP_BYTES + (H_MAX * TAB_FLOATS + 1024) * 4;
  auto bar = [&](int i) { return bars + 8u * i; };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + SQ_BYTES + 2 * SKV_BYTES + SP_BYTES + (H_MAX * TAB_FLOATS + 1024) * 4 + 8 * NUM_BARS);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool two = T > 128;
  const int n0 = min(Tk, 128);          // keys visible to query tile 0 (causal)
  const uint32_t kv_blk = static_cast<uint32_t>(Tk) * 128u;  // bytes of one 64-column block of K or V

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm_q)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm_kv)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm_o)) : "memory");
  } else if (warp == 1 && lane == 0) {
    for (int i = 0; i < NUM_BARS; ++i) {
      const int cnt = (i == B_P0 || i == B_P1 || i == B_OE0 || i == B_OE1) ? 256 : (i == B_QKFREE ? 8 : 1);
      mbar_init(bar(i), cnt);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  } else if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(const_cast<uint32_t*>(tmem_slot))), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (warp >= 2) {
    // bias tables (one per head): tab[h][i] <-> delta = i - 127 = t - j; -inf above the diagonal
    // (models/fdm_vocaset.py:94-115), log2 domain
    for (int i = threadIdx.x - 64; i < H * TAB_FLOATS; i += 256) {
      const int hh = i / TAB_FLOATS, delta = (i - hh * TAB_FLOATS) - 127;
      tab[i] = delta < 0 ? -INFINITY : -(slopes[hh] * 1.4426950408889634f) * static_cast<float>(delta / period);
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tS1 = tmem, tS0 = tmem + 256, tO = tmem + 384;  // S_1 [0,208) | S_0 [256,384) | O [384,512)
  pdl_wait();

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer: the next item's Q_0/K and V are fetched while the current item is still in its softmax =====
      int it = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
        const uint32_t ph = it & 1, prev = ph ^ 1u;
        const int h = item % H, seq = item / H;
        const int qc = h * DH, kc = d_model + h * DH, vc = 2 * d_model + h * DH;
        if (it > 0) mbar_wait(bar(B_QKFREE), prev);  // previous item: S_1 done and tile-0 outputs left the Q buffer
        mbar_expect_tx(bar(B_Q0K), SQ_BYTES + 2 * kv_blk);
        tma_load_3d(sQ, &tm_q, bar(B_Q0K), qc, 0, seq);
        tma_load_3d(sQ + 16384, &tm_q, bar(B_Q0K), qc + 64, 0, seq);
        tma_load_3d(sK, &tm_kv, bar(B_Q0K), kc, 0, seq);
        tma_load_3d(sK + kv_blk, &tm_kv, bar(B_Q0K), kc + 64, 0, seq);
        if (it > 0) mbar_wait(bar(two ? B_O1 : B_O0), prev);  // previous item's last P.V has consumed V
        mbar_expect_tx(bar(B_V), 2 * kv_blk);
        tma_load_3d(sV, &tm_kv, bar(B_V), vc, 0, seq);
        tma_load_3d(sV + kv_blk, &tm_kv, bar(B_V), vc + 64, 0, seq);
        if (two) {
          mbar_wait(bar(B_S0), ph);  // S_0's MMAs have consumed Q_0
          mbar_expect_tx(bar(B_Q1), SQ_BYTES);
          tma_load_3d(sQ, &tm_q, bar(B_Q1), qc, 128, seq);
          tma_load_3d(sQ + 16384, &tm_q, bar(B_Q1), qc + 64, 128, seq);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===== MMA issuer =====
      auto qk = [&](uint32_t tmem_d, int n) {  // S = Q K[0:n]^T, K dimension = head dim (2 blocks x 4 steps of 16)
        const uint32_t idesc = make_idesc(128, n, 0);
#pragma unroll
        for (int kb = 0; kb < 2; ++kb)
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma(tmem_d, make_desc(sQ + kb * 16384, 16, 1024) + 2u * k, make_desc(sK + kb * kv_blk, 16, 1024) + 2u * k, idesc,
                 (kb | k) != 0 ? 1u : 0u);
      };
      auto pv = [&](int nkeys) {  // O = P[:, 0:nkeys] V[0:nkeys, :], K dimension = keys (steps of 16)
        const uint32_t idesc = make_idesc(128, DH, 1);
        for (int ks = 0; ks < nkeys / 16; ++ks) {
          const uint64_t adesc = make_desc(sP + (ks >> 2) * 16384, 16, 1024) + 2u * (ks & 3);
          const uint64_t bdesc = make_desc(sV + ks * 2048, kv_blk, 1024);  // 16 key rows x 128 B further per step
          umma(tO, adesc, bdesc, idesc, ks != 0 ? 1u : 0u);
        }
      };
      int it = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
        const uint32_t ph = it & 1, prev = ph ^ 1u;
        mbar_wait(bar(B_Q0K), ph);
        tcgen05_fence_after();
        qk(tS0, n0);
        umma_commit(bar(B_S0));
        if (two) {
          mbar_wait(bar(B_Q1), ph);
          tcgen05_fence_after();
          qk(tS1, Tk);
          umma_commit(bar(B_S1));
        }
        mbar_wait(bar(B_P0), ph);
        mbar_wait(bar(B_V), ph);
        if (it > 0) mbar_wait(bar(two ? B_OE1 : B_OE0), prev);  // previous item's last epilogue has drained O
        tcgen05_fence_after();
        pv(n0);
        umma_commit(bar(B_O0));
        if (two) {
          mbar_wait(bar(B_P1), ph);
          mbar_wait(bar(B_OE0), ph);
          tcgen05_fence_after();
          pv(Tk);
          umma_commit(bar(B_O1));
        }
      }
    }
  } else {
    // ===== softmax + epilogue: two threads per query row (warpgroup 0: warps 2-5, warpgroup 1: warps 6-9), each owns a
    //       contiguous half of the key columns / output columns; partial row max and sum are exchanged through smem =====
    const int quad = warp & 3;                                    // TMEM lane quadrant this warp may access
    const int wg = (warp - 2) >> 2;
    const int r = quad * 32 + lane;                               // row inside the tile = TMEM lane
    const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t p_row = sP + r * 128;
    const int rsw = r & 7;

    // S columns [c_lo, c_hi) of this thread, multiples of 16
    auto split = [&](int ncols, int& c_lo, int& c_hi) {
      const int half = ((ncols / 16 + 1) / 2) * 16;
      c_lo = wg == 0 ? 0 : half;
      c_hi = wg == 0 ? half : ncols;
    };
    // Both passes stream S from TMEM in 16-column chunks with the next chunk's tcgen05.ld in flight during the math
    // (two statically indexed register buffers).
    auto row_max = [&](uint32_t tS, int c_lo, int c_hi, const float* tb) {
      float m = -INFINITY;
      uint32_t va[16], vb[16];
      auto chunk = [&](const uint32_t (&v)[16], int c0) {
#pragma unroll
        for (int i = 0; i < 16; ++i) m = fmaxf(m, fmaf(__uint_as_float(v[i]), scale2, tb[-(c0 + i)]));
      };
      if (c_lo < c_hi) tmem_ld16(tS + lane_off + c_lo, va);
      for (int c0 = c_lo; c0 < c_hi; c0 += 32) {
        tmem_ld_wait();
        if (c0 + 16 < c_hi) tmem_ld16(tS + lane_off + c0 + 16, vb);
        chunk(va, c0);
        if (c0 + 16 < c_hi) {
          tmem_ld_wait();
          if (c0 + 32 < c_hi) tmem_ld16(tS + lane_off + c0 + 32, va);
          chunk(vb, c0 + 16);
        }
      }
      return m;
    };
    auto exp_store = [&](uint32_t tS, int c_lo, int c_hi, const float* tb, float m) {  // row-sum part; P (bf16) -> smem
      float l = 0.f;
      uint32_t va[16], vb[16];
      auto chunk = [&](const uint32_t (&v)[16], int c0) {
        float p[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          p[i] = fast_exp2(fmaf(__uint_as_float(v[i]), scale2, tb[-(c0 + i)] - m));
          l += p[i];
        }
        const uint32_t blk = p_row + (c0 >> 6) * 16384;
        const int ch = (c0 & 63) >> 3;  // 16-byte chunk (8 keys) inside the 128-byte row
        st_shared_v4(blk + ((ch ^ rsw) << 4), pack2(p[0], p[1]), pack2(p[2], p[3]), pack2(p[4], p[5]), pack2(p[6], p[7]));
        st_shared_v4(blk + (((ch + 1) ^ rsw) << 4), pack2(p[8], p[9]), pack2(p[10], p[11]), pack2(p[12], p[13]), pack2(p[14], p[15]));
      };
      if (c_lo < c_hi) tmem_ld16(tS + lane_off + c_lo, va);
      for (int c0 = c_lo; c0 < c_hi; c0 += 32) {
        tmem_ld_wait();
        if (c0 + 16 < c_hi) tmem_ld16(tS + lane_off + c0 + 16, vb);
        chunk(va, c0);
        if (c0 + 16 < c_hi) {
          tmem_ld_wait();
          if (c0 + 32 < c_hi) tmem_ld16(tS + lane_off + c0 + 32, va);
          chunk(vb, c0 + 16);
        }
      }
      return l;
    };
    // O[:, 64*wg : 64*wg+64] / rowsum -> bf16 -> this warp's 4 KB swizzled staging tile -> TMA store (rows >= T clipped).
    // (Measured: 16-byte stores straight from registers are slower here - 77 vs 70 us per layer.)
    auto epilogue = [&](int tile, float inv, uint32_t sbuf, int h, int seq) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint32_t v[16];
        tmem_ld16(tO + lane_off + wg * 64 + q * 16, v);
        tmem_ld_wait();
        uint32_t o[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = pack2(__uint_as_float(v[2 * i]) * inv, __uint_as_float(v[2 * i + 1]) * inv);
        st_shared_v4(sbuf + lane * 128 + (((2 * q) ^ (lane & 7)) << 4), o[0], o[1], o[2], o[3]);
        st_shared_v4(sbuf + lane * 128 + (((2 * q + 1) ^ (lane & 7)) << 4), o[4], o[5], o[6], o[7]);
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) {
        tma_store_3d(&tm_o, sbuf, h * DH + wg * 64, tile * 128 + quad * 32, seq);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
    };
    const int ew = warp - 2;  // 0..7: staging slice

    int it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      const uint32_t ph = it & 1;
      const int h = item % H, seq = item / H;
      const float* tb0 = tab + h * TAB_FLOATS + r + 127;        // tb[-j] = bias(t - j), t = r
      const float* tb1 = tb0 + 128;                             // t = 128 + r
      int lo0, hi0, lo1, hi1;
      split(n0, lo0, hi0);
      split(Tk, lo1, hi1);
      float* xm0 = xch, *xl0 = xch + 256, *xm1 = xch + 512, *xl1 = xch + 768;  // [2][128] each
      // ---- tile 0: max, exchange, exp + P store ----
      mbar_wait(bar(B_S0), ph);
      tcgen05_fence_after();
      float m0 = row_max(tS0, lo0, hi0, tb0);
      xm0[wg * 128 + r] = m0;
      if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // previous item's tile-1 stores left the P region
      named_bar_sync(1, 256);
      m0 = fmaxf(m0, xm0[(wg ^ 1) * 128 + r]);
      float l0 = exp_store(tS0, lo0, hi0, tb0, m0);
      fence_async_smem();
      tcgen05_fence_before();
      mbar_arrive(bar(B_P0));
      xl0[wg * 128 + r] = l0;
      // ---- tile 1: max pass while the tensor core runs P_0 V ----
      float m1 = 0.f;
      if (two) {
        mbar_wait(bar(B_S1), ph);  // also: Q_1 has been consumed, the Q buffer can stage the tile-0 outputs
        tcgen05_fence_after();
        m1 = row_max(tS1, lo1, hi1, tb1);
        xm1[wg * 128 + r] = m1;
      }
      named_bar_sync(2, 256);
      l0 += xl0[(wg ^ 1) * 128 + r];
      if (two) m1 = fmaxf(m1, xm1[(wg ^ 1) * 128 + r]);
      mbar_wait(bar(B_O0), ph);  // O_0 complete; P buffer free
      tcgen05_fence_after();
      epilogue(0, 1.f / l0, sQ + ew * 4096u, h, seq);
      tcgen05_fence_before();
      mbar_arrive(bar(B_OE0));  // O accumulator drained
      if (two) {
        float l1 = exp_store(tS1, lo1, hi1, tb1, m1);
        fence_async_smem();
        tcgen05_fence_before();
        mbar_arrive(bar(B_P1));
        xl1[wg * 128 + r] = l1;
        if (lane == 0) {  // tile-0 stores have left the Q buffer: the producer may prefetch the next item's Q_0 / K
          asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          mbar_arrive(bar(B_QKFREE));
        }
        named_bar_sync(3, 256);
        l1 += xl1[(wg ^ 1) * 128 + r];
        mbar_wait(bar(B_O1), ph);
        tcgen05_fence_after();
        epilogue(1, 1.f / l1, sP + ew * 4096u, h, seq);  // P buffer is free once O_1 is complete
        tcgen05_fence_before();
        mbar_arrive(bar(B_OE1));
      } else if (lane == 0) {
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        mbar_arrive(bar(B_QKFREE));
      }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }

  pdl_trigger();
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                        const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                        CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
PFN_tmapEncodeTiled encode_fn() {
  static PFN_tmapEncodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_tmapEncodeTiled>(p);
  }
  return fn;
}

// 3-D bf16 map over a [sequences, T (t_stride rows apart), cols] activation: rows >= T are out of bounds (zero on
// load, clipped on store), so neighbouring sequences never leak into a tile.
bool make_map3(CUtensorMap* out, const void* ptr, int64_t cols, int64_t T, int64_t S, int64_t ld, int64_t t_stride, int box_rows) {
  PFN_tmapEncodeTiled enc = encode_fn();
  if (!enc) return false;
  cuuint64_t dims[3] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(T), static_cast<cuuint64_t>(S)};
  cuuint64_t strides[2] = {static_cast<cuuint64_t>(ld) * 2, static_cast<cuuint64_t>(t_stride) * ld * 2};
  cuuint32_t box[3] = {64, static_cast<cuuint32_t>(box_rows), 1};
  cuuint32_t estr[3] = {1, 1, 1};
  return enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace

int fdm_attention_tc_try(const fdm_attn_args& a, cudaStream_t stream, bool* handled) {
  *handled = false;
  static const bool enabled = [] { const char* e = getenv("FDM_B200_ATTN_TC"); return !(e && e[0] == '0'); }();  // (default path: attention_tc2.cu; this kernel runs when that one declines or FDM_B200_ATTN_TC=1)
  if (!enabled || a.dtype != FDM_BF16 || a.dh != DH || a.bias_mode != 1 || a.T > TK_MAX || a.T < 16 || a.H > H_MAX) return 0;
  // Q, K, V must be the three column groups of ONE packed [rows, 3d] buffer (what the denoiser's in_proj GEMM writes)
  const int64_t d = a.H * a.dh;
  const __nv_bfloat16* q = reinterpret_cast<const __nv_bfloat16*>(a.Q);
  if (reinterpret_cast<const __nv_bfloat16*>(a.K) != q + d || reinterpret_cast<const __nv_bfloat16*>(a.V) != q + 2 * d) return 0;
  if (a.ldq != a.ldk || a.ldq != a.ldv || a.ldq % 8 || a.ldo % 8 || a.B > 65535) return 0;
  if ((reinterpret_cast<uintptr_t>(a.Q) | reinterpret_cast<uintptr_t>(a.O)) & 15u) return 0;
  const int T = static_cast<int>(a.T), Tk = (T + 15) / 16 * 16;
  CUtensorMap tq, tkv, to;
  if (!make_map3(&tq, a.Q, 3 * d, T, a.B, a.ldq, a.t_stride, 128) || !make_map3(&tkv, a.Q, 3 * d, T, a.B, a.ldq, a.t_stride, Tk) ||
      !make_map3(&to, a.O, d, T, a.B, a.ldo, a.t_stride, 32))
    return 0;
  static bool attr = false;
  if (!attr) {
    FDM_CHECK_CUDA(cudaFuncSetAttribute(attn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    attr = true;
  }
  *handled = true;
  const int64_t n_items = a.H * a.B;  // (sequence, head) pairs, persistent CTAs
  const int grid = static_cast<int>(n_items < fdm_sm_count() ? n_items : fdm_sm_count());
  FDM_CHECK_CUDA(fdm_launch_pdl(attn_tc_kernel, dim3(grid), dim3(NUM_THREADS), SMEM_BYTES, stream, 1, tq, tkv, to, T, Tk,
                                static_cast<int>(d), static_cast<int>(a.H), static_cast<int>(n_items), a.scale * 1.4426950408889634f,
                                a.slopes, a.period));
  return 0;
}
