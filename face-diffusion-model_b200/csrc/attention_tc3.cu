// General tcgen05/TMEM self-attention (bf16): head dim 64 / 128 / 256, any number of keys (causal + periodic ALiBi up to
// T = 512), for every shape attention_tc2.cu (head dim 128, T <= 208, all keys at once) does not take:
//   * BIWI's FDM, 4 heads x 256 (models/fdm.py:10-52), T = 149 at 6 s - 22 % of the BIWI denoising step on mma.sync before;
//   * HuBERT-large / wav2vec2-base encoder attention, 16 / 12 heads x 64, N = 198 ... 498 (models/hubert.py:91-137);
//   * the EVQ-VAE transformer (8 x 128, unmasked, models/lib/base_models.py:138-174) and the VOCASET FDM at 10 s (T = 498).
//
// Work item = (sequence, head, 128-row query tile), persistent CTAs, heaviest (latest causal) tiles first. Keys stream
// through a shared-memory ring in blocks of 64; a score tile spans ST blocks (128 keys for head dim <= 128, 64 for head dim
// 256 where shared memory only holds 64-key P tiles). Exact two-pass softmax, no rescaling of the accumulator:
//   pass 1   S_t = Q K_t^T (tcgen05.mma into a double-buffered TMEM tile) -> row maximum of the raw scores
//   pass 2   S_t again -> P_t = exp2(S_t * scale + bias - max) as bf16 in the K-major swizzled A-operand layout (double
//            buffered) -> O += P_t V_t (V as an MN-major B operand) ; O / rowsum -> bf16 -> TMA store
// The second Q K^T costs tensor-pipe time the kernel has to spare (the softmax warps are the bottleneck: TMEM read bandwidth
// per lane quadrant and the MUFU pipe, see attention_tc2.cu) and keeps the arithmetic identical to the single-tile kernel:
// the maximum is exact, so no lazy-rescale path and no data-dependent timing.
// Warp roles (19 warps): 0 K/V producer (TMA), 1 MMA issuer, 2-17 softmax + epilogue (TMEM lane quadrant = warp % 4, four
// threads per query row, 16 keys of a block each), 18 query loads + output stores (so the K/V ring keeps prefetching the
// next item while an output tile drains). Every wait is an mbarrier try_wait loop with a watchdog trap.
// Head dim <= 128: two query buffers and the output tile staged in the (by then idle) P buffers, so an item switch costs no
// load / store round trip; head dim 256: one 64 KB query buffer that doubles as the staging tile.
// TMEM: O [0, DH) | S_0 [256, 256 + 64 ST) | S_1 [256 + 64 ST, 256 + 128 ST).
#include "tc_common.cuh"
#include <stdlib.h>

using namespace tc;

namespace {

constexpr int BK = 64;  // keys per block
constexpr int SM_WARPS = 16;
constexpr int SM_THREADS = SM_WARPS * 32;
constexpr int NUM_THREADS = 64 + SM_THREADS + 32;
constexpr int TAB_FLOATS = 640;  // bias table of the item's head: index = (t - j) + 128, t - j in [-127, 511]
constexpr int T_MAX_CAUSAL = 512;
constexpr int MAX_SLOTS = 8;
enum { B_QFULL = 0, B_QEMPTY = 2, B_STGFREE = 4, B_OFULL, B_OEMPTY, B_OUTFULL, B_SFULL, B_SEMPTY = B_SFULL + 2, B_PFULL = B_SEMPTY + 2, B_PEMPTY = B_PFULL + 2,
       B_KVFULL = B_PEMPTY + 2, B_KVEMPTY = B_KVFULL + MAX_SLOTS, NUM_BARS = B_KVEMPTY + MAX_SLOTS };

template <int DH>
struct Cfg3 {
  static constexpr int G = DH / 64;               // 64-column groups of the head dim
  static constexpr int Q_BYTES = 128 * DH * 2;    // query tile: G k-blocks of [128 rows][128 B]; reused as the output staging tile
  static constexpr int SLOT_BYTES = BK * DH * 2;  // one K or V block: G groups of [64 keys][128 B]
  static constexpr int NSLOT = DH == 256 ? 3 : (DH == 128 ? 5 : 8);
  // head dim <= 128: two query buffers (the next item's tile lands while this one computes) and the output tile is staged in
  // the P buffers; head dim 256: one query buffer that doubles as the output staging tile (64 KB each)
  static constexpr int QBUF = DH == 256 ? 1 : 2;
  static constexpr int ST = DH == 256 ? 1 : 2;    // key blocks (ring slots) per score tile
  static constexpr int KT = BK * ST;              // keys per score tile
  static constexpr int P_BYTES = 128 * KT * 2;    // ST k-blocks of [128 rows][128 B]
  static constexpr int OFF_KV = QBUF * Q_BYTES;
  static constexpr int OFF_P = OFF_KV + NSLOT * SLOT_BYTES;
  static constexpr int OFF_TAB = OFF_P + 2 * P_BYTES;
  static constexpr int OFF_XCH = OFF_TAB + TAB_FLOATS * 4;
  static constexpr int OFF_BAR = OFF_XCH + 2 * 4 * 128 * 4;
  static constexpr int SMEM_BYTES = OFF_BAR + 8 * NUM_BARS + 16 + 1024;  // + TMEM slot + manual 1024-byte alignment
  static_assert(SMEM_BYTES <= 232448, "shared memory budget");
  static_assert(NSLOT <= MAX_SLOTS, "ring too deep");
  static_assert(QBUF == 1 || Q_BYTES <= 2 * P_BYTES, "the output tile must fit the P buffers");
};

struct Ring {  // position in a ring of N single-use-per-lap buffers
  int idx = 0;
  uint32_t lap = 0;
  __device__ __forceinline__ void next(int n) {
    if (++idx == n) { idx = 0; lap ^= 1u; }
  }
};

template <int DH, bool CAUSAL>
__global__ void __launch_bounds__(NUM_THREADS, 1)
attn_tc3_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_o, const int T, const int H,
                const int n_sh, const int n_qt, const int n_items, const float scale2, const float* __restrict__ slopes,
                const int period) {
  using C = Cfg3<DH>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const uint32_t sQ = base, sKV = base + C::OFF_KV, sP = base + C::OFF_P;
  float* tab = reinterpret_cast<float*>(smem + C::OFF_TAB);
  float* xch = reinterpret_cast<float*>(smem + C::OFF_XCH);  // [2][4 parts][128 rows]: row maxima, row sums
  const uint32_t bars = base + C::OFF_BAR;
  auto bar = [&](int i) { return bars + 8u * i; };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + C::OFF_BAR + 8 * NUM_BARS);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_k);
    tma_prefetch_desc(&tm_v);
  } else if (warp == 18 && lane == 0) {
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_o);
  } else if (warp == 1 && lane == 0) {
    for (int i = 0; i < NUM_BARS; ++i) {
      // barriers the softmax warps arrive on take ONE arrival per warp (lane 0 after __syncwarp)
      const bool sm = i == B_OEMPTY || i == B_OUTFULL || i == B_SEMPTY || i == B_SEMPTY + 1 || i == B_PFULL || i == B_PFULL + 1;  // (B_QEMPTY, B_STGFREE: one arrival)
      mbar_init(bar(i), sm ? SM_WARPS : 1);
    }
    fence_barrier_init();
  } else if (warp == 2) {
    tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_slot)), 512u);
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tO = tmem, tS = tmem + 256;
  pdl_wait();

  // item -> (sequence, head, query tile): tiles in DEscending order first (a causal tile's cost grows with its index)
  auto decode = [&](int item, int& seq, int& h, int& qt) {
    const int sh = item % n_sh;
    qt = n_qt - 1 - item / n_sh;
    seq = sh / H;
    h = sh - seq * H;
  };
  auto blocks_of = [&](int qt) {  // key blocks a query tile attends to
    const int keys = CAUSAL ? min(T, 128 * qt + 128) : T;
    return (keys + BK - 1) / BK;
  };

  if (warp == 0) {
    if (lane == 0) {
      // ===== K/V producer: pass 1 wants the K blocks in order; pass 2 consumes tile by tile K(0), K(1), V(0), K(2), V(1), ... =====
      Ring r;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        int seq, h, qt;
        decode(item, seq, h, qt);
        const int n = blocks_of(qt);
        auto load = [&](const CUtensorMap* m, int blk) {
          mbar_wait(bar(B_KVEMPTY + r.idx), r.lap ^ 1u, "attn_tc3 producer");
          mbar_expect_tx(bar(B_KVFULL + r.idx), C::SLOT_BYTES);
          const uint32_t dst = sKV + r.idx * C::SLOT_BYTES;
#pragma unroll
          for (int g = 0; g < C::G; ++g) tma_load_3d(dst + g * (BK * 128), m, bar(B_KVFULL + r.idx), h * DH + 64 * g, BK * blk, seq);
          r.next(C::NSLOT);
        };
        for (int b = 0; b < n; ++b) load(&tm_k, b);
        for (int b0 = 0; b0 < n; b0 += C::ST) {  // blocks [b0, b1) = one score tile
          const int b1 = min(b0 + C::ST, n);
          for (int b = b0; b < b1; ++b) load(&tm_k, b);
          if (b0 > 0)
            for (int b = b0 - C::ST; b < b0; ++b) load(&tm_v, b);
        }
        for (int b = (n - 1) / C::ST * C::ST; b < n; ++b) load(&tm_v, b);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===== MMA issuer =====
      constexpr uint32_t idesc_qk = make_idesc_bf16(128, BK, 0);
      constexpr uint32_t idesc_pv = make_idesc_bf16(128, DH, 1);
      Ring r, rs, rp;  // K/V ring, S buffers, P buffers
      uint32_t sQcur = sQ;
      auto qk = [&](int nb) {  // S[rs.idx][:, 64 s2 ...] = Q K_blk^T for the nb blocks of a score tile (ring slots r.idx ...)
        mbar_wait(bar(B_SEMPTY + rs.idx), rs.lap ^ 1u, "attn_tc3 mma S empty");
        for (int s2 = 0; s2 < nb; ++s2) {
          mbar_wait(bar(B_KVFULL + r.idx), r.lap, "attn_tc3 mma K");
          tcgen05_fence_after();
          const uint32_t kb0 = sKV + r.idx * C::SLOT_BYTES;
#pragma unroll
          for (int g = 0; g < C::G; ++g)
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16(tS + C::KT * rs.idx + BK * s2, make_desc_sw128(sQcur + g * 16384, 16, 1024) + 2u * k,
                        make_desc_sw128(kb0 + g * (BK * 128), 16, 1024) + 2u * k, idesc_qk, (g | k) != 0 ? 1u : 0u);
          umma_commit(bar(B_KVEMPTY + r.idx));
          r.next(C::NSLOT);
        }
        umma_commit(bar(B_SFULL + rs.idx));
        rs.next(2);
      };
      int it = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
        int seq, h, qt;
        decode(item, seq, h, qt);
        const int n = blocks_of(qt);
        auto pv = [&](int tl, int nb) {  // O (+)= P[rp.idx] V_tile
          mbar_wait(bar(B_PFULL + rp.idx), rp.lap, "attn_tc3 mma P");
          if (tl == 0 && it > 0) mbar_wait(bar(B_OEMPTY), (it - 1) & 1, "attn_tc3 mma O empty");  // previous epilogue drained O
          for (int s2 = 0; s2 < nb; ++s2) {
            mbar_wait(bar(B_KVFULL + r.idx), r.lap, "attn_tc3 mma V");
            tcgen05_fence_after();
            const uint32_t vb = sKV + r.idx * C::SLOT_BYTES, pb = sP + rp.idx * C::P_BYTES + s2 * 16384;
#pragma unroll
            for (int ks = 0; ks < BK / 16; ++ks)
              umma_bf16(tO, make_desc_sw128(pb, 16, 1024) + 2u * ks, make_desc_sw128(vb + ks * 2048, BK * 128, 1024), idesc_pv,
                        (tl | s2 | ks) != 0 ? 1u : 0u);
            umma_commit(bar(B_KVEMPTY + r.idx));
            r.next(C::NSLOT);
          }
          umma_commit(bar(B_PEMPTY + rp.idx));
          rp.next(2);
        };
        const int nt = (n + C::ST - 1) / C::ST;
        auto nb_of = [&](int tl) { return min(C::ST, n - C::ST * tl); };
        const int qb = C::QBUF == 2 ? (it & 1) : 0;
        sQcur = sQ + qb * C::Q_BYTES;
        mbar_wait(bar(B_QFULL + qb), C::QBUF == 2 ? ((it >> 1) & 1) : (it & 1), "attn_tc3 mma Q");
        for (int tl = 0; tl < nt; ++tl) qk(nb_of(tl));  // pass 1
        for (int tl = 0; tl < nt; ++tl) {               // pass 2
          qk(nb_of(tl));
          if (tl == nt - 1) umma_commit(bar(B_QEMPTY + qb));  // every Q K^T of this item has read the query buffer
          if (tl > 0) pv(tl - 1, C::ST);
        }
        pv(nt - 1, nb_of(nt - 1));
        umma_commit(bar(B_OFULL));
      }
    }
  } else if (warp == 18) {
    if (lane == 0) {
      // ===== query loads + output stores =====
      auto load_q = [&](int item, int qb) {
        int seq, h, qt;
        decode(item, seq, h, qt);
        mbar_expect_tx(bar(B_QFULL + qb), C::Q_BYTES);
#pragma unroll
        for (int g = 0; g < C::G; ++g) tma_load_3d(sQ + qb * C::Q_BYTES + g * 16384, &tm_q, bar(B_QFULL + qb), h * DH + 64 * g, 128 * qt, seq);
      };
      int it = 0;
      if (C::QBUF == 2 && static_cast<int>(blockIdx.x) < n_items) load_q(blockIdx.x, 0);
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
        int seq, h, qt;
        decode(item, seq, h, qt);
        const int next = item + gridDim.x;
        if (C::QBUF == 2) {
          // the next item's query tile into the other buffer as soon as the item before this one has released it
          if (next < n_items) {
            const int nb = (it + 1) & 1;
            if (it >= 1) mbar_wait(bar(B_QEMPTY + nb), ((it - 1) >> 1) & 1, "attn_tc3 Q empty");
            load_q(next, nb);
          }
        } else {
          load_q(item, 0);  // (single buffer: free since the previous item's output store has read it)
          if (next < n_items) {  // the next query tile into L2
            int s2, h2, q2;
            decode(next, s2, h2, q2);
#pragma unroll
            for (int g = 0; g < C::G; ++g) tma_prefetch_l2_3d(&tm_q, h2 * DH + 64 * g, 128 * q2, s2);
          }
        }
        mbar_wait(bar(B_OUTFULL), it & 1, "attn_tc3 store");
        const uint32_t stg = C::QBUF == 2 ? sP : sQ;
#pragma unroll
        for (int g = 0; g < C::G; ++g) tma_store_3d(&tm_o, stg + g * 16384, h * DH + 64 * g, 128 * qt, seq);
        bulk_commit();
        bulk_wait_read_all();  // the staging tile has been read
        if (C::QBUF == 2) mbar_arrive(bar(B_STGFREE));  // ... the P buffers may take the next item's probabilities
      }
      bulk_wait_all();
    }
  } else {
    // ===== softmax + epilogue =====
    const int quad = warp & 3;
    const int part = (warp - 2) >> 2;  // which 16 keys of a block / which DH / 4 output columns
    const int r = quad * 32 + lane;    // row inside the tile = TMEM lane
    const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
    const int rsw = r & 7;
    const int st = threadIdx.x - 64;   // 0 .. 511
    float* xm = xch, *xl = xch + 512;
    Ring rs, rp;
    int it = 0, tab_h = -1;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      int seq, h, qt;
      decode(item, seq, h, qt);
      const int n = blocks_of(qt);
      const int t = 128 * qt + r;  // query index of this thread's row
      if (CAUSAL && h != tab_h) {   // bias table of this head, log2 domain; -inf above the diagonal (warp-uniform branch)
        const float sl = -(slopes[h] * 1.4426950408889634f);
        for (int i = st; i < TAB_FLOATS; i += SM_THREADS) {
          const int delta = i - 128;
          tab[i] = delta < 0 ? -INFINITY : sl * static_cast<float>(delta / period);
        }
        tab_h = h;
      }
      // keys of block b this thread owns: j = 64 b + 16 part + e; a block is `full` when every key is valid for every row
      const int full_blocks = CAUSAL ? min(128 * qt + 1, T) / BK : T / BK;
      const int nt = (n + C::ST - 1) / C::ST;
      uint32_t s[C::ST][16];
      auto next_scores = [&](int nb) {  // this thread's 16 keys of each of the tile's nb blocks, all loads in flight before one wait
        mbar_wait(bar(B_SFULL + rs.idx), rs.lap, "attn_tc3 softmax S");
        tcgen05_fence_after();
#pragma unroll
        for (int s2 = 0; s2 < C::ST; ++s2)
          if (s2 < nb) tmem_ld16(tS + lane_off + C::KT * rs.idx + BK * s2 + 16 * part, s[s2]);
        tmem_ld_wait();
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(B_SEMPTY + rs.idx));
        rs.next(2);
      };
      // ---- pass 1: row maximum of the raw scores (the ALiBi bias is <= 0, so max_j s_j * scale bounds the logits) ----
      float mloc = -INFINITY;
#pragma unroll 1
      for (int tl = 0; tl < nt; ++tl) {
        const int nb = min(C::ST, n - C::ST * tl);
        next_scores(nb);
#pragma unroll
        for (int s2 = 0; s2 < C::ST; ++s2) {
          const int b = C::ST * tl + s2;
          if (s2 < nb) {
            if (b < full_blocks) {
#pragma unroll
              for (int e = 0; e < 16; ++e) mloc = fmaxf(mloc, __uint_as_float(s[s2][e]));
            } else {
              const int j0 = BK * b + 16 * part;
#pragma unroll
              for (int e = 0; e < 16; ++e) {
                const int j = j0 + e;
                const bool ok = j < T && (!CAUSAL || j <= t);
                mloc = fmaxf(mloc, ok ? __uint_as_float(s[s2][e]) : -INFINITY);
              }
            }
          }
        }
      }
      xm[part * 128 + r] = mloc;
      named_bar_sync(1, SM_THREADS);
      float m2 = fmaxf(fmaxf(xm[r], xm[128 + r]), fmaxf(xm[256 + r], xm[384 + r])) * scale2;
      if (m2 == -INFINITY) m2 = 0.f;  // (rows past the sequence end)
      // ---- pass 2: P = exp2(s * scale2 + bias - m2) -> bf16, swizzled K-major [128 rows][64 keys] ----
      float l = 0.f;
      const float* tb = tab + 128 + t;  // tb[-j] = bias(t - j)
#pragma unroll 1
      for (int tl = 0; tl < nt; ++tl) {
        const int nb = min(C::ST, n - C::ST * tl);
        next_scores(nb);
        if (C::QBUF == 2 && tl == 0 && it > 0) mbar_wait(bar(B_STGFREE), (it - 1) & 1, "attn_tc3 softmax staging");  // (long since done)
        mbar_wait(bar(B_PEMPTY + rp.idx), rp.lap ^ 1u, "attn_tc3 softmax P empty");
#pragma unroll
        for (int s2 = 0; s2 < C::ST; ++s2) {
          if (s2 < nb) {
            const int b = C::ST * tl + s2;
            const int j0 = BK * b + 16 * part;
            float p[16];
            if (CAUSAL) {
              const float* tk = tb - j0;
#pragma unroll
              for (int e = 0; e < 16; ++e) p[e] = fast_exp2(fmaf(__uint_as_float(s[s2][e]), scale2, tk[-e] - m2));
            } else if (b < full_blocks) {
#pragma unroll
              for (int e = 0; e < 16; ++e) p[e] = fast_exp2(fmaf(__uint_as_float(s[s2][e]), scale2, -m2));
            } else {
#pragma unroll
              for (int e = 0; e < 16; ++e) p[e] = j0 + e < T ? fast_exp2(fmaf(__uint_as_float(s[s2][e]), scale2, -m2)) : 0.f;
            }
#pragma unroll
            for (int e = 0; e < 16; ++e) l += p[e];
            const uint32_t prow = sP + rp.idx * C::P_BYTES + s2 * 16384 + r * 128;
            st_shared_v4(prow + (((2 * part) ^ rsw) << 4), pack_bf16x2(p[0], p[1]), pack_bf16x2(p[2], p[3]), pack_bf16x2(p[4], p[5]),
                         pack_bf16x2(p[6], p[7]));
            st_shared_v4(prow + (((2 * part + 1) ^ rsw) << 4), pack_bf16x2(p[8], p[9]), pack_bf16x2(p[10], p[11]),
                         pack_bf16x2(p[12], p[13]), pack_bf16x2(p[14], p[15]));
          }
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(B_PFULL + rp.idx));
        rp.next(2);
      }
      xl[part * 128 + r] = l;
      named_bar_sync(2, SM_THREADS);
      const float inv = 1.f / ((xl[r] + xl[128 + r]) + (xl[256 + r] + xl[384 + r]));
      // ---- epilogue: O[row, part * DH/4 ...] / rowsum -> bf16 -> staging tile (query buffer, [G][128 rows][128 B] swizzled) ----
      mbar_wait(bar(B_OFULL), it & 1, "attn_tc3 softmax O");
      tcgen05_fence_after();
      constexpr int NC = DH / 64;  // 16-column chunks per thread
      uint32_t v[NC][16];
#pragma unroll
      for (int c = 0; c < NC; ++c) tmem_ld16(tO + lane_off + part * (DH / 4) + 16 * c, v[c]);
      tmem_ld_wait();
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        const int col = part * (DH / 4) + 16 * c;  // first of 16 output columns
        const uint32_t grp = (C::QBUF == 2 ? sP : sQ) + (col >> 6) * 16384 + r * 128;
        const int ch = (col & 63) >> 3;
        uint32_t o[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = pack_bf16x2(__uint_as_float(v[c][2 * i]) * inv, __uint_as_float(v[c][2 * i + 1]) * inv);
        st_shared_v4(grp + ((ch ^ rsw) << 4), o[0], o[1], o[2], o[3]);
        st_shared_v4(grp + (((ch + 1) ^ rsw) << 4), o[4], o[5], o[6], o[7]);
      }
      fence_async_smem();
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(bar(B_OUTFULL));
        mbar_arrive(bar(B_OEMPTY));
      }
    }
  }

  pdl_trigger();
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    tmem_dealloc(tmem, 512u);
  }
}

template <int DH, bool CAUSAL>
int launch3(const fdm_attn_args& a, cudaStream_t stream) {
  using C = Cfg3<DH>;
  static bool attr = false;
  if (!attr) {
    FDM_CHECK_CUDA(cudaFuncSetAttribute(attn_tc3_kernel<DH, CAUSAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    attr = true;
  }
  const int T = static_cast<int>(a.T);
  const int64_t d = a.H * a.dh;
  CUtensorMap tq, tk, tv, to;
  FDM_CHECK_ARG(make_map3_bf16(&tq, a.Q, d, T, a.B, a.ldq, a.t_stride, 128) && make_map3_bf16(&tk, a.K, d, T, a.B, a.ldk, a.t_stride, BK) &&
                    make_map3_bf16(&tv, a.V, d, T, a.B, a.ldv, a.t_stride, BK) && make_map3_bf16(&to, a.O, d, T, a.B, a.ldo, a.t_stride, 128),
                "fdm_self_attention: cuTensorMapEncodeTiled failed (T=%d d=%lld)", T, (long long)d);
  const int n_qt = (T + 127) / 128;
  const int64_t n_sh = a.B * a.H, n_items = n_sh * n_qt;
  const int grid = static_cast<int>(n_items < fdm_sm_count() ? n_items : fdm_sm_count());
  FDM_CHECK_CUDA(fdm_launch_pdl(attn_tc3_kernel<DH, CAUSAL>, dim3(grid), dim3(NUM_THREADS), C::SMEM_BYTES, stream, 1, tq, tk, tv, to, T,
                                static_cast<int>(a.H), static_cast<int>(n_sh), n_qt, static_cast<int>(n_items),
                                a.scale * 1.4426950408889634f, a.slopes, a.period));
  return 0;
}

}  // namespace

int fdm_attention_tc3_try(const fdm_attn_args& a, cudaStream_t stream, bool* handled) {
  *handled = false;
  static const bool on = [] { const char* e = getenv("FDM_B200_ATTN_TC3"); return !(e && e[0] == '0'); }();
  if (!on || a.dtype != FDM_BF16 || a.scale <= 0.f || a.T < 1) return 0;
  if (a.dh != 64 && a.dh != 128 && a.dh != 256) return 0;
  // head dim 64 (audio encoders): the mma.sync kernel is still faster there (68 vs 106 us at 64 x 16 x 198): opt-in
  static const bool dh64 = [] { const char* e = getenv("FDM_B200_ATTN_TC3_DH64"); return e && e[0] == '1'; }();
  if (a.dh == 64 && !dh64) return 0;
  if (a.bias_mode != 0 && a.bias_mode != 1) return 0;
  if (a.bias_mode == 1 && a.T > T_MAX_CAUSAL) return 0;
  if (a.ldq % 8 || a.ldk % 8 || a.ldv % 8 || a.ldo % 8) return 0;
  if ((reinterpret_cast<uintptr_t>(a.Q) | reinterpret_cast<uintptr_t>(a.K) | reinterpret_cast<uintptr_t>(a.V) | reinterpret_cast<uintptr_t>(a.O)) & 15u)
    return 0;
  if (a.B * a.H * ((a.T + 127) / 128) >= (1ll << 31)) return 0;
  *handled = true;
  const bool causal = a.bias_mode == 1;
  switch (a.dh) {
    case 64: return causal ? launch3<64, true>(a, stream) : launch3<64, false>(a, stream);
    case 128: return causal ? launch3<128, true>(a, stream) : launch3<128, false>(a, stream);
    default: return causal ? launch3<256, true>(a, stream) : launch3<256, false>(a, stream);
  }
}
