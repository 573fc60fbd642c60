// Tensor-core self-attention for bf16 activations (head dim 64 / 128): flash-style online softmax with
// mma.sync.m16n8k16 (bf16 in, fp32 accumulate), K/V streamed through a double-buffered, XOR-swizzled shared
// memory ring with cp.async, Q fragments and the O accumulator held in registers. The periodic-ALiBi bias and
// the causal mask are evaluated per score from (t - j) with an exact integer reciprocal; causal key blocks
// beyond the diagonal are skipped. CTA = 4 warps x 16 query rows; grid = (ceil(T/64), heads, sequences).
// (T <= 600 here, one head's K/V is at most 150 KB: the work per CTA is too small to amortise a TMEM/tcgen05
//  pipeline, so attention stays on the legacy warp-level MMA path; the GEMMs around it use tcgen05.)
#include "common.cuh"

namespace {

constexpr int QB = 64, KB = 64, THREADS = 128;

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

// tile [rows][DH] bf16, 16-byte chunks XOR-swizzled by (row & 7)
template <int DH>
__device__ __forceinline__ uint32_t tile_off(int row, int chunk) {
  return static_cast<uint32_t>(row * (DH * 2) + ((chunk ^ (row & 7)) << 4));
}

template <int DH>
__device__ __forceinline__ void load_tile(uint32_t smem_base, const __nv_bfloat16* g, int64_t ld, int row0, int n_rows_valid,
                                          int rows) {
  constexpr int CH = DH / 8;  // 16-byte chunks per row
  for (int i = threadIdx.x; i < rows * CH; i += THREADS) {
    const int r = i / CH, c = i - r * CH;
    const bool ok = row0 + r < n_rows_valid;
    cp_async16(smem_base + tile_off<DH>(r, c), g + static_cast<int64_t>(ok ? row0 + r : 0) * ld + c * 8, ok);
  }
}

template <int DH>
__global__ void __launch_bounds__(THREADS) attn_mma_kernel(const fdm_attn_args a, const uint32_t period_magic) {
  constexpr int KS = DH / 16;   // k-steps of Q.K^T
  constexpr int NT = DH / 8;    // n-tiles of P.V
  constexpr int TILE = KB * DH * 2;
  extern __shared__ __align__(128) uint8_t smem[];
  const uint32_t sQ = smem_u32(smem);
  const uint32_t sK = sQ + QB * DH * 2;       // 2 stages
  const uint32_t sV = sK + 2 * TILE;          // 2 stages

  const int T = static_cast<int>(a.T);
  const int q0 = blockIdx.x * QB, h = blockIdx.y;
  const int64_t row0 = static_cast<int64_t>(blockIdx.z) * a.t_stride;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, tq = lane & 3;
  const bool causal = a.bias_mode == 1;
  const float LOG2E = 1.4426950408889634f;
  const float slope2 = causal ? a.slopes[h] * LOG2E : 0.f;
  const float scale2 = a.scale * LOG2E;

  const __nv_bfloat16* Qg = reinterpret_cast<const __nv_bfloat16*>(a.Q) + row0 * a.ldq + static_cast<int64_t>(h) * DH;
  const __nv_bfloat16* Kg = reinterpret_cast<const __nv_bfloat16*>(a.K) + row0 * a.ldk + static_cast<int64_t>(h) * DH;
  const __nv_bfloat16* Vg = reinterpret_cast<const __nv_bfloat16*>(a.V) + row0 * a.ldv + static_cast<int64_t>(h) * DH;

  const int k_end = causal ? min(T, q0 + QB) : T;
  const int n_kb = (k_end + KB - 1) / KB;

  load_tile<DH>(sQ, Qg, a.ldq, q0, T, QB);
  load_tile<DH>(sK, Kg, a.ldk, 0, T, KB);
  load_tile<DH>(sV, Vg, a.ldv, 0, T, KB);
  cp_async_commit();
  if (n_kb > 1) {
    load_tile<DH>(sK + TILE, Kg, a.ldk, KB, T, KB);
    load_tile<DH>(sV + TILE, Vg, a.ldv, KB, T, KB);
  }
  cp_async_commit();
  cp_async_wait<1>();
  __syncthreads();

  // Q fragments for this warp's 16 rows
  uint32_t qf[KS][4];
#pragma unroll
  for (int ks = 0; ks < KS; ++ks) {
    const int r = warp * 16 + (lane & 15);
    const int c = ks * 2 + (lane >> 4);
    ldsm_x4(sQ + tile_off<DH>(r, c), qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3]);
  }

  float o[NT][4];
#pragma unroll
  for (int i = 0; i < NT; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
  const int t_row[2] = {q0 + warp * 16 + g, q0 + warp * 16 + g + 8};

  for (int kb = 0; kb < n_kb; ++kb) {
    const uint32_t kbuf = sK + (kb & 1) * TILE, vbuf = sV + (kb & 1) * TILE;
    const int j0 = kb * KB;
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
    // ---- scale, bias, mask, online softmax (log2 domain) ----
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int n = 0; n < 8; ++n) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int r = e >> 1;
        const int j = j0 + n * 8 + tq * 2 + (e & 1);
        float x = s[n][e] * scale2;
        bool valid = j < T;
        if (causal) {
          const int dlt = t_row[r] - j;
          valid = valid && dlt >= 0;
          const uint32_t q = (static_cast<uint32_t>(dlt < 0 ? 0 : dlt) * period_magic) >> 16;  // floor(dlt / period), exact
          x -= slope2 * static_cast<float>(q);
        }
        x = valid ? x : -INFINITY;
        s[n][e] = x;
        mx[r] = fmaxf(mx[r], x);
      }
    }
    float corr[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
      const float m_new = fmaxf(m_run[r], mx[r]);
      corr[r] = m_new == -INFINITY ? 1.f : exp2f(m_run[r] - m_new);
      m_run[r] = m_new;
    }
    float rs[2] = {0.f, 0.f};
    uint32_t pf[4][4];  // P as A fragments: 4 k-steps of 16 keys
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      float p[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int r = e >> 1;
        p[e] = m_run[r] == -INFINITY ? 0.f : exp2f(s[n][e] - m_run[r]);
        rs[r] += p[e];
      }
      // accumulator tile n covers keys n*8..n*8+7: k-step n/2, low (a0,a1) or high (a2,a3) half
      pf[n >> 1][(n & 1) * 2 + 0] = pack_bf16(p[0], p[1]);
      pf[n >> 1][(n & 1) * 2 + 1] = pack_bf16(p[2], p[3]);
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      rs[r] += __shfl_xor_sync(0xffffffffu, rs[r], 1);
      rs[r] += __shfl_xor_sync(0xffffffffu, rs[r], 2);
      l_run[r] = l_run[r] * corr[r] + rs[r];
    }
#pragma unroll
    for (int i = 0; i < NT; ++i) {
      o[i][0] *= corr[0]; o[i][1] *= corr[0];
      o[i][2] *= corr[1]; o[i][3] *= corr[1];
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
    // ---- pipeline: refill this stage with block kb+2, make block kb+1 visible ----
    __syncthreads();
    if (kb + 2 < n_kb) {
      load_tile<DH>(kbuf, Kg, a.ldk, (kb + 2) * KB, T, KB);
      load_tile<DH>(vbuf, Vg, a.ldv, (kb + 2) * KB, T, KB);
    }
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
  }

  // ---- normalise, stage through shared memory (Q tile is free), coalesced 16-byte stores ----
  const float inv0 = 1.f / l_run[0], inv1 = 1.f / l_run[1];
#pragma unroll
  for (int i = 0; i < NT; ++i) {
    const int r0 = warp * 16 + g, r1 = r0 + 8;
    const uint32_t w0 = sQ + tile_off<DH>(r0, i) + tq * 4;
    const uint32_t w1 = sQ + tile_off<DH>(r1, i) + tq * 4;
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
      asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(sQ + tile_off<DH>(r, c)));
      *reinterpret_cast<uint4*>(Og + static_cast<int64_t>(q0 + r) * a.ldo + c * 8) = v;
    }
  }
}

template <int DH>
int launch(const fdm_attn_args& a, cudaStream_t stream) {
  const size_t smem = QB * DH * 2 + 4 * KB * DH * 2;
  static bool attr = false;
  if (!attr) {
    FDM_CHECK_CUDA(cudaFuncSetAttribute(attn_mma_kernel<DH>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    attr = true;
  }
  uint32_t magic = 0;
  if (a.bias_mode == 1) magic = 65536u / static_cast<uint32_t>(a.period) + 1u;
  dim3 grid(static_cast<unsigned>(ceil_div64(a.T, QB)), static_cast<unsigned>(a.H), static_cast<unsigned>(a.B));
  attn_mma_kernel<DH><<<grid, THREADS, smem, stream>>>(a, magic);
  FDM_CHECK_LAUNCH();
  return 0;
}

}  // namespace

int fdm_attention_mma_try(const fdm_attn_args& a, cudaStream_t stream, bool* handled) {
  *handled = false;
  if (a.dtype != FDM_BF16 || (a.dh != 64 && a.dh != 128)) return 0;
  // 16-byte cp.async / vector stores need aligned rows; the exact reciprocal needs (T * period) < 65536
  const uintptr_t al = reinterpret_cast<uintptr_t>(a.Q) | reinterpret_cast<uintptr_t>(a.K) | reinterpret_cast<uintptr_t>(a.V) |
                       reinterpret_cast<uintptr_t>(a.O);
  if ((al & 15u) != 0 || a.ldq % 8 || a.ldk % 8 || a.ldv % 8 || a.ldo % 8) return 0;
  if (a.bias_mode == 1 && (a.period > 63 || a.T > 1024)) return 0;
  *handled = true;
  return a.dh == 64 ? launch<64>(a, stream) : launch<128>(a, stream);
}
