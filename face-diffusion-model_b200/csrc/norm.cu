// Row LayerNorm with fused residual adds (one warp per row, values kept in registers, two-pass
// mean/variance in fp32) and LeakyReLU + InstanceNorm over time for the VQ-decoder expander.
#include "common.cuh"
#include <stdlib.h>

namespace {

constexpr int MAX_V4 = 8;  // float4 groups per lane -> d <= 32*4*8 = 1024

__device__ __forceinline__ float4 load4(const void* p, int dtype, int64_t idx) {
  if (dtype == FDM_BF16) {
    const uint2 u = *reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(p) + idx);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
    const float2 a = __bfloat1622float2(h[0]), b = __bfloat1622float2(h[1]);
    return make_float4(a.x, a.y, b.x, b.y);
  }
  return *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p) + idx);
}
__device__ __forceinline__ void store4(void* p, int dtype, int64_t idx, float4 v) {
  if (dtype == FDM_BF16) {
    uint2 u;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
    h[0] = __floats2bfloat162_rn(v.x, v.y);
    h[1] = __floats2bfloat162_rn(v.z, v.w);
    *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p) + idx) = u;
  } else {
    *reinterpret_cast<float4*>(reinterpret_cast<float*>(p) + idx) = v;
  }
}

__device__ __forceinline__ void ln_inplace(float4 (&v)[MAX_V4], int nv, int lane, int d, float eps, const float* g,
                                           const float* b) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < MAX_V4; ++i)
    if (i < nv && (i * 32 + lane) * 4 < d) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  const float mean = warp_sum(s) / static_cast<float>(d);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < MAX_V4; ++i)
    if (i < nv && (i * 32 + lane) * 4 < d) {
      const float a = v[i].x - mean, bb = v[i].y - mean, c = v[i].z - mean, e = v[i].w - mean;
      q += (a * a + bb * bb) + (c * c + e * e);
    }
  const float rstd = 1.f / sqrtf(warp_sum(q) / static_cast<float>(d) + eps);
#pragma unroll
  for (int i = 0; i < MAX_V4; ++i)
    if (i < nv && (i * 32 + lane) * 4 < d) {
      const int c0 = (i * 32 + lane) * 4;
      const float4 gg = *reinterpret_cast<const float4*>(g + c0);
      const float4 bv = *reinterpret_cast<const float4*>(b + c0);
      v[i].x = (v[i].x - mean) * rstd * gg.x + bv.x;
      v[i].y = (v[i].y - mean) * rstd * gg.y + bv.y;
      v[i].z = (v[i].z - mean) * rstd * gg.z + bv.z;
      v[i].w = (v[i].w - mean) * rstd * gg.w + bv.w;
    }
}

__global__ void __launch_bounds__(256) layernorm_kernel(const fdm_norm_args a) {
  const int lane = threadIdx.x & 31;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= a.rows) return;
  const int d = static_cast<int>(a.d);
  const int nv = (d + 127) / 128;
  float4 v[MAX_V4];
#pragma unroll
  for (int i = 0; i < MAX_V4; ++i) {
    const int c0 = (i * 32 + lane) * 4;
    if (i < nv && c0 < d) {
      v[i] = load4(a.x, a.x_dtype, row * a.ldx + c0);
      if (a.r1) {
        const float4 r = load4(a.r1, a.r1_dtype, row * a.ldr1 + c0);
        v[i].x += r.x; v[i].y += r.y; v[i].z += r.z; v[i].w += r.w;
      }
    }
  }
  if (a.g1) ln_inplace(v, nv, lane, d, a.eps, a.g1, a.b1);
  if (a.act1 != FDM_ACT_NONE) {
#pragma unroll
    for (int i = 0; i < MAX_V4; ++i)
      if (i < nv && (i * 32 + lane) * 4 < d) {
        v[i].x = apply_act(v[i].x, a.act1); v[i].y = apply_act(v[i].y, a.act1);
        v[i].z = apply_act(v[i].z, a.act1); v[i].w = apply_act(v[i].w, a.act1);
      }
  }
  if (a.g2) {
    const float* vec = a.vec2 ? a.vec2 + static_cast<int64_t>(*a.vec_index_dev) * d : nullptr;
    const int64_t r2row = a.r2_rows > 0 ? row % a.r2_rows : row;
#pragma unroll
    for (int i = 0; i < MAX_V4; ++i) {
      const int c0 = (i * 32 + lane) * 4;
      if (i < nv && c0 < d) {
        if (a.r2) {
          const float4 r = load4(a.r2, a.r2_dtype, r2row * a.ldr2 + c0);
          v[i].x += r.x; v[i].y += r.y; v[i].z += r.z; v[i].w += r.w;
        }
        if (vec) {
          const float4 r = *reinterpret_cast<const float4*>(vec + c0);
          v[i].x += r.x; v[i].y += r.y; v[i].z += r.z; v[i].w += r.w;
        }
      }
    }
    ln_inplace(v, nv, lane, d, a.eps, a.g2, a.b2);
  }
#pragma unroll
  for (int i = 0; i < MAX_V4; ++i) {
    const int c0 = (i * 32 + lane) * 4;
    if (i < nv && c0 < d) {
      store4(a.out, a.out_dtype, row * a.ldo + c0, v[i]);
      if (a.out2) store4(a.out2, a.out2_dtype, row * a.ldo2 + c0, v[i]);
    }
  }
}

// ---- fast path: d in {512, 1024}, one dtype for x / r1 / r2 / out, no second output -----------------------------
// Two rows per warp (gamma/beta loads amortised over both), 16-byte loads/stores, compile-time trip counts and
// dtype, 32-bit indexing inside the row. The generic kernel above needed ~1600 warp-instructions per row and was
// issue-bound at ~25% of HBM speed; this one needs ~350.
template <typename T> struct Chunk;
template <> struct Chunk<float> {
  static constexpr int CE = 4;
  static __device__ __forceinline__ void load(const float* p, float (&o)[4]) {
    const float4 t = *reinterpret_cast<const float4*>(p);
    o[0] = t.x; o[1] = t.y; o[2] = t.z; o[3] = t.w;
  }
  static __device__ __forceinline__ void store(float* p, const float (&o)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(o[0], o[1], o[2], o[3]);
  }
};
template <> struct Chunk<__nv_bfloat16> {
  static constexpr int CE = 8;
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&o)[8]) {
    const uint4 u = *reinterpret_cast<const uint4*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float2 f = __bfloat1622float2(h[q]);
      o[2 * q] = f.x; o[2 * q + 1] = f.y;
    }
  }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&o)[8]) {
    uint4 u;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
    for (int q = 0; q < 4; ++q) h[q] = __floats2bfloat162_rn(o[2 * q], o[2 * q + 1]);
    *reinterpret_cast<uint4*>(p) = u;
  }
};

template <int EPL, int CE>
__device__ __forceinline__ void ln2rows(float (&v)[2][EPL], int lane, float eps, const float* __restrict__ g,
                                        const float* __restrict__ b) {
  constexpr int NCH = EPL / CE;
  constexpr float inv_d = 1.f / (EPL * 32);
  float s0 = 0.f, s1 = 0.f;
#pragma unroll
  for (int e = 0; e < EPL; ++e) { s0 += v[0][e]; s1 += v[1][e]; }
  const float m0 = warp_sum(s0) * inv_d, m1 = warp_sum(s1) * inv_d;
  float q0 = 0.f, q1 = 0.f;
#pragma unroll
  for (int e = 0; e < EPL; ++e) {
    v[0][e] -= m0; v[1][e] -= m1;
    q0 = fmaf(v[0][e], v[0][e], q0); q1 = fmaf(v[1][e], v[1][e], q1);
  }
  const float r0 = 1.f / sqrtf(warp_sum(q0) * inv_d + eps), r1 = 1.f / sqrtf(warp_sum(q1) * inv_d + eps);
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    const int c0 = (lane + 32 * i) * CE;
#pragma unroll
    for (int k = 0; k < CE / 4; ++k) {
      const float4 gg = __ldg(reinterpret_cast<const float4*>(g + c0) + k);
      const float4 bb = __ldg(reinterpret_cast<const float4*>(b + c0) + k);
      const float gv[4] = {gg.x, gg.y, gg.z, gg.w}, bv[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int idx = i * CE + k * 4 + e;
        v[0][idx] = fmaf(v[0][idx] * r0, gv[e], bv[e]);
        v[1][idx] = fmaf(v[1][idx] * r1, gv[e], bv[e]);
      }
    }
  }
}

// MODE 0: every option decided at run time. MODE 1: plain LayerNorm (g1 only). MODE 2: the denoiser's fused pair
// LN1 -> + cross-attention cache (r2) + time row (vec2) -> LN2. The specialisations drop the uniform branches and the
// code behind them (same lesson as the GEMM epilogue: they cost registers and issue slots even when not taken).
template <typename T, int EPL, int MODE>
__global__ void __launch_bounds__(128) layernorm_fast_kernel(const fdm_norm_args a) {
  constexpr int CE = Chunk<T>::CE;
  constexpr int NCH = EPL / CE;
  const int lane = threadIdx.x & 31;
  const int64_t row0 = (static_cast<int64_t>(blockIdx.x) * 4 + (threadIdx.x >> 5)) * 2;
  pdl_trigger();
  pdl_wait();
  if (row0 >= a.rows) return;
  const bool two = row0 + 1 < a.rows;
  const int64_t rows[2] = {row0, two ? row0 + 1 : row0};
  float v[2][EPL];
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const T* xp = reinterpret_cast<const T*>(a.x) + rows[r] * a.ldx;
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      float t[CE];
      Chunk<T>::load(xp + (lane + 32 * i) * CE, t);
#pragma unroll
      for (int e = 0; e < CE; ++e) v[r][i * CE + e] = t[e];
    }
    if (MODE == 0 && a.r1) {
      const T* rp = reinterpret_cast<const T*>(a.r1) + rows[r] * a.ldr1;
#pragma unroll
      for (int i = 0; i < NCH; ++i) {
        float t[CE];
        Chunk<T>::load(rp + (lane + 32 * i) * CE, t);
#pragma unroll
        for (int e = 0; e < CE; ++e) v[r][i * CE + e] += t[e];
      }
    }
  }
  if (MODE != 0 || a.g1) ln2rows<EPL, CE>(v, lane, a.eps, a.g1, a.b1);
  if (MODE == 0 && a.act1 == FDM_ACT_GELU_ERF) {
#pragma unroll
    for (int e = 0; e < EPL; ++e) { v[0][e] = act_gelu_erf(v[0][e]); v[1][e] = act_gelu_erf(v[1][e]); }
  }
  if (MODE == 2 || (MODE == 0 && a.g2)) {
    if (MODE == 2 || a.r2) {
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int64_t rr = a.r2_rows > 0 ? rows[r] % a.r2_rows : rows[r];
        const T* rp = reinterpret_cast<const T*>(a.r2) + rr * a.ldr2;
#pragma unroll
        for (int i = 0; i < NCH; ++i) {
          float t[CE];
          Chunk<T>::load(rp + (lane + 32 * i) * CE, t);
#pragma unroll
          for (int e = 0; e < CE; ++e) v[r][i * CE + e] += t[e];
        }
      }
    }
    if (MODE == 2 || a.vec2) {
      const float* vec = a.vec2 + static_cast<int64_t>(*a.vec_index_dev) * (EPL * 32);
#pragma unroll
      for (int i = 0; i < NCH; ++i)
#pragma unroll
        for (int k = 0; k < CE / 4; ++k) {
          const float4 t = __ldg(reinterpret_cast<const float4*>(vec + (lane + 32 * i) * CE) + k);
          const int idx = i * CE + k * 4;
          v[0][idx] += t.x; v[0][idx + 1] += t.y; v[0][idx + 2] += t.z; v[0][idx + 3] += t.w;
          v[1][idx] += t.x; v[1][idx + 1] += t.y; v[1][idx + 2] += t.z; v[1][idx + 3] += t.w;
        }
    }
    ln2rows<EPL, CE>(v, lane, a.eps, a.g2, a.b2);
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    if (r == 1 && !two) break;
    T* op = reinterpret_cast<T*>(a.out) + rows[r] * a.ldo;
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      float t[CE];
#pragma unroll
      for (int e = 0; e < CE; ++e) t[e] = v[r][i * CE + e];
      Chunk<T>::store(op + (lane + 32 * i) * CE, t);
    }
  }
}

template <typename T, int EPL>
bool launch_fast_ln(const fdm_norm_args& a, cudaStream_t s) {
  const unsigned grid = static_cast<unsigned>(ceil_div64(a.rows, 8));
  const bool plain = a.g1 && !a.r1 && a.act1 == FDM_ACT_NONE && !a.g2;
  const bool pair = a.g1 && !a.r1 && a.act1 == FDM_ACT_NONE && a.g2 && a.r2 && a.vec2;
  if (plain) return fdm_launch_pdl(layernorm_fast_kernel<T, EPL, 1>, dim3(grid), dim3(128), 0, s, 1, a) == cudaSuccess;
  if (pair) return fdm_launch_pdl(layernorm_fast_kernel<T, EPL, 2>, dim3(grid), dim3(128), 0, s, 1, a) == cudaSuccess;
  return fdm_launch_pdl(layernorm_fast_kernel<T, EPL, 0>, dim3(grid), dim3(128), 0, s, 1, a) == cudaSuccess;
}
template <typename T>
bool try_fast_ln(const fdm_norm_args& a, cudaStream_t s) {
  if (a.d == 1024) return launch_fast_ln<T, 32>(a, s);
  if (a.d == 512) return launch_fast_ln<T, 16>(a, s);
  return false;
}

// ---- hot-loop LayerNorm kernels of the denoiser step (bf16, d in {512, 1024}) ------------------------------------------
// The warp-per-two-rows kernel above ran at 2.9-4.0 TB/s: 125 registers (16 warps per SM), ~720 warp instructions per row of
// the fused pair (gamma / beta re-read from L1 for every row pair, the cross-attention rows fetched only after the first
// LayerNorm), long-scoreboard bound (profiles/r01_i_layernorm_full.md). Two kernels replace it:
//   * plain LayerNorm (norm3): column-owner layout. A thread owns 8 consecutive columns (one 16-byte chunk) of R rows, d / 8
//     threads cover a row; gamma / beta of those columns live in registers, every load of the CTA is in flight before the
//     first use, row statistics are a reduce-scatter over the R rows a lane holds plus one shared-memory exchange between the
//     d / 256 warps of a row.
//   * fused pair (norm1 -> + cross-attention cache + time row -> norm2): one warp per row, x and the cache row loaded up
//     front, gamma / beta / (beta1 + time row) staged once per CTA in shared memory, ~75 registers (24+ warps per SM), no
//     block-wide synchronisation in the row loop.
// Row statistics are (mean, M2) pairs merged with Chan's equal-count formula: two-pass accuracy in one reduction.
__device__ __forceinline__ void chan_merge(float& m, float& q, float rm, float rq, float n_half) {
  const float dl = rm - m;
  m = fmaf(0.5f, dl, m);
  q = fmaf(dl * dl, n_half, q + rq);
}
template <int R>
__device__ __forceinline__ void stat_reduce_scatter(float (&m)[R], float (&q)[R], int lane) {
  // in: per-lane (mean, M2) of 8 elements of each of R rows. out: m[0], q[0] = the statistics of 256 elements of row
  // `stat_row<R>(lane)`, identical in the 32 / R lanes that share that row.
  float n_half = 4.f;  // half the element count of each side of a merge
  int half = R / 2;
#pragma unroll
  for (int mask = 16; mask >= 1; mask >>= 1) {
    if (half >= 1) {
      const bool up = (lane & mask) != 0;
#pragma unroll
      for (int i = 0; i < R / 2; ++i) {
        if (i < half) {
          const float sm = up ? m[i] : m[i + half], sq = up ? q[i] : q[i + half];
          float km = up ? m[i + half] : m[i], kq = up ? q[i + half] : q[i];
          chan_merge(km, kq, __shfl_xor_sync(0xffffffffu, sm, mask), __shfl_xor_sync(0xffffffffu, sq, mask), n_half);
          m[i] = km;
          q[i] = kq;
        }
      }
      half >>= 1;
    } else {
      chan_merge(m[0], q[0], __shfl_xor_sync(0xffffffffu, m[0], mask), __shfl_xor_sync(0xffffffffu, q[0], mask), n_half);
    }
    n_half *= 2.f;
  }
}
template <int R>
__device__ __forceinline__ int stat_row(int lane) {  // which of the R rows a lane holds after stat_reduce_scatter
  return R == 8 ? (lane >> 2) : (R == 4 ? (lane >> 3) : (R == 2 ? (lane >> 4) : 0));
}

__device__ __forceinline__ void unpack8(const uint4& u, float* o) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    o[2 * k] = __uint_as_float(w[k] << 16);
    o[2 * k + 1] = __uint_as_float(w[k] & 0xffff0000u);
  }
}
__device__ __forceinline__ uint4 pack8(const float* v) {
  uint4 u;
  __nv_bfloat162 h;
  h = __floats2bfloat162_rn(v[0], v[1]); u.x = *reinterpret_cast<uint32_t*>(&h);
  h = __floats2bfloat162_rn(v[2], v[3]); u.y = *reinterpret_cast<uint32_t*>(&h);
  h = __floats2bfloat162_rn(v[4], v[5]); u.z = *reinterpret_cast<uint32_t*>(&h);
  h = __floats2bfloat162_rn(v[6], v[7]); u.w = *reinterpret_cast<uint32_t*>(&h);
  return u;
}
template <int N>
__device__ __forceinline__ void local_stat(const float (&v)[N], float& m, float& q) {
  float s = 0.f;
#pragma unroll
  for (int e = 0; e < N; ++e) s += v[e];
  m = s * (1.f / N);
  q = 0.f;
#pragma unroll
  for (int e = 0; e < N; ++e) {
    const float c = v[e] - m;
    q = fmaf(c, c, q);
  }
}
__device__ __forceinline__ void load8f(const float* p, float (&o)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w; o[4] = b.x; o[5] = b.y; o[6] = b.z; o[7] = b.w;
}

template <int D, int R, bool RES>
__global__ void __launch_bounds__(128) layernorm_cols_kernel(const fdm_norm_args a) {
  constexpr int TPR = D / 8;     // threads per row
  constexpr int WPR = TPR / 32;  // warps per row: 4 (d = 1024) or 2 (d = 512)
  constexpr int RPP = 128 / TPR; // rows per pass of the CTA
  __shared__ float2 st[R][RPP][WPR];
  const int tid = threadIdx.x, lane = tid & 31;
  const int sub = tid / TPR, ct = tid % TPR, wr = ct >> 5;
  const int col0 = ct * 8;
  const int64_t row_base = static_cast<int64_t>(blockIdx.x) * (R * RPP) + sub;
  const int64_t last = a.rows - 1;
  pdl_trigger();
  float g1[8], b1[8];  // weights: independent of the previous kernel, so these loads overlap its tail
  load8f(a.g1 + col0, g1);
  load8f(a.b1 + col0, b1);
  pdl_wait();
  const __nv_bfloat16* xp = reinterpret_cast<const __nv_bfloat16*>(a.x);
  uint4 xr[R], rr[RES ? R : 1];
#pragma unroll
  for (int i = 0; i < R; ++i) xr[i] = *reinterpret_cast<const uint4*>(xp + min(row_base + i * RPP, last) * a.ldx + col0);
  if (RES) {  // residual rows (x + f(x) formed here in fp32: the sum is never rounded to bf16)
    const __nv_bfloat16* r1p = reinterpret_cast<const __nv_bfloat16*>(a.r1);
#pragma unroll
    for (int i = 0; i < R; ++i) rr[i] = *reinterpret_cast<const uint4*>(r1p + min(row_base + i * RPP, last) * a.ldr1 + col0);
  }
  float m[R], q[R], v[R][8];
#pragma unroll
  for (int i = 0; i < R; ++i) {
    unpack8(xr[i], v[i]);
    if (RES) {
      float t[8];
      unpack8(rr[i], t);
#pragma unroll
      for (int e = 0; e < 8; ++e) v[i][e] += t[e];
    }
    local_stat<8>(v[i], m[i], q[i]);
  }
  stat_reduce_scatter<R>(m, q, lane);
  if ((lane & (32 / R - 1)) == 0) st[stat_row<R>(lane)][sub][wr] = make_float2(m[0], q[0]);
  __syncthreads();
#pragma unroll
  for (int i = 0; i < R; ++i) {
    float mean = 0.f, M2 = 0.f, pm[WPR];
#pragma unroll
    for (int w = 0; w < WPR; ++w) {
      const float2 p = st[i][sub][w];
      pm[w] = p.x;
      mean += p.x;
      M2 += p.y;
    }
    mean *= 1.f / WPR;
#pragma unroll
    for (int w = 0; w < WPR; ++w) M2 = fmaf((pm[w] - mean) * (pm[w] - mean), 256.f, M2);
    const float rstd = 1.f / sqrtf(M2 * (1.f / D) + a.eps), nmr = -mean * rstd;
#pragma unroll
    for (int e = 0; e < 8; ++e) v[i][e] = fmaf(fmaf(v[i][e], rstd, nmr), g1[e], b1[e]);
    const int64_t row = row_base + i * RPP;
    if (row <= last) *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(a.out) + row * a.ldo + col0) = pack8(v[i]);
  }
}

// one warp per row. PAIR: out = LN(LN(x [+ r1]; g1, b1) + r2[row % r2_rows] + vec2[*vec_index_dev]; g2, b2); else: out = LN(x [+ r1]; g1, b1)
template <int D, bool RES, bool PAIR>
__global__ void __launch_bounds__(256) layernorm_rows_kernel(const fdm_norm_args a, const int rows_per_cta) {
  constexpr int NCH = D / 256;  // 16-byte chunks per lane: chunk k covers columns (32 k + lane) * 8 ...
  constexpr int EPL = NCH * 8;
  // gamma1, beta1 (+ time row), gamma2, beta2; float4 i of an array (columns 4 i ...) lives at psw(i): the two float4 a lane
  // needs per chunk are 32 float4 apart, so a quarter-warp's 16-byte reads are conflict-free
  __shared__ float4 sp[PAIR ? 4 : 2][D / 4];
  auto psw = [](int i) { return ((i >> 6) * 2 + (i & 1)) * 32 + ((i >> 1) & 31); };
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  pdl_trigger();
  for (int i = tid; i < D / 4; i += 256) {
    sp[0][psw(i)] = __ldg(reinterpret_cast<const float4*>(a.g1) + i);
    if (PAIR) {
      sp[2][psw(i)] = __ldg(reinterpret_cast<const float4*>(a.g2) + i);
      sp[3][psw(i)] = __ldg(reinterpret_cast<const float4*>(a.b2) + i);
    } else {
      sp[1][psw(i)] = __ldg(reinterpret_cast<const float4*>(a.b1) + i);
    }
  }
  pdl_wait();
  if (PAIR) {
    const float4* vec = reinterpret_cast<const float4*>(a.vec2 + static_cast<int64_t>(*a.vec_index_dev) * D);
    for (int i = tid; i < D / 4; i += 256) {
      const float4 b = __ldg(reinterpret_cast<const float4*>(a.b1) + i), t = __ldg(vec + i);
      sp[1][psw(i)] = make_float4(b.x + t.x, b.y + t.y, b.z + t.z, b.w + t.w);
    }
  }
  __syncthreads();
  const int64_t row_end = min(a.rows, (static_cast<int64_t>(blockIdx.x) + 1) * rows_per_cta);
  const __nv_bfloat16* xp = reinterpret_cast<const __nv_bfloat16*>(a.x);
  const __nv_bfloat16* rp = reinterpret_cast<const __nv_bfloat16*>(a.r2);
  for (int64_t row = static_cast<int64_t>(blockIdx.x) * rows_per_cta + warp; row < row_end; row += 8) {
    const int64_t rr = !PAIR ? 0 : (a.r2_rows > 0 ? (row < a.r2_rows ? row : (row < 2 * a.r2_rows ? row - a.r2_rows : row % a.r2_rows)) : row);
    uint4 xr[NCH], cr[PAIR ? NCH : 1], r1r[RES ? NCH : 1];
#pragma unroll
    for (int k = 0; k < NCH; ++k) xr[k] = *reinterpret_cast<const uint4*>(xp + row * a.ldx + (32 * k + lane) * 8);
    if (RES) {
      const __nv_bfloat16* r1p = reinterpret_cast<const __nv_bfloat16*>(a.r1) + row * a.ldr1;
#pragma unroll
      for (int k = 0; k < NCH; ++k) r1r[k] = *reinterpret_cast<const uint4*>(r1p + (32 * k + lane) * 8);
    }
    if (PAIR) {
#pragma unroll
      for (int k = 0; k < NCH; ++k) cr[k] = *reinterpret_cast<const uint4*>(rp + rr * a.ldr2 + (32 * k + lane) * 8);
    }
    float v[EPL];
#pragma unroll
    for (int k = 0; k < NCH; ++k) unpack8(xr[k], v + 8 * k);
    if (RES) {  // x + f(x) formed in fp32 (the sum is never rounded to bf16)
#pragma unroll
      for (int k = 0; k < NCH; ++k) {
        float t[8];
        unpack8(r1r[k], t);
#pragma unroll
        for (int e = 0; e < 8; ++e) v[8 * k + e] += t[e];
      }
    }
#pragma unroll
    for (int pass = 0; pass < (PAIR ? 2 : 1); ++pass) {
      float m, q;
      local_stat<EPL>(v, m, q);
      float n_half = 0.5f * EPL;
#pragma unroll
      for (int mask = 16; mask >= 1; mask >>= 1) {
        chan_merge(m, q, __shfl_xor_sync(0xffffffffu, m, mask), __shfl_xor_sync(0xffffffffu, q, mask), n_half);
        n_half *= 2.f;
      }
      const float rstd = 1.f / sqrtf(q * (1.f / D) + a.eps), nmr = -m * rstd;
#pragma unroll
      for (int k = 0; k < NCH; ++k) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const float4 g = sp[2 * pass][(2 * k + h) * 32 + lane], b = sp[2 * pass + 1][(2 * k + h) * 32 + lane];
          float* o = v + 8 * k + 4 * h;
          o[0] = fmaf(fmaf(o[0], rstd, nmr), g.x, b.x);
          o[1] = fmaf(fmaf(o[1], rstd, nmr), g.y, b.y);
          o[2] = fmaf(fmaf(o[2], rstd, nmr), g.z, b.z);
          o[3] = fmaf(fmaf(o[3], rstd, nmr), g.w, b.w);
        }
      }
      if (PAIR && pass == 0) {
#pragma unroll
        for (int k = 0; k < NCH; ++k) {
          float c[8];
          unpack8(cr[k], c);
#pragma unroll
          for (int e = 0; e < 8; ++e) v[8 * k + e] += c[e];
        }
      }
    }
    __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(a.out) + row * a.ldo;
#pragma unroll
    for (int k = 0; k < NCH; ++k) *reinterpret_cast<uint4*>(op + (32 * k + lane) * 8) = pack8(v + 8 * k);
  }
}

template <int D, bool RES>
bool launch_hot_ln(const fdm_norm_args& a, bool pair, cudaStream_t s) {
  static const int plain_rows = [] { const char* e = getenv("FDM_B200_LN_PLAIN_ROWS"); return e ? atoi(e) : 1; }();
  if (pair || plain_rows) {
    // ~3 CTAs of 8 warps per SM, each staging the parameters (16 KB for the pair at d = 1024) once for its share of the rows
    const int64_t ctas = (pair ? 3 : 4) * static_cast<int64_t>(fdm_sm_count());
    int64_t per = ceil_div64(a.rows, ctas);
    per = (per + 7) / 8 * 8;
    const unsigned grid = static_cast<unsigned>(ceil_div64(a.rows, per));
    if (pair) return fdm_launch_pdl(layernorm_rows_kernel<D, RES, true>, dim3(grid), dim3(256), 0, s, 1, a, static_cast<int>(per)) == cudaSuccess;
    return fdm_launch_pdl(layernorm_rows_kernel<D, RES, false>, dim3(grid), dim3(256), 0, s, 1, a, static_cast<int>(per)) == cudaSuccess;
  }
  constexpr int R = 4;
  const unsigned grid = static_cast<unsigned>(ceil_div64(a.rows, R * (128 * 8 / D)));
  return fdm_launch_pdl(layernorm_cols_kernel<D, R, RES>, dim3(grid), dim3(128), 0, s, 1, a) == cudaSuccess;
}
// bf16 plain / fused-pair LayerNorm of the denoiser step. FDM_B200_LN_HOT: 0 = warp-per-two-rows kernel for both,
// 1 = new plain kernel only, 2 = both (default)
bool try_hot_ln(const fdm_norm_args& a, cudaStream_t s) {
  static const int mode = [] { const char* e = getenv("FDM_B200_LN_HOT"); return e ? atoi(e) : 2; }();
  if (mode <= 0 || a.x_dtype != FDM_BF16 || a.act1 != FDM_ACT_NONE || !a.g1) return false;
  const bool plain = !a.g2, pair = a.g2 && a.r2 && a.vec2;
  if (!(plain || (pair && mode >= 2))) return false;
  const bool res = a.r1 != nullptr;  // (same dtype / alignment as x: checked by the caller)
  if (a.d == 1024) return res ? launch_hot_ln<1024, true>(a, pair, s) : launch_hot_ln<1024, false>(a, pair, s);
  if (a.d == 512) return res ? launch_hot_ln<512, true>(a, pair, s) : launch_hot_ln<512, false>(a, pair, s);
  return false;
}

// block = 32 channels x 8 time-lanes; grid = (C/32, B)
__global__ void __launch_bounds__(256) leaky_instnorm_kernel(const void* x, int x_dtype, void* out, int out_dtype, int T,
                                                             int64_t t_stride, int64_t out_t_stride, int C, float slope,
                                                             float eps, const float* gamma, const float* beta, int post_act) {
  __shared__ float red[8][33];
  const int cx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cx;
  const int64_t base = static_cast<int64_t>(blockIdx.y) * t_stride * C;
  const int64_t obase = static_cast<int64_t>(blockIdx.y) * out_t_stride * C;
  const bool ok = c < C;
  float s = 0.f;
  for (int t = ty; t < T; t += 8)
    if (ok) {
      float v = ld_as_float(x, x_dtype, base + static_cast<int64_t>(t) * C + c);
      s += v > 0.f ? v : slope * v;
    }
  red[ty][cx] = s;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) tot += red[i][cx];
  const float mean = tot / static_cast<float>(T);
  __syncthreads();
  float q = 0.f;
  for (int t = ty; t < T; t += 8)
    if (ok) {
      float v = ld_as_float(x, x_dtype, base + static_cast<int64_t>(t) * C + c);
      v = (v > 0.f ? v : slope * v) - mean;
      q += v * v;
    }
  red[ty][cx] = q;
  __syncthreads();
  tot = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) tot += red[i][cx];
  const float rstd = 1.f / sqrtf(tot / static_cast<float>(T) + eps);
  const float gm = (gamma && ok) ? gamma[c] : 1.f, bt = (beta && ok) ? beta[c] : 0.f;
  for (int t = ty; t < T; t += 8)
    if (ok) {
      float v = ld_as_float(x, x_dtype, base + static_cast<int64_t>(t) * C + c);
      v = (v > 0.f ? v : slope * v);
      v = (v - mean) * rstd * gm + bt;
      if (post_act == FDM_ACT_GELU_ERF) v = act_gelu_erf(v);
      st_from_float(out, out_dtype, obase + static_cast<int64_t>(t) * C + c, v);
    }
}

inline bool ok_align(const void* p, int dtype, int64_t ld) {
  const uintptr_t a = dtype == FDM_BF16 ? 8 : 16;
  return p == nullptr || ((reinterpret_cast<uintptr_t>(p) % a) == 0 && ld % 4 == 0);
}

}  // namespace

extern "C" int fdm_layernorm(const fdm_norm_args* args, void* stream) {
  FDM_CHECK_ARG(args != nullptr, "fdm_layernorm: null args");
  const fdm_norm_args& a = *args;
  FDM_CHECK_ARG(a.x && a.out, "fdm_layernorm: null x/out");
  FDM_CHECK_ARG(a.rows >= 0 && a.d > 0 && a.d % 4 == 0 && a.d <= 128 * MAX_V4, "fdm_layernorm: d=%lld must be a multiple of 4 and <= %d",
                (long long)a.d, 128 * MAX_V4);
  FDM_CHECK_ARG(ok_align(a.x, a.x_dtype, a.ldx) && ok_align(a.r1, a.r1_dtype, a.ldr1) && ok_align(a.r2, a.r2_dtype, a.ldr2) &&
                    ok_align(a.out, a.out_dtype, a.ldo) && ok_align(a.out2, a.out2_dtype, a.ldo2),
                "fdm_layernorm: operands must be vector-aligned with strides %% 4 == 0");
  FDM_CHECK_ARG(!a.vec2 || a.vec_index_dev, "fdm_layernorm: vec2 needs vec_index_dev");
  FDM_CHECK_ARG((!a.g1 || a.b1) && (!a.g2 || a.b2), "fdm_layernorm: gamma without beta");
  if (a.rows == 0) return 0;
  {
    const bool same = (!a.r1 || a.r1_dtype == a.x_dtype) && (!a.r2 || a.r2_dtype == a.x_dtype) && a.out_dtype == a.x_dtype;
    auto al16 = [](const void* p, int64_t ld, int esz) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) % 16 == 0 && (ld * esz) % 16 == 0); };
    const int esz = a.x_dtype == FDM_BF16 ? 2 : 4;
    const bool aligned = al16(a.x, a.ldx, esz) && al16(a.r1, a.ldr1, esz) && al16(a.r2, a.ldr2, esz) && al16(a.out, a.ldo, esz) &&
                         al16(a.g1, 0, 4) && al16(a.b1, 0, 4) && al16(a.g2, 0, 4) && al16(a.b2, 0, 4) && al16(a.vec2, 0, 4);
    if (same && aligned && !a.out2 && (a.act1 == FDM_ACT_NONE || a.act1 == FDM_ACT_GELU_ERF) && (a.d == 512 || a.d == 1024)) {
      cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
      if (try_hot_ln(a, st)) {
        FDM_CHECK_LAUNCH();
        return 0;
      }
      const bool ok = a.x_dtype == FDM_BF16 ? try_fast_ln<__nv_bfloat16>(a, st) : try_fast_ln<float>(a, st);
      if (ok) {
        FDM_CHECK_LAUNCH();
        return 0;
      }
    }
  }
  const int warps = 8;
  layernorm_kernel<<<static_cast<unsigned>(ceil_div64(a.rows, warps)), warps * 32, 0, reinterpret_cast<cudaStream_t>(stream)>>>(a);
  FDM_CHECK_LAUNCH();
  return 0;
}

extern "C" int fdm_leaky_instnorm(const void* x, int32_t x_dtype, void* out, int32_t out_dtype, int64_t B, int64_t T,
                                  int64_t t_stride, int64_t out_t_stride, int64_t C, float slope, float eps, const float* gamma,
                                  const float* beta, int32_t post_act, void* stream) {
  FDM_CHECK_ARG(x && out && B > 0 && T > 0 && C > 0 && t_stride >= T && out_t_stride >= T, "fdm_leaky_instnorm: bad arguments");
  FDM_CHECK_ARG(x != out || t_stride == out_t_stride, "fdm_leaky_instnorm: in-place needs equal strides");
  FDM_CHECK_ARG(post_act == FDM_ACT_NONE || post_act == FDM_ACT_GELU_ERF, "fdm_leaky_instnorm: post_act must be NONE or GELU_ERF");
  dim3 grid(static_cast<unsigned>(ceil_div64(C, 32)), static_cast<unsigned>(B));
  leaky_instnorm_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, x_dtype, out, out_dtype, static_cast<int>(T),
                                                                                  t_stride, out_t_stride, static_cast<int>(C), slope, eps, gamma, beta, post_act);
  FDM_CHECK_LAUNCH();
  return 0;
}

namespace {
// one thread per row: (mean, rstd) from the per-column-group partial sums a GEMM epilogue wrote (fixed summation order)
__global__ void __launch_bounds__(256) ln_stats_finalize_kernel(const float* __restrict__ partials, int64_t M, int parts, float inv_d,
                                                                float eps, float* __restrict__ mean_rstd) {
  pdl_trigger();
  pdl_wait();
  const int64_t row = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (row >= M) return;
  const float2* p = reinterpret_cast<const float2*>(partials) + row * parts;
  float s = 0.f, q = 0.f;
  for (int i = 0; i < parts; ++i) {
    const float2 v = p[i];
    s += v.x;
    q += v.y;
  }
  const float mean = s * inv_d;
  const float var = fmaxf(q * inv_d - mean * mean, 0.f);
  reinterpret_cast<float2*>(mean_rstd)[row] = make_float2(mean, 1.f / sqrtf(var + eps));
}
}  // namespace

extern "C" int fdm_ln_stats_finalize(const float* partials, int64_t M, int64_t parts, int64_t d, float eps, float* mean_rstd,
                                     void* stream) {
  FDM_CHECK_ARG(partials && mean_rstd && M > 0 && parts > 0 && d > 0, "fdm_ln_stats_finalize: bad arguments");
  const unsigned grid = static_cast<unsigned>(ceil_div64(M, 256));
  FDM_CHECK_CUDA(fdm_launch_pdl(ln_stats_finalize_kernel, dim3(grid), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), 1, partials, M,
                                static_cast<int>(parts), 1.f / static_cast<float>(d), eps, mean_rstd));
  return 0;
}
