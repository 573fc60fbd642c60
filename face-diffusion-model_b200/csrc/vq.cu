// fdm_vq_quantize: EVQ-VAE nearest-code search with a DEFINED fp32 expression so that indices are bit-exact
// against the CPU oracle (oracle/vq_ref.c):
//     d_j = (zz + ee_j) - 2 * dot_j,   zz = sum_k z_k^2, ee_j = sum_k e_jk^2, dot_j = sum_k z_k e_jk,
// every sum a sequential fmaf chain over k = 0..D-1 starting at +0.0f; argmin with the lowest index on ties.
// Persistent CTAs keep the (per-clip) codebook slice resident in shared memory; each CTA tile is 64 latent
// rows x all codes, register-tiled 8 rows x 8 codes per thread, then a warp-shuffle (d, index) min-reduction.
// The gather of the winning code and the (B, L, D) -> (B, D, L) permute of the reference are fused into the
// store phase.
#include "common.cuh"

namespace {

constexpr int ROWS = 64;      // latent rows per tile
constexpr int THREADS = 256;  // 8 warps; warp w owns rows 8w..8w+7, lane owns codes lane + 32*j
constexpr int MAXJ = 8;       // up to 256 codes

template <int D>
__global__ void __launch_bounds__(THREADS) vq_kernel(const float* __restrict__ z, const float* __restrict__ codebook,
                                                     const int64_t* __restrict__ code_offset, int64_t B, int64_t L, int n_codes,
                                                     int64_t* __restrict__ indices, float* __restrict__ zq_bdl,
                                                     float* __restrict__ zq_rows) {
  constexpr int DP = D + 4;  // padded row: conflict-free float4 reads across lanes
  extern __shared__ float sm[];
  float* cb = sm;                   // [n_codes][DP]
  float* zt = cb + 256 * DP;        // [ROWS][DP]
  float* ee = zt + ROWS * DP;       // [256]
  float* zz = ee + 256;             // [ROWS]
  int* widx = reinterpret_cast<int*>(zz + ROWS);  // [ROWS]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ncj = n_codes / 32;
  const int64_t tiles_per_clip = (L + ROWS - 1) / ROWS;
  const int64_t num_tiles = B * tiles_per_clip;
  int64_t loaded_off = -1;

  for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
    const int64_t b = tile / tiles_per_clip;
    const int64_t l0 = (tile - b * tiles_per_clip) * ROWS;
    const int nrows = static_cast<int>(min(static_cast<int64_t>(ROWS), L - l0));
    const int64_t off = code_offset ? code_offset[b] : 0;
    __syncthreads();  // previous tile's store phase is done with cb / widx
    if (off != loaded_off) {
      for (int i = tid; i < n_codes * (D / 4); i += THREADS) {
        const int c = i / (D / 4), k4 = i - c * (D / 4);
        *reinterpret_cast<float4*>(cb + c * DP + k4 * 4) =
            *reinterpret_cast<const float4*>(codebook + (off + c) * D + k4 * 4);
      }
      loaded_off = off;
      __syncthreads();
      for (int c = tid; c < n_codes; c += THREADS) {
        float s = 0.f;
        for (int k = 0; k < D; ++k) s = fmaf(cb[c * DP + k], cb[c * DP + k], s);
        ee[c] = s;
      }
    }
    const float* zsrc = z + (b * L + l0) * D;
    for (int i = tid; i < ROWS * (D / 4); i += THREADS) {
      const int r = i / (D / 4), k4 = i - r * (D / 4);
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < nrows) v = *reinterpret_cast<const float4*>(zsrc + static_cast<int64_t>(r) * D + k4 * 4);
      *reinterpret_cast<float4*>(zt + r * DP + k4 * 4) = v;
    }
    __syncthreads();
    if (tid < ROWS) {
      float s = 0.f;
      for (int k = 0; k < D; ++k) s = fmaf(zt[tid * DP + k], zt[tid * DP + k], s);
      zz[tid] = s;
    }

    float dot[8][MAXJ];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < MAXJ; ++j) dot[i][j] = 0.f;
#pragma unroll 2
    for (int k4 = 0; k4 < D / 4; ++k4) {
      float4 zv[8], ev[MAXJ];
#pragma unroll
      for (int i = 0; i < 8; ++i) zv[i] = *reinterpret_cast<const float4*>(zt + (warp * 8 + i) * DP + k4 * 4);
#pragma unroll
      for (int j = 0; j < MAXJ; ++j)
        ev[j] = j < ncj ? *reinterpret_cast<const float4*>(cb + (lane + 32 * j) * DP + k4 * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < MAXJ; ++j) {
          float a = dot[i][j];
          a = fmaf(zv[i].x, ev[j].x, a);
          a = fmaf(zv[i].y, ev[j].y, a);
          a = fmaf(zv[i].z, ev[j].z, a);
          a = fmaf(zv[i].w, ev[j].w, a);
          dot[i][j] = a;
        }
    }
    __syncthreads();  // zz ready (and every warp done reading zt)
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = warp * 8 + i;
      const float zzr = zz[r];
      float best = INFINITY;
      int bidx = 0x7fffffff;
#pragma unroll
      for (int j = 0; j < MAXJ; ++j)
        if (j < ncj) {
          const int c = lane + 32 * j;
          const float dist = __fsub_rn(__fadd_rn(zzr, ee[c]), __fmul_rn(2.f, dot[i][j]));
          if (dist < best || (dist == best && c < bidx)) { best = dist; bidx = c; }
        }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float od = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
        if (od < best || (od == best && oi < bidx)) { best = od; bidx = oi; }
      }
      // NaN / Inf row: no distance compares below +inf (the oracle's strict '<' from +inf keeps index 0; vq_tc.cu agrees)
      if (bidx == 0x7fffffff || !(best < INFINITY)) bidx = 0;
      if (lane == 0) {
        widx[r] = bidx;
        if (r < nrows && indices) indices[b * L + l0 + r] = bidx;
      }
    }
    __syncthreads();
    if (zq_bdl) {  // [b, k, l0 + r]: r fastest -> coalesced 256-byte runs
      for (int i = tid; i < D * ROWS; i += THREADS) {
        const int k = i / ROWS, r = i - k * ROWS;
        if (r < nrows) zq_bdl[(b * D + k) * L + l0 + r] = cb[widx[r] * DP + k];
      }
    }
    if (zq_rows) {
      for (int i = tid; i < ROWS * D; i += THREADS) {
        const int r = i / D, k = i - r * D;
        if (r < nrows) zq_rows[(b * L + l0 + r) * D + k] = cb[widx[r] * DP + k];
      }
    }
  }
}

template <int D>
int launch_vq(const float* z, const float* codebook, const int64_t* code_offset, int64_t B, int64_t L, int n_codes,
              int64_t* indices, float* zq_bdl, float* zq_rows, cudaStream_t stream) {
  constexpr int DP = D + 4;
  const size_t smem = sizeof(float) * (256 * DP + ROWS * DP + 256 + ROWS) + sizeof(int) * ROWS;
  static bool attr = false;
  if (!attr) {
    FDM_CHECK_CUDA(cudaFuncSetAttribute(vq_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    attr = true;
  }
  const int64_t tiles = B * ceil_div64(L, ROWS);
  const int per_sm = smem * 2 <= 220 * 1024 ? 2 : 1;
  const int64_t cap = static_cast<int64_t>(fdm_sm_count()) * per_sm;
  const int grid = static_cast<int>(tiles < cap ? tiles : cap);
  vq_kernel<D><<<grid, THREADS, smem, stream>>>(z, codebook, code_offset, B, L, n_codes, indices, zq_bdl, zq_rows);
  FDM_CHECK_LAUNCH();
  return 0;
}

}  // namespace

// tensor-core filter + exact recheck (vq_tc.cu), D = 64 / 128
int fdm_vq_tc_launch(const float* z, const float* codebook, const int64_t* code_offset, int64_t B, int64_t L, int64_t D,
                     int n_codes, int64_t* indices, float* zq_bdl, float* zq_rows, unsigned long long* recheck_rows,
                     float* dbg_acc, cudaStream_t stream);

extern "C" int fdm_vq_quantize_ex(const float* z, const float* codebook, const int64_t* code_offset, int64_t B, int64_t L,
                                  int64_t D, int64_t n_codes, int64_t* indices, float* zq_bdl, float* zq_rows, int32_t algo,
                                  uint64_t* recheck_rows, float* dbg_acc, void* stream) {
  FDM_CHECK_ARG(z && codebook && B > 0 && L > 0, "fdm_vq_quantize: bad arguments");
  FDM_CHECK_ARG(n_codes > 0 && n_codes <= 256 && n_codes % 32 == 0, "fdm_vq_quantize: n_codes must be a multiple of 32, <= 256");
  FDM_CHECK_ARG(reinterpret_cast<uintptr_t>(z) % 16 == 0 && reinterpret_cast<uintptr_t>(codebook) % 16 == 0,
                "fdm_vq_quantize: z and codebook must be 16-byte aligned");
  FDM_CHECK_ARG(algo >= FDM_VQ_AUTO && algo <= FDM_VQ_TENSOR, "fdm_vq_quantize: unknown algo %d", algo);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (algo == FDM_VQ_AUTO) algo = (D == 64 || D == 128) ? FDM_VQ_TENSOR : FDM_VQ_FFMA;
  if (algo == FDM_VQ_TENSOR) {
    FDM_CHECK_ARG(D == 64 || D == 128, "fdm_vq_quantize: the tensor-core path needs D = 64 or 128 (got %lld)", (long long)D);
    return fdm_vq_tc_launch(z, codebook, code_offset, B, L, D, static_cast<int>(n_codes), indices, zq_bdl, zq_rows,
                            reinterpret_cast<unsigned long long*>(recheck_rows), dbg_acc, s);
  }
  FDM_CHECK_ARG(!dbg_acc, "fdm_vq_quantize: dbg_acc is a tensor-core path output");
  switch (D) {
    case 32: return launch_vq<32>(z, codebook, code_offset, B, L, static_cast<int>(n_codes), indices, zq_bdl, zq_rows, s);
    case 64: return launch_vq<64>(z, codebook, code_offset, B, L, static_cast<int>(n_codes), indices, zq_bdl, zq_rows, s);
    case 128: return launch_vq<128>(z, codebook, code_offset, B, L, static_cast<int>(n_codes), indices, zq_bdl, zq_rows, s);
    default: FDM_CHECK_ARG(false, "fdm_vq_quantize: D=%lld not in {32,64,128}", (long long)D);
  }
  return 0;
}

extern "C" int fdm_vq_quantize(const float* z, const float* codebook, const int64_t* code_offset, int64_t B, int64_t L,
                               int64_t D, int64_t n_codes, int64_t* indices, float* zq_bdl, float* zq_rows, void* stream) {
  return fdm_vq_quantize_ex(z, codebook, code_offset, B, L, D, n_codes, indices, zq_bdl, zq_rows, FDM_VQ_AUTO, nullptr, nullptr,
                            stream);
}
