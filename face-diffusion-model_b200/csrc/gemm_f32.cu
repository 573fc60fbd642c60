// fdm_gemm_f32: fp32 FFMA GEMM with a fixed (sequential-k) accumulation order.
// This is the "fp32 mode" of the path: it exists so that the end-to-end result can be compared with the
// reference PyTorch fp32 path at 1e-4 (BASELINE.json north_star); throughput runs use fdm_gemm_bf16.
// Same operand convention and fused epilogue as the tensor-core kernel, including implicit 1-D convolution.
#include "common.cuh"

namespace {

constexpr int BM = 128, BN = 128, BK = 16, THREADS = 256, PAD = 4;

struct F32Params {
  const float* A; const float* W; const float* bias; const void* residual; void* C;
  int64_t lda, ldw, ldr, ldc;
  int M, N, K;
  int res_dtype, out_dtype, act;
  int taps, tap_k, tap_row_shift;
  int vec_a, vec_w;
};

__device__ __forceinline__ const float* a_ptr(const F32Params& p, int64_t m, int k) {
  if (p.taps > 1) {
    const int tap = k / p.tap_k;
    const int c = k - tap * p.tap_k;
    return p.A + (m + static_cast<int64_t>(tap) * p.tap_row_shift) * p.lda + c;
  }
  return p.A + m * p.lda + k;
}

__global__ void __launch_bounds__(THREADS) gemm_f32_kernel(const F32Params p) {
  __shared__ float As[BK][BM + PAD];
  __shared__ float Bs[BK][BN + PAD];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int64_t m0 = static_cast<int64_t>(blockIdx.y) * BM;
  const int n0 = blockIdx.x * BN;

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < p.K; k0 += BK) {
    // each thread stages 2 x 4 consecutive k of A and of W
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int idx = tid + i * THREADS;
      const int r = idx >> 2, kq = (idx & 3) * 4;
      const int k = k0 + kq;
      float va[4] = {0.f, 0.f, 0.f, 0.f}, vw[4] = {0.f, 0.f, 0.f, 0.f};
      const int64_t m = m0 + r;
      if (m < p.M) {
        if (p.vec_a && k + 3 < p.K) {
          float4 t = *reinterpret_cast<const float4*>(a_ptr(p, m, k));
          va[0] = t.x; va[1] = t.y; va[2] = t.z; va[3] = t.w;
        } else {
#pragma unroll
          for (int q = 0; q < 4; ++q)
            if (k + q < p.K) va[q] = *a_ptr(p, m, k + q);
        }
      }
      const int n = n0 + r;
      if (n < p.N) {
        const float* w = p.W + static_cast<int64_t>(n) * p.ldw + k;
        if (p.vec_w && k + 3 < p.K) {
          float4 t = *reinterpret_cast<const float4*>(w);
          vw[0] = t.x; vw[1] = t.y; vw[2] = t.z; vw[3] = t.w;
        } else {
#pragma unroll
          for (int q = 0; q < 4; ++q)
            if (k + q < p.K) vw[q] = w[q];
        }
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        As[kq + q][r] = va[q];
        Bs[kq + q][r] = vw[q];
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[8], b[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = As[kk][ty * 8 + i];
#pragma unroll
      for (int j = 0; j < 8; ++j) b[j] = Bs[kk][tx * 8 + j];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t m = m0 + ty * 8 + i;
    if (m >= p.M) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int n = n0 + tx * 8 + j;
      if (n >= p.N) continue;
      float v = acc[i][j];
      if (p.bias) v += p.bias[n];
      v = apply_act(v, p.act);
      if (p.residual) v += ld_as_float(p.residual, p.res_dtype, m * p.ldr + n);
      st_from_float(p.C, p.out_dtype, m * p.ldc + n, v);
    }
  }
}

}  // namespace

extern "C" int fdm_gemm_f32(const fdm_gemm_args* args, void* stream) {
  FDM_CHECK_ARG(args != nullptr, "fdm_gemm_f32: null args");
  const fdm_gemm_args& a = *args;
  FDM_CHECK_ARG(a.A && a.W && a.C, "fdm_gemm_f32: null operand");
  FDM_CHECK_ARG(a.M > 0 && a.N > 0 && a.K > 0, "fdm_gemm_f32: empty problem");
  FDM_CHECK_ARG(a.M < (1ll << 31) && a.N < (1ll << 31) && a.K < (1ll << 31), "fdm_gemm_f32: dimension too large");
  if (a.taps > 1) FDM_CHECK_ARG(a.tap_k > 0 && a.K == a.taps * a.tap_k, "fdm_gemm_f32: implicit conv needs K == taps*tap_k");
  FDM_CHECK_ARG(!a.a_ln && !a.res_ln && !a.stats_out, "fdm_gemm_f32: LayerNorm folding is a bf16-path feature");
  FDM_CHECK_ARG(a.a_group_cols == 0, "fdm_gemm_f32: the grouped mode is a bf16-path feature");
  F32Params p;
  p.A = static_cast<const float*>(a.A);
  p.W = static_cast<const float*>(a.W);
  p.bias = a.bias;
  p.residual = a.residual;
  p.C = a.C;
  p.lda = a.lda; p.ldw = a.ldw; p.ldr = a.ldr; p.ldc = a.ldc;
  p.M = static_cast<int>(a.M); p.N = static_cast<int>(a.N); p.K = static_cast<int>(a.K);
  p.res_dtype = a.res_dtype; p.out_dtype = a.out_dtype; p.act = a.act;
  p.taps = a.taps > 1 ? a.taps : 1;
  p.tap_k = static_cast<int>(a.taps > 1 ? a.tap_k : a.K);
  p.tap_row_shift = static_cast<int>(a.taps > 1 ? a.tap_row_shift : 0);
  p.vec_a = (reinterpret_cast<uintptr_t>(a.A) % 16 == 0) && a.lda % 4 == 0 && p.tap_k % 4 == 0;
  p.vec_w = (reinterpret_cast<uintptr_t>(a.W) % 16 == 0) && a.ldw % 4 == 0;
  dim3 grid(static_cast<unsigned>(ceil_div64(a.N, BN)), static_cast<unsigned>(ceil_div64(a.M, BM)));
  gemm_f32_kernel<<<grid, THREADS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p);
  FDM_CHECK_LAUNCH();
  return 0;
}
