// fdm_self_attention, exact-fp32 variant: flash-style online softmax on FFMA with K/V tiles staged in shared
// memory. Used for the fp32 parity mode and as the fallback for shapes the tensor-core variant
// (attention_mma.cu) does not take. The FaceFormer-style periodic ALiBi bias and the causal mask are computed
// in-kernel from (head, t - j, period); the (h, T, T) mask tensor of the reference is never materialised.
#include "common.cuh"

int fdm_attention_mma_try(const fdm_attn_args& a, cudaStream_t stream, bool* handled);  // attention_mma.cu
int fdm_attention_tc_try(const fdm_attn_args& a, cudaStream_t stream, bool* handled);   // attention_tc.cu
int fdm_attention_tc2_try(const fdm_attn_args& a, cudaStream_t stream, bool* handled);  // attention_tc2.cu
int fdm_attention_tc3_try(const fdm_attn_args& a, cudaStream_t stream, bool* handled);  // attention_tc3.cu

namespace {

constexpr int QB = 32;       // queries per CTA
constexpr int KB = 64;       // keys per tile
constexpr int WARPS = 8;     // 4 queries per warp
constexpr int QPW = QB / WARPS;

template <int NI>  // dh = 32 * NI
__global__ void __launch_bounds__(WARPS * 32) attn_f32_kernel(const fdm_attn_args a) {
  constexpr int DH = 32 * NI;
  extern __shared__ float sm[];
  float* Qs = sm;                      // [QB][DH]
  float* Ks = Qs + QB * DH;            // [KB][DH + 1]
  float* Vs = Ks + KB * (DH + 1);      // [KB][DH]

  const int T = static_cast<int>(a.T);
  const int q0 = blockIdx.x * QB;
  const int h = blockIdx.y;
  const int64_t row0 = static_cast<int64_t>(blockIdx.z) * a.t_stride;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t hoff = static_cast<int64_t>(h) * DH;
  const bool causal = a.bias_mode == 1;
  const float slope = causal ? a.slopes[h] : 0.f;

  for (int idx = threadIdx.x; idx < QB * DH; idx += blockDim.x) {
    const int qi = idx / DH, dd = idx - qi * DH;
    const int t = q0 + qi;
    Qs[idx] = t < T ? ld_as_float(a.Q, a.dtype, (row0 + t) * a.ldq + hoff + dd) : 0.f;
  }

  float m[QPW], l[QPW], o[QPW][NI];
#pragma unroll
  for (int i = 0; i < QPW; ++i) {
    m[i] = -INFINITY;
    l[i] = 0.f;
#pragma unroll
    for (int j = 0; j < NI; ++j) o[i][j] = 0.f;
  }

  const int q_last = min(q0 + QB, T) - 1;
  const int k_end = causal ? q_last + 1 : T;
  for (int k0 = 0; k0 < k_end; k0 += KB) {
    __syncthreads();  // previous tile fully consumed (also orders the Qs fill before first use)
    for (int idx = threadIdx.x; idx < KB * DH; idx += blockDim.x) {
      const int kj = idx / DH, dd = idx - kj * DH;
      const int j = k0 + kj;
      float kv = 0.f, vv = 0.f;
      if (j < T) {
        kv = ld_as_float(a.K, a.dtype, (row0 + j) * a.ldk + hoff + dd);
        vv = ld_as_float(a.V, a.dtype, (row0 + j) * a.ldv + hoff + dd);
      }
      Ks[kj * (DH + 1) + dd] = kv;
      Vs[kj * DH + dd] = vv;
    }
    __syncthreads();

    float s[QPW][2];
#pragma unroll
    for (int i = 0; i < QPW; ++i) s[i][0] = s[i][1] = 0.f;
    const float* kr0 = Ks + lane * (DH + 1);
    const float* kr1 = Ks + (lane + 32) * (DH + 1);
#pragma unroll 4
    for (int dd = 0; dd < DH; ++dd) {
      const float ka = kr0[dd], kb = kr1[dd];
#pragma unroll
      for (int i = 0; i < QPW; ++i) {
        const float qv = Qs[(warp * QPW + i) * DH + dd];
        s[i][0] = fmaf(qv, ka, s[i][0]);
        s[i][1] = fmaf(qv, kb, s[i][1]);
      }
    }
    float p[QPW][2];
#pragma unroll
    for (int i = 0; i < QPW; ++i) {
      const int t = q0 + warp * QPW + i;
      float sv[2];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int j = k0 + lane + 32 * e;
        float x = s[i][e] * a.scale;
        bool valid = j < T;
        if (causal) {
          valid = valid && j <= t;
          x -= slope * static_cast<float>((t - j) / a.period);
        }
        sv[e] = valid ? x : -INFINITY;
      }
      const float m_new = fmaxf(m[i], warp_max(fmaxf(sv[0], sv[1])));
      // m_new is finite for every in-range query after the first tile (key 0 is always visible)
      const float corr = m_new == -INFINITY ? 1.f : expf(m[i] - m_new);
      p[i][0] = m_new == -INFINITY ? 0.f : expf(sv[0] - m_new);
      p[i][1] = m_new == -INFINITY ? 0.f : expf(sv[1] - m_new);
      l[i] = l[i] * corr + warp_sum(p[i][0] + p[i][1]);
      m[i] = m_new;
#pragma unroll
      for (int j = 0; j < NI; ++j) o[i][j] *= corr;
    }
    const int kmax = min(KB, k_end - k0);
    for (int kj = 0; kj < kmax; ++kj) {
      float vv[NI];
#pragma unroll
      for (int j = 0; j < NI; ++j) vv[j] = Vs[kj * DH + lane + 32 * j];
#pragma unroll
      for (int i = 0; i < QPW; ++i) {
        const float pk = __shfl_sync(0xffffffffu, kj < 32 ? p[i][0] : p[i][1], kj & 31);
#pragma unroll
        for (int j = 0; j < NI; ++j) o[i][j] = fmaf(pk, vv[j], o[i][j]);
      }
    }
  }

#pragma unroll
  for (int i = 0; i < QPW; ++i) {
    const int t = q0 + warp * QPW + i;
    if (t < T) {
      const float inv = 1.f / l[i];
#pragma unroll
      for (int j = 0; j < NI; ++j)
        st_from_float(a.O, a.dtype, (row0 + t) * a.ldo + hoff + lane + 32 * j, o[i][j] * inv);
    }
  }
}

template <int NI>
int launch_f32(const fdm_attn_args& a, cudaStream_t stream) {
  constexpr int DH = 32 * NI;
  const size_t smem = sizeof(float) * (QB * DH + KB * (DH + 1) + KB * DH);
  static bool attr = false;
  if (!attr) {
    FDM_CHECK_CUDA(cudaFuncSetAttribute(attn_f32_kernel<NI>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    attr = true;
  }
  dim3 grid(static_cast<unsigned>(ceil_div64(a.T, QB)), static_cast<unsigned>(a.H), static_cast<unsigned>(a.B));
  attn_f32_kernel<NI><<<grid, WARPS * 32, smem, stream>>>(a);
  FDM_CHECK_LAUNCH();
  return 0;
}

}  // namespace

extern "C" int fdm_self_attention(const fdm_attn_args* args, void* stream) {
  FDM_CHECK_ARG(args != nullptr, "fdm_self_attention: null args");
  const fdm_attn_args& a = *args;
  FDM_CHECK_ARG(a.Q && a.K && a.V && a.O, "fdm_self_attention: null operand");
  FDM_CHECK_ARG(a.B > 0 && a.T > 0 && a.H > 0 && a.t_stride >= a.T, "fdm_self_attention: bad sizes");
  FDM_CHECK_ARG(a.B <= 65535 && a.H <= 65535, "fdm_self_attention: B and H must be <= 65535");
  FDM_CHECK_ARG(a.bias_mode == 0 || (a.bias_mode == 1 && a.slopes && a.period > 0), "fdm_self_attention: bad bias mode");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (a.dtype == FDM_BF16) {
    bool handled = false;
    int rc = fdm_attention_tc2_try(a, s, &handled);  // tcgen05/TMEM kernel, 16 softmax warps: head dim 128, T <= 208, causal or unmasked
    if (rc != 0 || handled) return rc;
    rc = fdm_attention_tc_try(a, s, &handled);       // first-generation tcgen05 kernel (FDM_B200_ATTN_TC=1)
    if (rc != 0 || handled) return rc;
    rc = fdm_attention_tc3_try(a, s, &handled);      // general tcgen05 kernel: head dim 64 / 128 / 256, key blocks of 64, two-pass softmax
    if (rc != 0 || handled) return rc;
    rc = fdm_attention_mma_try(a, s, &handled);      // mma.sync kernel: head dim 64 / 128, any bias mode
    if (rc != 0 || handled) return rc;
  }
  switch (a.dh) {
    case 32: return launch_f32<1>(a, s);
    case 64: return launch_f32<2>(a, s);
    case 128: return launch_f32<4>(a, s);
    case 256: return launch_f32<8>(a, s);
    default: FDM_CHECK_ARG(false, "fdm_self_attention: head dim %lld not in {32,64,128,256}", (long long)a.dh);
  }
  return 0;
}
