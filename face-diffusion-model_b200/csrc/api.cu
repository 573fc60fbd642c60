// Library-level entry points: error string, device checks.
#include "common.cuh"
#include <stdarg.h>
#include <string.h>
#include <stdlib.h>

static thread_local char g_err[512] = "";

void fdm_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char* fdm_last_error(void) { return g_err; }

extern "C" int fdm_abi_version(void) { return 4; }  // 3: split-bf16 GEMM operands, fdm_ddpm_args.seed_dev, fdm_split_bf16x2; 4: fdm_gemm_args.a_group_cols / splitk_ws, tensor-core VQ for D = 128

int fdm_sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

bool fdm_pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("FDM_B200_PDL");
    // on unless FDM_B200_PDL=0: neutral on the power-capped VOCASET step (4.44 vs 4.45 ms), +5.6 % on the launch-bound
    // MEAD step (d = 512: 1.026 -> 0.978 ms for 74 kernels)
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

extern "C" int fdm_device_info(int32_t* sm_count, int32_t* cc_major, int32_t* cc_minor) {
  int dev = 0, major = 0, minor = 0, sms = 0;
  FDM_CHECK_CUDA(cudaGetDevice(&dev));
  FDM_CHECK_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  FDM_CHECK_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
  FDM_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  if (sm_count) *sm_count = sms;
  if (cc_major) *cc_major = major;
  if (cc_minor) *cc_minor = minor;
  FDM_CHECK_ARG(major == 10, "libfdm_b200 is built for sm_100a only; device %d is sm_%d%d", dev, major, minor);
  return fdm_gemm_init_device();  // per-device constants (allocated here, never inside a stream capture)
}
