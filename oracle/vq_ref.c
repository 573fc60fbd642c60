/* CPU oracle for the EVQ-VAE nearest-code search (TEST INFRASTRUCTURE, see oracle/__init__.py).
 *
 * Restates models/lib/quantizer.py:35-64 and models/vq_vae_emotion.py:221-252 of the reference:
 *     d = sum(z**2, dim=1, keepdim) + sum(e**2, dim=1) - 2 * z @ e.T ;  idx = argmin(d, dim=1)
 * with a DEFINED fp32 evaluation order (the reference's BLAS order is unspecified): every sum is a
 * sequential fmaf chain over k = 0..D-1 from +0.0f, d_j = (zz + ee_j) - 2*dot_j with each operation
 * rounded to fp32, and the lowest index wins ties (torch.argmin semantics).
 * Build: gcc -O2 -ffp-contract=off -shared -fPIC vq_ref.c -o _build/libvq_ref.so -lm
 */
#include <math.h>
#include <stdint.h>

void vq_ref_quantize(const float* z, const float* codebook, int64_t rows, int64_t D, int64_t n_codes,
                     int64_t* indices, float* dist_best, float* dist_second) {
  for (int64_t r = 0; r < rows; ++r) {
    const float* zr = z + r * D;
    float zz = 0.0f;
    for (int64_t k = 0; k < D; ++k) zz = fmaf(zr[k], zr[k], zz);
    float best = INFINITY, second = INFINITY;
    int64_t bi = 0;
    for (int64_t j = 0; j < n_codes; ++j) {
      const float* e = codebook + j * D;
      float ee = 0.0f, dot = 0.0f;
      for (int64_t k = 0; k < D; ++k) ee = fmaf(e[k], e[k], ee);
      for (int64_t k = 0; k < D; ++k) dot = fmaf(zr[k], e[k], dot);
      volatile float s = zz + ee;
      volatile float t2 = 2.0f * dot;
      const float d = s - t2;
      if (d < best) { second = best; best = d; bi = j; }
      else if (d < second) second = d;
    }
    indices[r] = bi;
    if (dist_best) dist_best[r] = best;
    if (dist_second) dist_second[r] = second;
  }
}
