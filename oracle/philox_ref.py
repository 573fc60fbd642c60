"""numpy restatement of the counter-based sampler noise used for throughput runs (TEST INFRASTRUCTURE).

Philox4x32-10 keyed by (seed); counter = (elem//4 lo, elem//4 hi, step t, global clip index); four uint32
-> two Box-Muller pairs. Mirrors face-diffusion-model_b200/csrc/ddpm.cu philox_normal4. The reference draws
torch.randn_like per step (diffusion_BIWI_encoder_decoder.py:654); this generator replaces it only in
throughput runs, parity runs inject host noise into both implementations.
"""
import numpy as np

M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    c0, c1, c2, c3 = [np.asarray(c, dtype=np.uint64) & 0xFFFFFFFF for c in (c0, c1, c2, c3)]
    k0, k1 = np.uint64(k0 & 0xFFFFFFFF), np.uint64(k1 & 0xFFFFFFFF)
    for _ in range(10):
        p0 = np.uint64(M0) * c0
        p1 = np.uint64(M1) * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & np.uint64(0xFFFFFFFF)
        hi1, lo1 = p1 >> np.uint64(32), p1 & np.uint64(0xFFFFFFFF)
        c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
        k0 = (k0 + np.uint64(W0)) & np.uint64(0xFFFFFFFF)
        k1 = (k1 + np.uint64(W1)) & np.uint64(0xFFFFFFFF)
    return c0, c1, c2, c3


def philox_normal(seed: int, clip: int, t: int, n: int) -> np.ndarray:
    """n (multiple of 4) standard normals for one clip at step t, float32."""
    assert n % 4 == 0
    e4 = np.arange(n // 4, dtype=np.uint64)
    z = np.zeros_like(e4)
    r0, r1, r2, r3 = philox4x32_10(e4 & np.uint64(0xFFFFFFFF), e4 >> np.uint64(32), z + np.uint64(t), z + np.uint64(clip),
                                   seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    inv24 = np.float32(2.0 ** -24)
    u0 = ((r0 >> np.uint64(8)) + np.uint64(1)).astype(np.float32) * inv24
    u1 = (r1 >> np.uint64(8)).astype(np.float32) * inv24
    u2 = ((r2 >> np.uint64(8)) + np.uint64(1)).astype(np.float32) * inv24
    u3 = (r3 >> np.uint64(8)).astype(np.float32) * inv24
    two_pi = np.float32(6.283185307179586)
    ra = np.sqrt(np.float32(-2.0) * np.log(u0)).astype(np.float32)
    rb = np.sqrt(np.float32(-2.0) * np.log(u2)).astype(np.float32)
    out = np.stack([ra * np.cos(two_pi * u1), ra * np.sin(two_pi * u1), rb * np.cos(two_pi * u3), rb * np.sin(two_pi * u3)], axis=1)
    return out.astype(np.float32).reshape(-1)
