"""A/B of the GEMM's tail split-K (lib.splitk_enabled) on the shapes whose last wave is sparsely filled.
Run on a B200: python tools/tailk_probe.py"""
import json, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "face-diffusion-model_b200"))
import torch
from fdm_b200 import lib

dev = torch.device("cuda:0")
lib.require_device()
shapes = [("vocaset64 qkv", 25344, 3072, 1024), ("vocaset64 out", 25344, 1024, 1024), ("vocaset64 ffn1", 25344, 2048, 1024),
          ("vocaset64 ffn2", 25344, 1024, 2048), ("biwi16 qkv", 4768, 3072, 1024), ("biwi16 out", 4768, 1024, 1024),
          ("biwi16 ffn1", 4768, 2048, 1024), ("biwi16 ffn2", 4768, 1024, 2048), ("mead32 qkv", 12736, 1536, 512),
          ("mead32 out", 12736, 512, 512), ("mead32 ffn1", 12736, 1024, 512), ("mead32 ffn2", 12736, 512, 1024),
          ("biwi32 out", 9536, 1024, 1024), ("biwi64 out", 19072, 1024, 1024)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, n=30):
    for _ in range(5):
        fn()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    for e0, e1 in ev:
        flush.zero_()
        e0.record()
        fn()
        e1.record()
    torch.cuda.synchronize()
    t = sorted(e0.elapsed_time(e1) for e0, e1 in ev)
    return 1e3 * t[len(t) // 2]


out = {}
for name, M, N, K in shapes:
    a = torch.randn(M, K, device=dev).bfloat16()
    w = (torch.randn(N, K, device=dev) / K ** 0.5).bfloat16()
    c = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    r = {}
    for rep in range(2):
        for on in (True, False):
            lib.splitk_enabled = on
            us = timed(lambda: lib.gemm(a, w, c))
            r.setdefault("tailk" if on else "plain", []).append(round(us, 1))
    lib.splitk_enabled = False
    tiles = -(-M // 256) * -(-N // 256)
    out[name] = {"M": M, "N": N, "K": K, "tiles": tiles, "us": r, "TFLOPs_tailk": round(2 * M * N * K / min(r["tailk"]) / 1e6, 1),
                 "TFLOPs_plain": round(2 * M * N * K / min(r["plain"]) / 1e6, 1)}
    print(name, json.dumps(out[name]), flush=True)
