"""BIWI latent width (D = 128): tensor-core filter + exact recheck against the FFMA kernel, all rows, with timings.
Run on a B200: python tools/vq_d128_probe.py [clips]"""
import json, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "face-diffusion-model_b200"))
import torch
from fdm_b200 import lib

dev = torch.device("cuda:0")
clips = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
L, D, codes = 149 * 8, 128, 256
g = torch.Generator(device="cpu").manual_seed(7)
cb = (torch.randn(codes, D, generator=g) * 0.5).to(dev)
z = torch.randn(clips, L, D, device=dev)
out = {"rows": clips * L, "D": D}


def timed(fn, n, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for name, algo, n in (("tensor", lib.VQ_TENSOR, 10), ("ffma", lib.VQ_FFMA, 2)):
    for zq in (False, True):
        ms = timed(lambda: lib.vq_quantize(z, cb, codes, want_bdl=zq, algo=algo), n)
        alg = clips * L * (4 * D + 8 + (4 * D if zq else 0))
        out[f"{name}_{'zq' if zq else 'idx'}"] = {"ms": round(ms, 4), "GBps": round(alg / ms / 1e6, 1)}
cnt = torch.zeros(1, dtype=torch.int64, device=dev)
it, _, _ = lib.vq_quantize(z, cb, codes, want_bdl=False, algo=lib.VQ_TENSOR, recheck_rows=cnt)
i_f, _, _ = lib.vq_quantize(z, cb, codes, want_bdl=False, algo=lib.VQ_FFMA)
out["equal_all_rows"] = bool(torch.equal(it, i_f))
out["recheck_fraction"] = cnt.item() / (clips * L)
print(json.dumps(out))
