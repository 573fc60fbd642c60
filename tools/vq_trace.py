#!/usr/bin/env python
"""Timeline of the tensor-core VQ kernel's pipeline roles for CTA 0 (needs a library built with -DVQ_TRACE, see
vq_tc.cu: `make -C face-diffusion-model_b200/csrc trace` writes tools/_trace/libfdm_trace.so; run with
FDM_B200_LIB=$PWD/tools/_trace/libfdm_trace.so). Prints per-tile clock64 stamps relative to the first one."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "face-diffusion-model_b200")]
import torch
from fdm_b200 import lib
lib.require_device()
dev = torch.device("cuda:0")
D, n = int(os.environ.get("VQ_D", "64")), 256
clips = int(os.environ.get("VQ_CLIPS", "256"))
L = int(os.environ.get("VQ_L", str(498 * 16)))
g = torch.Generator(device="cpu").manual_seed(1)
z = torch.randn(clips, L, D, device=dev)
cb = torch.randn(n, D, generator=g).to(dev)
want = os.environ.get("VQ_ZQ", "0") == "1"
for _ in range(3):
    lib.vq_quantize(z, cb, n, want_bdl=want, algo=lib.VQ_TENSOR)
buf = torch.zeros(clips * L * n, dtype=torch.float32, device=dev)  # trace build writes int64 stamps at the start
lib.vq_quantize(z, cb, n, want_bdl=want, algo=lib.VQ_TENSOR, dbg_acc=buf)
torch.cuda.synchronize()
st = buf[:96 * 32].view(torch.int64).view(96, 16).cpu()
t0 = st[st > 0].min().item()
names = ["cv_wait", "cv_go", "cv_done", "mma_in", "mma_tE", "mma_aF", "mma_iss", "ep_wait", "ep_go", "ep_scan", "ep_rel", "ep_out", "ld_iss",
         "cv0_ld", "cv0_go", "cv0_done"]  # (D = 128: the last three are the first K-half, cv_* the second)
print("tile " + " ".join(f"{n:>8}" for n in names))
for i in range(8, 72):
    print(f"{i:4d} " + " ".join(f"{(st[i, k].item() - t0) if st[i, k] > 0 else -1:8d}" for k in range(len(names))))
