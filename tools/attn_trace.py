#!/usr/bin/env python
"""Per-phase cycle stamps of attention_tc2.cu (instrumented build: `make -C face-diffusion-model_b200/csrc trace`), CTA 0,
items 1-6: where the ~16k cycles of one (sequence, head) item go. Run with FDM_B200_LIB=tools/_trace/libfdm_trace.so."""
import ctypes as C
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ.setdefault("FDM_B200_LIB", os.path.join(ROOT, "tools", "_trace", "libfdm_trace.so"))
sys.path[:0] = [ROOT, os.path.join(ROOT, "face-diffusion-model_b200")]
import torch  # noqa: E402
from fdm_b200 import lib  # noqa: E402

h = lib.require_device()
dev = torch.device("cuda:0")
B, H, dh, T = 128, 8, 128, int(os.environ.get("T", "198"))
causal = os.environ.get("CAUSAL", "1") == "1"
d = H * dh
qkv = torch.randn(B * T, 3 * d, device=dev).bfloat16()
o = torch.empty(B * T, d, device=dev, dtype=torch.bfloat16)
slopes = torch.tensor([2.0 ** (-(2.0 ** -(math.log2(H) - 3)) * (i + 1)) for i in range(H)], device=dev) if causal else None
for _ in range(3):
    lib.self_attention(qkv[:, 0:], qkv[:, d:], qkv[:, 2 * d:], o, B, T, T, H, dh, 1.0 / math.sqrt(dh), slopes=slopes, period=30)
torch.cuda.synchronize()
buf = (C.c_longlong * (8 * 2 * 16))()
h.fdm_attn_trace_read.restype = C.c_int
assert h.fdm_attn_trace_read(buf) == 0
names_s = ["start", "S0 ready", "ld+max0", "bar1", "exp0+P0", "S1 ready", "ld+max1", "bar2", "O0 ready", "epi0", "exp1+P1", "bar3", "O1 ready", "epi1"]
names_m = ["top", "K,Q0,OE0 in", "Q1 in", "P0 in", "V in", "pv0 issued", "P1 in", "OE1 in", "pv1 issued"]
for it in range(1, 6):
    s = [buf[(it * 2) * 16 + i] for i in range(14)]
    m = [buf[(it * 2 + 1) * 16 + i] for i in range(9)]
    t0 = s[0]
    print(f"item {it}: softmax warp " + " | ".join(f"{n} {v - t0}" for n, v in zip(names_s, s)))
    print(f"         mma thread   " + " | ".join(f"{n} {v - t0}" for n, v in zip(names_m, m)))
    nxt = buf[((it + 1) * 2) * 16]
    print(f"         item period {nxt - t0} cycles")

buf2 = (C.c_longlong * (8 * 16 * 8))()
h.fdm_attn_trace2_read.restype = C.c_int
assert h.fdm_attn_trace2_read(buf2) == 0
it = 3
t0 = buf[(it * 2) * 16]
print("item 3, per softmax warp (quad, part): max0 done | exp0 done | O0 seen | epi0 done | exp1 done")
for w in range(16):
    v = [buf2[(it * 16 + w) * 8 + i] - t0 for i in range(6)]
    print(f"  warp {w + 2} (quad {(w + 2) & 3}, part {w >> 2}): {v[4]} | {v[5]} | {v[0]} | {v[1]} | {v[2]}")
