import os, sys, math, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'face-diffusion-model_b200')]
from fdm_b200 import lib
lib.require_device()
dev = torch.device('cuda:0')
S, H, T, dh = 128, 8, 198, 128
d = H * dh
g = torch.Generator(device='cpu').manual_seed(0)
qkv = torch.randn(S * T, 3 * d, generator=g).to(dev).bfloat16()
out = torch.zeros(S * T, d, device=dev, dtype=torch.bfloat16)
slopes = torch.tensor([2.0 ** -(i + 1) for i in range(H)], device=dev)
def run():
    lib.self_attention(qkv[:, 0:], qkv[:, d:], qkv[:, 2 * d:], out, S, T, T, H, dh, 1 / math.sqrt(dh), slopes=slopes, period=30)
for _ in range(5): run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50): run()
e1.record(); torch.cuda.synchronize()
print(os.environ.get('FDM_B200_ATTN_TC'), 'us per call', e0.elapsed_time(e1) * 1000 / 50, 'checksum', out.float().abs().mean().item())
# reference check on a few sequences
q, k, v = [t.float().view(S, T, H, dh)[:2].permute(0, 2, 1, 3) for t in qkv.split(d, dim=-1)]
i = torch.arange(T, device=dev)[:, None]; j = torch.arange(T, device=dev)[None, :]
bias = (-slopes[:, None, None] * ((i - j) // 30).float()[None]).masked_fill((j > i)[None], float('-inf'))
ref = ((q @ k.transpose(-1, -2)) / math.sqrt(dh) + bias).softmax(-1) @ v
got = out.view(S, T, H, dh)[:2].permute(0, 2, 1, 3).float()
print('rel err', ((got - ref).norm() / ref.norm()).item())
