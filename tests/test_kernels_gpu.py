"""Per-kernel numerics: each C-ABI kernel against a plain PyTorch fp32 reference of the same op (GPU)."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-12)).item()


GEMM_SHAPES = [
    # M, N, K
    (128, 256, 64), (256, 256, 128), (200, 1024, 1024), (198, 3072, 1024), (792, 1024, 2048),
    (1000, 512, 512), (130, 64, 64), (257, 192, 320), (12672, 1024, 1024), (333, 15069, 1024),
    (5000, 2048, 512), (4099, 1280, 192), (25344, 3072, 1024),  # CTA-pair (cta_group::2) kernel, with M / N / K tails
    (2100, 15069, 256), (700, 1001, 128),  # fp32 rows that are not 16-byte aligned: 1-D TMA row stores, ragged last chunk
]


@pytest.mark.parametrize("M,N,K", GEMM_SHAPES)
@pytest.mark.parametrize("out_dtype", [torch.bfloat16, torch.float32])
def test_gemm_bf16_tcgen05(cuda_dev, M, N, K, out_dtype):
    from fdm_b200 import lib
    g = torch.Generator(device="cpu").manual_seed(M * 7 + N * 3 + K)
    a = torch.randn(M, K, generator=g).to(cuda_dev).bfloat16()
    w = (torch.randn(N, K, generator=g) / math.sqrt(K)).to(cuda_dev).bfloat16()
    bias = torch.randn(N, generator=g).to(cuda_dev)
    out = torch.full((M, N), float("nan"), device=cuda_dev, dtype=out_dtype)
    lib.gemm(a, w, out, bias=bias)
    torch.cuda.synchronize()
    ref = a.float() @ w.float().t() + bias
    tol = 1e-2 if out_dtype == torch.bfloat16 else 2e-5
    assert torch.isfinite(out.float()).all()
    assert _rel(out, ref) < tol
    err = (out.float() - ref).abs().max().item()
    assert err < (0.1 if out_dtype == torch.bfloat16 else 1e-3), err


@pytest.mark.parametrize("act", [1, 2, 3, 4, 5])
def test_gemm_bf16_epilogue(cuda_dev, act):
    from fdm_b200 import lib
    import torch.nn.functional as F
    M, N, K = 300, 1024, 512
    g = torch.Generator(device="cpu").manual_seed(act)
    a = torch.randn(M, K, generator=g).to(cuda_dev).bfloat16()
    w = (torch.randn(N, K, generator=g) / math.sqrt(K)).to(cuda_dev).bfloat16()
    bias = torch.randn(N, generator=g).to(cuda_dev)
    res = torch.randn(M, N, generator=g).to(cuda_dev)
    fns = {1: F.relu, 2: F.mish, 3: F.gelu, 4: lambda x: F.gelu(x, approximate="tanh"), 5: lambda x: F.leaky_relu(x, 0.2)}
    ref = fns[act](a.float() @ w.float().t() + bias) + res
    for res_t in (res, res.bfloat16()):
        out = torch.empty(M, N, device=cuda_dev)
        lib.gemm(a, w, out, bias=bias, act=act, residual=res_t)
        torch.cuda.synchronize()
        assert _rel(out, fns[act](a.float() @ w.float().t() + bias) + res_t.float()) < 2e-5


@pytest.mark.parametrize("M,N,K", [(300, 1024, 512), (130, 200, 64), (1000, 3072, 1024), (128, 64, 128), (6000, 1024, 2048),
                                   (5003, 1288, 320)])
def test_gemm_bf16_out_with_bf16_residual_tma_path(cuda_dev, M, N, K):
    """bf16 output + bf16 residual: both move through TMA (staging tile); includes the in-place x = x + f(x) form."""
    from fdm_b200 import lib
    g = torch.Generator(device="cpu").manual_seed(M + N)
    a = torch.randn(M, K, generator=g).to(cuda_dev).bfloat16()
    w = (torch.randn(N, K, generator=g) / math.sqrt(K)).to(cuda_dev).bfloat16()
    bias = torch.randn(N, generator=g).to(cuda_dev)
    res = torch.randn(M, N, generator=g).to(cuda_dev).bfloat16()
    ref = torch.relu(a.float() @ w.float().t() + bias) + res.float()
    out = torch.full((M, N), float("nan"), device=cuda_dev, dtype=torch.bfloat16)
    lib.gemm(a, w, out, bias=bias, act=lib.ACT_RELU, residual=res)
    x = res.clone()
    lib.gemm(a, w, x, bias=bias, act=lib.ACT_RELU, residual=x)  # in place
    torch.cuda.synchronize()
    assert torch.isfinite(out.float()).all()
    assert _rel(out, ref) < 6e-3
    assert torch.equal(out, x)


@pytest.mark.parametrize("M,N,K", [(300, 1024, 512), (130, 200, 64), (128, 64, 128), (6000, 1024, 2048), (5003, 1288, 320),
                                   (12672, 1024, 1024), (25344, 1024, 2048), (4099, 1280, 192), (700, 72, 64)])
@pytest.mark.parametrize("out_dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("resmma", [0, 1])
def test_gemm_bf16_residual_through_mma(cuda_dev, M, N, K, out_dtype, resmma):
    """bf16 residual, no activation (out-projection / FFN2 of the denoiser layers), both residual paths: the TMA-residual
    epilogue (resmma = 0) and the residual tile accumulated by the tensor core as identity k-blocks (resmma = 1,
    fdm_gemm_set_option). All three tile kernels, M / N / K tails, fp32 and bf16 outputs, the in-place form x = x + f(x)."""
    from fdm_b200 import lib
    lib.gemm_set_option("resmma", resmma)
    g = torch.Generator(device="cpu").manual_seed(M + 3 * N + K)
    a = torch.randn(M, K, generator=g).to(cuda_dev).bfloat16()
    w = (torch.randn(N, K, generator=g) / math.sqrt(K)).to(cuda_dev).bfloat16()
    bias = torch.randn(N, generator=g).to(cuda_dev)
    ld = (N + 7) // 8 * 8  # residual rows 16-byte aligned (what the TMA path needs); N itself may be ragged
    res_buf = (3 * torch.randn(M, ld, generator=g)).to(cuda_dev).bfloat16()
    res = res_buf[:, :N]
    ref = a.float() @ w.float().t() + bias + res.float()
    out = torch.full((M, N), float("nan"), device=cuda_dev, dtype=out_dtype)
    lib.gemm(a, w, out, bias=bias, residual=res)
    torch.cuda.synchronize()
    assert torch.isfinite(out.float()).all()
    assert _rel(out, ref) < (6e-3 if out_dtype == torch.bfloat16 else 2e-5)
    err = (out.float() - ref).abs().max().item()
    assert err < (0.15 if out_dtype == torch.bfloat16 else 2e-3), err
    if out_dtype == torch.bfloat16:
        x_buf = res_buf.clone()
        lib.gemm(a, w, x_buf[:, :N], bias=bias, residual=x_buf[:, :N])  # in place
        torch.cuda.synchronize()
        assert torch.equal(out, x_buf[:, :N])
    lib.gemm_set_option("resmma", 0)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_gemm_implicit_conv(cuda_dev, dtype):
    """Conv1d(k=5, pad=2 replicate) and a stride-2 k=3 conv as shifted-row / overlapping-row GEMMs."""
    from fdm_b200 import lib
    import torch.nn.functional as F
    B, T, Cin, Cout = 3, 50, 128, 192
    g = torch.Generator(device="cpu").manual_seed(5)
    x = torch.randn(B, T, Cin, generator=g).to(cuda_dev)
    wt = (torch.randn(Cout, Cin, 5, generator=g) / math.sqrt(5 * Cin)).to(cuda_dev)
    bias = torch.randn(Cout, generator=g).to(cuda_dev)
    if dtype == torch.bfloat16:
        x, wt = x.bfloat16().float(), wt.bfloat16().float()
    ref = F.conv1d(F.pad(x.double().transpose(1, 2), (2, 2), mode="replicate"), wt.double(), bias.double()).transpose(1, 2).float()
    xp = torch.empty(B, T + 4, Cin, device=cuda_dev, dtype=dtype)
    lib.pad_time(x.to(dtype).contiguous(), xp, B, T, Cin, 2, 2, 1)
    wk = wt.permute(0, 2, 1).reshape(Cout, 5 * Cin).contiguous().to(dtype)  # [Cout, tap, Cin]
    out = torch.zeros(B * (T + 4), Cout, device=cuda_dev)
    lib.gemm(xp, wk, out, bias=bias, M=B * (T + 4) - 4, lda=Cin, a_rows=B * (T + 4), taps=5, tap_k=Cin, tap_row_shift=1)
    torch.cuda.synchronize()
    got = out.view(B, T + 4, Cout)[:, :T]
    assert _rel(got, ref) < (2e-5 if dtype == torch.float32 else 1e-4)

    # stride-2, k=3, no padding: rows overlap (lda = 2*Cin, K = 3*Cin)
    Lin = 64
    x2 = torch.randn(B, Lin, Cin, generator=g).to(cuda_dev)
    w2 = (torch.randn(Cout, Cin, 3, generator=g) / math.sqrt(3 * Cin)).to(cuda_dev)
    if dtype == torch.bfloat16:
        x2, w2 = x2.bfloat16().float(), w2.bfloat16().float()
    ref2 = F.conv1d(x2.double().transpose(1, 2), w2.double(), None, stride=2).transpose(1, 2).float()  # B, 31, Cout
    Lout = (Lin - 3) // 2 + 1
    w2k = w2.permute(0, 2, 1).reshape(Cout, 3 * Cin).contiguous().to(dtype)
    xin = torch.zeros(B * Lin + 2, Cin, device=cuda_dev, dtype=dtype)  # 2 slack rows for the last window
    xin[:B * Lin] = x2.reshape(B * Lin, Cin).to(dtype)
    out2 = torch.zeros(B * Lin // 2, Cout, device=cuda_dev)
    lib.gemm(xin, w2k, out2, M=B * Lin // 2, lda=2 * Cin, a_rows=B * Lin // 2, K=3 * Cin)
    torch.cuda.synchronize()
    got2 = out2.view(B, Lin // 2, Cout)[:, :Lout]
    assert _rel(got2, ref2) < (2e-5 if dtype == torch.float32 else 1e-4)


@pytest.mark.parametrize("M,N,K,act,res", [
    (4768, 1024, 1024, 0, False),    # BIWI at 16 clips per GPU: 76 tiles on 74 CTA pairs -> 2 tail tiles x 8 K-slices
    (4768, 3072, 1024, 0, False),    # 228 tiles = 3 waves + 6
    (4768, 1000, 2048, 1, True),     # ragged M and N edges inside the tail tiles, bias + ReLU + fp32 residual in the finaliser
    (12736, 1536, 512, 3, False),    # MEAD at 32 clips per GPU: 300 tiles = 4 waves + 4, 4 slices of 2 k-blocks
    (25344, 1024, 2048, 0, False),   # the headline FFN2: 396 tiles = 5 waves + 26 -> 2 slices each
])
@pytest.mark.parametrize("out_dtype", [torch.bfloat16, torch.float32])
def test_gemm_tail_split_k(cuda_dev, M, N, K, act, res, out_dtype):
    """Tail split-K (TailK in gemm_tc.cu): the tiles of a sparsely filled last wave are cut along K, partial fp32 tiles meet
    in the workspace and the last slice to arrive finishes the tile. Checked against torch, against the unsplit kernel,
    and for run-to-run determinism (fixed summation order: no atomics on data)."""
    from fdm_b200 import lib
    import torch.nn.functional as F
    g = torch.Generator(device="cpu").manual_seed(M + N + K)
    a = torch.randn(M, K, generator=g).to(cuda_dev).bfloat16()
    w = (torch.randn(N, K, generator=g) / math.sqrt(K)).to(cuda_dev).bfloat16()
    bias = torch.randn(N, generator=g).to(cuda_dev)
    r = torch.randn(M, N, generator=g).to(cuda_dev) if res else None
    fn = {0: lambda x: x, 1: F.relu, 3: F.gelu}[act]
    ref = fn(a.float() @ w.float().t() + bias) + (r if res else 0)
    outs = []
    was = lib.splitk_enabled
    try:
        lib.splitk_enabled = True  # opt-in feature (default off)
        for rep in range(3):
            out = torch.full((M, N), float("nan"), device=cuda_dev, dtype=out_dtype)
            lib.gemm(a, w, out, bias=bias, act=act, residual=r)
            outs.append(out)
        lib.splitk_enabled = False
        plain = torch.empty(M, N, device=cuda_dev, dtype=out_dtype)
        lib.gemm(a, w, plain, bias=bias, act=act, residual=r)
    finally:
        lib.splitk_enabled = was
    torch.cuda.synchronize()
    assert torch.isfinite(outs[0].float()).all()
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])
    tol = 1e-2 if out_dtype == torch.bfloat16 else 2e-5
    assert _rel(outs[0], ref) < tol
    assert (outs[0].float() - ref).abs().max().item() < (0.15 if out_dtype == torch.bfloat16 else 1e-3)
    # against the unsplit kernel: only the fp32 summation order of the tail tiles differs
    d = (outs[0].float() - plain.float()).abs().max().item()
    assert d < (0.07 if out_dtype == torch.bfloat16 else 2e-4), d
    if out_dtype == torch.float32:
        assert d > 0, "the split path did not run"  # (the fp32 summation order differs somewhere in the tail tiles)
    ws = lib._SPLITK_WS[cuda_dev.index]
    assert int(ws[:4096].view(torch.int32).abs().sum()) == 0  # the counters are back to zero


@pytest.mark.parametrize("split", [False, True])
def test_gemm_grouped_conv_one_launch(cuda_dev, split):
    """HubertPositionalConvEmbedding-shaped grouped Conv1d (groups of 64 channels, zero padding k/2, + GELU + residual) as ONE
    grouped launch (fdm_gemm_args.a_group_cols) against torch's grouped conv1d; plain bf16 and split-bf16 operands."""
    from fdm_b200 import lib
    import torch.nn.functional as F
    B, T, G, cg, k = 3, 70, 5, 64, 16
    Cn = G * cg
    g = torch.Generator(device="cpu").manual_seed(9)
    x = torch.randn(B, T, Cn, generator=g).to(cuda_dev)
    wt = (torch.randn(Cn, cg, k, generator=g) / math.sqrt(k * cg)).to(cuda_dev)
    bias = torch.randn(Cn, generator=g).to(cuda_dev)
    if not split:
        x, wt = x.bfloat16().float(), wt.bfloat16().float()
    pad = k // 2
    conv = F.conv1d(x.double().transpose(1, 2), wt.double(), bias.double(), padding=pad, groups=G)[:, :, :T].transpose(1, 2)
    ref = (F.gelu(conv) + x.double()).float()
    Tp = (T + 2 * pad + 7) // 8 * 8
    dt = torch.float32 if split else torch.bfloat16
    xpad = torch.zeros(B * Tp, Cn, device=cuda_dev, dtype=dt)
    lib.pad_time(x.to(dt).contiguous(), xpad, B, T, Cn, pad, Tp - T - pad, 0)
    wk = wt.permute(0, 2, 1).reshape(Cn, k * cg).contiguous().to(dt)  # [Cout, tap, cin of the group]
    out = torch.zeros(B * Tp, Cn, device=cuda_dev, dtype=dt)
    Mp = B * Tp - (k - 1)
    a_op, w_op = (lib.split(xpad), lib.split(wk)) if split else (xpad, wk)
    lib.gemm(a_op, w_op, out, bias=bias, act=lib.ACT_GELU_ERF, residual=xpad[pad:], M=Mp, lda=Cn, a_rows=B * Tp, taps=k, tap_k=cg,
             tap_row_shift=1, a_group_cols=cg)
    torch.cuda.synchronize()
    got = out.view(B, Tp, Cn)[:, :T].float()
    assert _rel(got, ref) < (3e-5 if split else 4e-3), _rel(got, ref)   # bf16 output rounding: 2^-9 per element
    # and it equals the per-group launches it replaces (same kernel, same k order)
    out_g = torch.zeros_like(out)
    for gi in range(G):
        a_g = lib.Split(a_op.hi[:, gi * cg:], a_op.lo[:, gi * cg:]) if split else xpad[:, gi * cg:]
        w_g = lib.Split(w_op.hi[gi * cg:(gi + 1) * cg], w_op.lo[gi * cg:(gi + 1) * cg]) if split else wk[gi * cg:(gi + 1) * cg]
        lib.gemm(a_g, w_g, out_g[:, gi * cg:(gi + 1) * cg], bias=bias[gi * cg:(gi + 1) * cg].contiguous(), act=lib.ACT_GELU_ERF,
                 residual=xpad[pad:, gi * cg:(gi + 1) * cg], M=Mp, lda=Cn, a_rows=B * Tp, taps=k, tap_k=cg, tap_row_shift=1)
    torch.cuda.synchronize()
    assert torch.equal(out_g[:Mp], out[:Mp])


def test_gemm_gelu_erf_polynomial_epilogue(cuda_dev):
    """The MUFU-free erf-GELU of plain bf16-output GEMMs (act_gelu_erf_poly): |error| < 1.3e-4 absolute against the exact
    GELU over the whole input range, i.e. invisible after the bf16 rounding of the output; fp32 outputs keep the 1.5e-7 form."""
    from fdm_b200 import lib
    import torch.nn.functional as F
    M, N, K = 256, 128, 64
    xs = torch.linspace(-12, 12, M * N).view(M, N)
    a = torch.zeros(M, K, device=cuda_dev, dtype=torch.bfloat16)
    w = torch.zeros(N, K, device=cuda_dev, dtype=torch.bfloat16)
    # W = two stacked 64 x 64 identities: out[:, j] = act(A[:, j mod 64]), the accumulator holds the bf16 inputs exactly
    w[:64] = torch.eye(64, device=cuda_dev, dtype=torch.bfloat16)
    w[64:] = torch.eye(64, device=cuda_dev, dtype=torch.bfloat16)
    xa = xs[:, :64].to(cuda_dev).bfloat16()
    a.copy_(xa)
    ref = F.gelu(xa.double())
    ref = torch.cat([ref, ref], 1)
    out_bf = torch.empty(M, N, device=cuda_dev, dtype=torch.bfloat16)
    out_f = torch.empty(M, N, device=cuda_dev)
    lib.gemm(a, w, out_bf, act=lib.ACT_GELU_ERF)
    lib.gemm(a, w, out_f, act=lib.ACT_GELU_ERF)
    torch.cuda.synchronize()
    assert (out_f.double() - ref).abs().max().item() < 2e-6
    err = (out_bf.double() - ref).abs()
    assert (err <= 1.3e-4 + ref.abs() * 2.0 ** -8).all(), err.max().item()
    big = torch.cat([xa.double().abs() > 4.1] * 2, 1)  # beyond the clamp: x (1 + O(1e-6)) or x O(1e-6)
    assert (err[big] <= 2e-5 * ref[big].abs().clamp_min(1.0) * 12 + ref[big].abs() * 2.0 ** -8).all()


@pytest.mark.parametrize("M,d,big", [(600, 1024, False), (333, 512, False), (25344, 1024, True)])
def test_gemm_layernorm_folding(cuda_dev, M, d, big):
    """LayerNorm folded into the neighbouring GEMMs (fdm_gemm_args a_ln / res_ln / stats_out + fdm_ln_stats_finalize):
    producer statistics, the consumer computed from un-normalised rows with pre-scaled weights, and the residual
    rebuilt element-wise, each against the plain fp32 statement LN(u) on the same bf16 operands."""
    from fdm_b200 import lib
    g = torch.Generator(device="cpu").manual_seed(M + d)
    dev = cuda_dev
    K0, N2 = 2 * d, 3 * d
    A0 = (torch.randn(M, K0, generator=g) * 0.5).to(dev).bfloat16()
    W0 = (torch.randn(d, K0, generator=g) / K0 ** 0.5).to(dev).bfloat16()
    b0 = (0.1 * torch.randn(d, generator=g)).to(dev)
    R0 = (torch.randn(M, d, generator=g) + 0.3).to(dev).bfloat16()  # non-zero row mean
    gamma = (1 + 0.1 * torch.randn(d, generator=g)).to(dev)
    beta = (0.1 * torch.randn(d, generator=g)).to(dev)
    # producer: u = A0 W0^T + b0 + R0 (bf16) with per-row partial statistics
    parts = d // 64
    u = torch.empty(M, d, device=dev, dtype=torch.bfloat16)
    stats = torch.zeros(M, parts, 2, device=dev)
    lib.gemm(A0, W0, u, bias=b0, residual=R0, stats_out=stats)
    u32 = A0.float() @ W0.float().t() + b0 + R0.float()
    assert (u.float() - u32).abs().max().item() < 0.05
    ref_parts = torch.stack([u32.view(M, parts, 64).sum(-1), (u32 ** 2).view(M, parts, 64).sum(-1)], dim=-1)
    assert (stats - ref_parts).abs().max().item() < 2e-3 * ref_parts.abs().max().item()
    mr = torch.empty(M, 2, device=dev)
    lib.ln_stats_finalize(stats, M, parts, d, mr)
    mean, var = u32.mean(-1), u32.var(-1, unbiased=False)
    assert (mr[:, 0] - mean).abs().max().item() < 1e-4
    assert ((mr[:, 1] - (var + 1e-5).rsqrt()) / mr[:, 1]).abs().max().item() < 1e-3
    y = torch.nn.functional.layer_norm(u.float(), (d,), gamma, beta, 1e-5)  # what the LayerNorm kernel would produce (fp32)
    # consumer: y W2^T + b2 from u with W' = W2 diag(gamma)
    W2 = (torch.randn(N2, d, generator=g) / d ** 0.5).to(dev)
    b2 = (0.1 * torch.randn(N2, generator=g)).to(dev)
    Wp = (W2 * gamma[None]).bfloat16()
    colsum = Wp.float().sum(-1).contiguous()
    bias_p = (b2 + W2 @ beta).contiguous()
    out = torch.empty(M, N2, device=dev, dtype=torch.bfloat16)
    lib.gemm(u, Wp, out, bias=bias_p, a_ln=mr, w_colsum=colsum)
    ref = y @ W2.t() + b2
    rel = ((out.float() - ref).norm() / ref.norm()).item()
    assert rel < 1e-2, rel
    # residual: A3 W3^T + b3 + y with y rebuilt from u inside the epilogue
    A3 = (torch.randn(M, d, generator=g)).to(dev).bfloat16()
    W3 = (torch.randn(d, d, generator=g) / d ** 0.5).to(dev).bfloat16()
    out3 = torch.empty(M, d, device=dev, dtype=torch.bfloat16)
    stats3 = torch.zeros(M, parts, 2, device=dev)
    lib.gemm(A3, W3, out3, bias=b0, residual=u, res_ln=mr, res_gamma=gamma, res_beta=beta, stats_out=stats3)
    ref3 = A3.float() @ W3.float().t() + b0 + y
    assert ((out3.float() - ref3).norm() / ref3.norm()).item() < 5e-3
    ref_parts3 = torch.stack([ref3.view(M, parts, 64).sum(-1), (ref3 ** 2).view(M, parts, 64).sum(-1)], dim=-1)
    assert (stats3 - ref_parts3).abs().max().item() < 5e-3 * ref_parts3.abs().max().item()
    torch.cuda.synchronize()


@pytest.mark.parametrize("M,N,K", [(198, 1024, 1024), (77, 130, 50), (512, 15069, 64), (300, 64, 2048)])
def test_gemm_f32(cuda_dev, M, N, K):
    from fdm_b200 import lib
    g = torch.Generator(device="cpu").manual_seed(M + N + K)
    a = torch.randn(M, K, generator=g).to(cuda_dev)
    w = (torch.randn(N, K, generator=g) / math.sqrt(K)).to(cuda_dev)
    bias = torch.randn(N, generator=g).to(cuda_dev)
    out = torch.empty(M, N, device=cuda_dev)
    lib.gemm(a, w, out, bias=bias, act=lib.ACT_MISH)
    torch.cuda.synchronize()
    ref = torch.nn.functional.mish((a.double() @ w.double().t() + bias.double())).float()
    assert _rel(out, ref) < 2e-6


@pytest.mark.parametrize("d", [512, 1024, 64])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_layernorm_fused(cuda_dev, d, dtype):
    from fdm_b200 import lib
    import torch.nn.functional as F
    rows = 333
    g = torch.Generator(device="cpu").manual_seed(d)
    mk = lambda *s: torch.randn(*s, generator=g).to(cuda_dev)
    x, r1, r2 = mk(rows, d).to(dtype), mk(rows, d).to(dtype), mk(rows, d)
    g1, b1, g2, b2 = mk(d), mk(d), mk(d), mk(d)
    vec = mk(10, d)
    idx = torch.tensor([7], dtype=torch.int32, device=cuda_dev)
    out = torch.empty(rows, d, device=cuda_dev, dtype=dtype)
    out2 = torch.empty(rows, d, device=cuda_dev, dtype=torch.float32 if dtype == torch.bfloat16 else torch.bfloat16)
    lib.layernorm(x, out, g1=g1, b1=b1, r1=r1, r2=r2, vec2=vec, vec_index_dev=idx, g2=g2, b2=b2, out2=out2)
    torch.cuda.synchronize()
    y = F.layer_norm(x.float() + r1.float(), (d,), g1, b1)
    ref = F.layer_norm(y + r2 + vec[7], (d,), g2, b2)
    tol = 1e-5 if dtype == torch.float32 else 5e-3
    assert _rel(out, ref) < tol
    assert _rel(out2, ref) < 5e-3
    # single LN + GELU (HuBERT conv stack)
    o3 = torch.empty(rows, d, device=cuda_dev, dtype=dtype)
    lib.layernorm(x, o3, g1=g1, b1=b1, act1=lib.ACT_GELU_ERF)
    assert _rel(o3, F.gelu(F.layer_norm(x.float(), (d,), g1, b1))) < tol


@pytest.mark.gpu
@pytest.mark.parametrize("d", [512, 1024])
@pytest.mark.parametrize("rows", [7, 64, 333, 2 * 198 * 3])
@pytest.mark.parametrize("offset", [0.0, 30.0])
def test_layernorm_hot_modes(cuda_dev, d, rows, offset):
    """The denoiser step's two bf16 LayerNorm launches (column-owner kernel): plain norm3 and the fused pair
    norm1 -> + cross cache (shared by the two guidance passes: r2 has rows / 2 rows) + time row -> norm2; ragged row counts,
    and inputs with a mean of 30 standard deviations (the (mean, M2) merge must not cancel)."""
    from fdm_b200 import lib
    import torch.nn.functional as F
    g = torch.Generator(device="cpu").manual_seed(d + rows)
    mk = lambda *s: torch.randn(*s, generator=g).to(cuda_dev)
    x = (mk(rows, d) + offset).bfloat16()
    half = rows // 2 if rows % 2 == 0 else rows
    r2 = mk(half, d).bfloat16()
    g1, b1, g2, b2 = 1 + 0.1 * mk(d), mk(d), 1 + 0.1 * mk(d), mk(d)
    vec = mk(10, d)
    idx = torch.tensor([3], dtype=torch.int32, device=cuda_dev)
    out = torch.full((rows + 1, d), 7.0, device=cuda_dev, dtype=torch.bfloat16)
    lib.layernorm(x, out[:rows], g1=g1, b1=b1)
    ref = F.layer_norm(x.float(), (d,), g1, b1)
    assert (out[:rows].float() - ref).abs().max() < 2e-2 * max(1.0, float(ref.abs().max()))
    assert _rel(out[:rows], ref) < 4e-3
    assert float(out[rows].float().min()) == 7.0  # nothing written past the last row
    out.fill_(7.0)
    lib.layernorm(x, out[:rows], g1=g1, b1=b1, r2=r2, vec2=vec, vec_index_dev=idx, g2=g2, b2=b2)
    ref2 = F.layer_norm(ref + r2.float().repeat(rows // half, 1) + vec[3], (d,), g2, b2)
    assert _rel(out[:rows], ref2) < 4e-3
    assert float(out[rows].float().min()) == 7.0
    # residual added inside the kernel (x + r1 in fp32), written in place over the residual rows
    res = mk(rows + 1, d).bfloat16()
    res[rows] = 7.0
    refr = F.layer_norm(x.float() + res[:rows].float(), (d,), g1, b1)
    buf = res.clone()
    lib.layernorm(x, buf[:rows], r1=buf[:rows], g1=g1, b1=b1)
    assert _rel(buf[:rows], refr) < 4e-3
    assert float(buf[rows].float().min()) == 7.0
    buf = res.clone()
    lib.layernorm(x, buf[:rows], r1=buf[:rows], g1=g1, b1=b1, r2=r2, vec2=vec, vec_index_dev=idx, g2=g2, b2=b2)
    refr2 = F.layer_norm(refr + r2.float().repeat(rows // half, 1) + vec[3], (d,), g2, b2)
    assert _rel(buf[:rows], refr2) < 4e-3
    assert float(buf[rows].float().min()) == 7.0


def _alibi_mask(H, T, period):
    # closed form of models/fdm_vocaset.py:94-115
    slopes = torch.tensor([2.0 ** (-(2.0 ** -(math.log2(H) - 3)) * (i + 1)) for i in range(H)])
    i = torch.arange(T)[:, None]
    j = torch.arange(T)[None, :]
    bias = -((i - j) // period).float()
    m = slopes[:, None, None] * bias[None]
    m = m.masked_fill((j > i)[None], float("-inf"))
    return slopes, m


@pytest.mark.parametrize("H,dh,T,causal", [(8, 128, 198, True), (4, 128, 99, True), (4, 256, 149, True),
                                            (8, 128, 70, False), (16, 64, 198, False),
                                            # tcgen05 kernel edge shapes: one / two query tiles, ragged and full last key chunk,
                                            # the unmasked EVQ-VAE decoder attention at 4 s / 8 s (T = 198 / 199)
                                            (8, 128, 128, True), (8, 128, 129, True), (8, 128, 208, True), (4, 128, 16, True),
                                            (4, 128, 161, True), (8, 128, 198, False), (8, 128, 199, False), (8, 128, 208, False),
                                            (8, 128, 128, False), (8, 128, 144, False)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_attention(cuda_dev, H, dh, T, causal, dtype):
    from fdm_b200 import lib
    B, d = 3, H * dh
    t_stride = T + 5
    g = torch.Generator(device="cpu").manual_seed(T)
    qkv = torch.randn(B, t_stride, 3 * d, generator=g).to(cuda_dev).to(dtype)
    out = torch.zeros(B * t_stride, d, device=cuda_dev, dtype=dtype)
    rows = qkv.view(B * t_stride, 3 * d)
    slopes = None
    scale = 1.0 / math.sqrt(dh)
    bias = torch.zeros(H, T, T)
    if causal:
        slopes, bias = _alibi_mask(H, T, 30)
        slopes = slopes.to(cuda_dev)
    lib.self_attention(rows[:, 0:], rows[:, d:], rows[:, 2 * d:], out, B, T, t_stride, H, dh, scale, slopes=slopes, period=30)
    torch.cuda.synchronize()
    q, k, v = [t.float().view(B, t_stride, H, dh)[:, :T].permute(0, 2, 1, 3) for t in qkv.split(d, dim=-1)]
    s = (q @ k.transpose(-1, -2)) * scale + bias.to(cuda_dev)
    ref = (s.softmax(-1) @ v).permute(0, 2, 1, 3).reshape(B, T, d)
    got = out.view(B, t_stride, d)[:, :T]
    assert _rel(got, ref) < (2e-5 if dtype == torch.float32 else 1e-2)


@pytest.mark.parametrize("H,dh,T,causal,B", [(4, 256, 149, True, 64), (4, 256, 192, True, 48), (4, 256, 298, True, 30),
                                             (16, 64, 498, False, 12), (12, 64, 299, False, 16), (16, 64, 199, False, 20),
                                             (8, 128, 498, True, 24), (8, 128, 512, True, 20), (8, 128, 300, False, 24),
                                             (8, 128, 498, False, 20), (2, 64, 65, True, 100), (1, 64, 1, False, 200),
                                             (3, 128, 64, True, 70), (4, 256, 33, False, 50)])
def test_attention_tcgen05_general(cuda_dev, H, dh, T, causal, B):
    """attention_tc3.cu: head dim 64 / 128 / 256, key blocks of 64 with the exact two-pass softmax. More (sequence, head,
    query tile) items than SMs, so every persistent CTA runs several items back to back (ring / barrier parities across
    items), ragged last key block and query tile, T = 1, the BIWI (4 x 256, T = 149), HuBERT (16 x 64, N = 498), wav2vec2
    (12 x 64, N = 299) and 10 s VOCASET (8 x 128, T = 498) shapes. Rows between the sequences must stay untouched."""
    from fdm_b200 import lib
    d = H * dh
    t_stride = T + 3
    g = torch.Generator(device="cpu").manual_seed(T + dh)
    qkv = torch.randn(B, t_stride, 3 * d, generator=g).to(cuda_dev).bfloat16()
    out = torch.full((B * t_stride, d), 3.0, device=cuda_dev, dtype=torch.bfloat16)
    rows = qkv.view(B * t_stride, 3 * d)
    slopes = None
    scale = 1.0 / math.sqrt(dh)
    bias = torch.zeros(H, T, T)
    if causal:
        slopes, bias = _alibi_mask(H, T, 25)
        slopes = slopes.to(cuda_dev)
    lib.self_attention(rows[:, 0:], rows[:, d:], rows[:, 2 * d:], out, B, T, t_stride, H, dh, scale, slopes=slopes, period=25)
    torch.cuda.synchronize()
    q, k, v = [t.float().view(B, t_stride, H, dh)[:, :T].permute(0, 2, 1, 3) for t in qkv.split(d, dim=-1)]
    s = (q @ k.transpose(-1, -2)) * scale + bias.to(cuda_dev)
    ref = (s.softmax(-1) @ v).permute(0, 2, 1, 3).reshape(B, T, d)
    got = out.view(B, t_stride, d)
    assert torch.isfinite(got.float()).all()
    assert _rel(got[:, :T], ref) < 1e-2
    assert (got[:, :T].float() - ref).abs().max() < 0.06
    assert bool((got[:, T:] == 3.0).all())  # padding rows between the sequences are not written


def test_ddpm_step_bit_exact(cuda_dev):
    from fdm_b200 import lib
    B, n = 4, 3168 * 64
    g = torch.Generator(device="cpu").manual_seed(0)
    mk = lambda: torch.randn(B, n, generator=g).to(cuda_dev)
    c, u, xt, nz = mk(), mk(), mk(), mk()
    tabs = [torch.rand(1000, generator=g).to(cuda_dev) for _ in range(3)]
    t = torch.tensor([999, 500, 1, 0], device=cuda_dev)
    out = torch.empty_like(xt)
    ob = torch.empty(B, n, device=cuda_dev, dtype=torch.bfloat16)
    lib.ddpm_step(c, xt, out, *tabs, x0_uncond=u, guidance=2.5, noise=nz, out_bf16=ob, t_per_clip=t)
    torch.cuda.synchronize()
    x0 = u + 2.5 * (c - u)
    e = lambda tab: tab[t][:, None]
    ref = e(tabs[0]) * x0 + e(tabs[1]) * xt + e(tabs[2]) * nz * (t > 0)[:, None]
    assert torch.equal(out, ref)
    assert torch.equal(ob, ref.bfloat16())
    # graph-style addressing: t from a device-resident step counter
    sched = torch.arange(999, -1, -1, dtype=torch.int32, device=cuda_dev)
    cur = torch.tensor([499], dtype=torch.int32, device=cuda_dev)
    t_dev = sched[499:500].clone()
    lib.ddpm_step(c, xt, out, *tabs, noise=nz, t_dev=t_dev)
    lib.advance_cursor(cur, sched, t_dev)
    torch.cuda.synchronize()
    assert cur.item() == 500 and t_dev.item() == 499
    assert torch.equal(out, tabs[0][500] * c + tabs[1][500] * xt + tabs[2][500] * nz)


def test_philox_normal_stats(cuda_dev):
    from fdm_b200 import lib
    out = torch.empty(8, 1 << 18, device=cuda_dev)
    lib.philox_normal(out, seed=1234, clip_index0=3, t=17)
    o2 = torch.empty(4, 1 << 18, device=cuda_dev)
    lib.philox_normal(o2, seed=1234, clip_index0=5, t=17)
    torch.cuda.synchronize()
    assert abs(out.mean().item()) < 5e-3 and abs(out.std().item() - 1) < 5e-3
    assert torch.equal(out[2:6], o2)  # keyed by global clip index: invariant to sharding
    assert not torch.equal(out[0], out[1])


@pytest.mark.parametrize("D,L", [(64, 3168), (128, 1192), (64, 100)])
def test_vq_quantize(cuda_dev, D, L):
    from fdm_b200 import lib
    B, n = 3, 256
    g = torch.Generator(device="cpu").manual_seed(D + L)
    z = torch.randn(B, L, D, generator=g).to(cuda_dev)
    cb = torch.randn(7 * n, D, generator=g).to(cuda_dev)
    off = torch.tensor([0, 256 * 3, 256 * 6], device=cuda_dev)
    idx, zq, zr = lib.vq_quantize(z, cb, n, code_offset=off, want_rows=True)
    torch.cuda.synchronize()
    for b in range(B):
        e = cb[off[b]:off[b] + n].double()
        d = (z[b].double() ** 2).sum(1, keepdim=True) + (e ** 2).sum(1) - 2 * z[b].double() @ e.t()
        ref = d.argmin(1)
        got = idx.view(B, L)[b]
        # fp32 vs fp64 argmin may differ only on near-ties
        bad = (got != ref)
        if bad.any():
            dd = d[bad]
            gap = (dd.gather(1, got[bad][:, None]) - dd.min(1, keepdim=True).values).abs().max().item()
            assert gap < 1e-4, gap
        assert torch.equal(zr[b], cb[off[b] + got])
        assert torch.equal(zq[b], zr[b].t())
    assert idx.min() >= 0 and idx.max() < n


def _vq_case(name, gen, D=64):
    """Synthetic (z (B, L, D), codebook (n_slices*n, D), n, offsets) cases for the tensor-core VQ path."""
    if name == "normal_mead":  # per-clip emotion slices, ragged last tile
        B, L, n = 5, 792, 256
        return torch.randn(B, L, D, generator=gen), torch.randn(7 * n, D, generator=gen), n, [0, 768, 1536, 1536, 256]
    if name == "reference_init":  # uniform(+-1/n) codes against unit-variance latents: all distances nearly equal
        B, L, n = 2, 3168, 256
        return torch.randn(B, L, D, generator=gen), (torch.rand(n, D, generator=gen) * 2 - 1) / n, n, None
    if name == "ragged":
        return torch.randn(3, 129, D, generator=gen), torch.randn(128, D, generator=gen), 128, None
    if name == "single_row":
        return torch.randn(4, 1, D, generator=gen), torch.randn(64, D, generator=gen), 64, None
    if name == "ties":  # duplicate and nearly-duplicate codes, latents on codes / midpoints / zero, wide dynamic range
        B, L, n = 2, 1000, 256
        cb = torch.randn(n, D, generator=gen)
        cb[100] = cb[7]                       # exact duplicate: lowest index must win
        cb[200] = cb[7]
        cb[150] = cb[31] * (1 + 1e-7)         # one-ulp-ish neighbours
        cb[151] = cb[31] + 1e-6 * torch.randn(D, generator=gen)
        z = torch.randn(B, L, D, generator=gen)
        pick = torch.randint(0, n, (B, L), generator=gen)
        z[:, 0:200] = cb[pick[:, 0:200]]
        other = torch.randint(0, n, (B, L), generator=gen)
        z[:, 200:400] = 0.5 * (cb[pick[:, 200:400]] + cb[other[:, 200:400]])
        z[:, 400:410] = 0.0
        z[:, 410:500] = cb[7] + 1e-6 * torch.randn(B, 90, D, generator=gen)
        z[1] *= 1e3
        z[1, 700:] *= 1e-9
        return z, cb, n, None
    raise KeyError(name)


@pytest.mark.parametrize("D", [64, 128])
@pytest.mark.parametrize("case", ["normal_mead", "reference_init", "ragged", "single_row", "ties"])
def test_vq_tensor_path_bit_exact(cuda_dev, case, D):
    """The tcgen05 filter + exact recheck must return the oracle's indices on every row (not just away from ties), for
    both latent widths of the reference (64: VOCASET / MEAD, 128: BIWI - two K-halves per tile)."""
    from fdm_b200 import lib
    from oracle import reference_ops as R  # checker
    gen = torch.Generator(device="cpu").manual_seed(sum(map(ord, case)) + D)
    z, cb, n, offs = _vq_case(case, gen, D)
    B, L, D = z.shape
    off = torch.tensor(offs, device=cuda_dev) if offs is not None else None
    cnt = torch.zeros(1, dtype=torch.int64, device=cuda_dev)
    idx, zq, zr = lib.vq_quantize(z.to(cuda_dev), cb.to(cuda_dev), n, code_offset=off, want_rows=True, algo=lib.VQ_TENSOR,
                                  recheck_rows=cnt)
    idx_f, _, _ = lib.vq_quantize(z.to(cuda_dev), cb.to(cuda_dev), n, code_offset=off, want_bdl=False, algo=lib.VQ_FFMA)
    torch.cuda.synchronize()
    idx, idx_f = idx.view(B, L).cpu(), idx_f.view(B, L).cpu()
    for b in range(B):
        o = offs[b] if offs is not None else 0
        ref, _, _ = R.vq_quantize(z[b], cb[o:o + n])
        assert torch.equal(idx[b], ref), (case, b, (idx[b] != ref).nonzero()[:5].tolist())
        assert torch.equal(idx_f[b], ref)
        assert torch.equal(zr[b].cpu(), cb[o + ref])
        assert torch.equal(zq[b].cpu(), cb[o + ref].t())
    frac = cnt.item() / (B * L)
    if case in ("normal_mead", "ragged"):
        assert frac < 0.02, frac       # random data: the exact pass is the rare path
    if case == "ties":
        assert cnt.item() >= 2 * 200   # the rows sitting on duplicated codes must have gone through it


@pytest.mark.parametrize("D", [64, 128])
def test_vq_nan_inf_rows_are_defined(cuda_dev, D):
    """A diverged sample (NaN / +-Inf latent rows) must not fault: no distance compares below +inf, so such a row takes
    code 0 on BOTH kernels (torch.argmin never crashes either; oracle/vq_ref.c starts from index 0), and every other
    row keeps the oracle's index (the FFMA kernel indexed shared memory with 0x7fffffff)."""
    from fdm_b200 import lib
    from oracle import reference_ops as R  # checker
    gen = torch.Generator(device="cpu").manual_seed(11 + D)
    B, L, n = 2, 300, 256
    z = torch.randn(B, L, D, generator=gen)
    cb = torch.randn(n, D, generator=gen)
    bad = [3, 64, 65, 127, 128, 299]
    z[0, bad[0]] = float("nan")
    z[0, bad[1], 5] = float("inf")
    z[0, bad[2], 0] = float("-inf")
    z[1, bad[3]] = float("inf")
    z[1, bad[4], D - 1] = float("nan")
    z[1, bad[5]] = float("nan")
    algos = [lib.VQ_FFMA, lib.VQ_TENSOR]
    for algo in algos:
        idx, zq, zr = lib.vq_quantize(z.to(cuda_dev), cb.to(cuda_dev), n, want_rows=True, algo=algo)
        torch.cuda.synchronize()  # an out-of-bounds shared-memory read would surface here as a sticky fault
        idx = idx.view(B, L).cpu()
        assert int(idx.min()) >= 0 and int(idx.max()) < n
        for b in range(B):
            ref, _, _ = R.vq_quantize(z[b], cb)
            assert torch.equal(idx[b], ref), (algo, b, (idx[b] != ref).nonzero()[:5].tolist())
        for b, r in ((0, bad[0]), (0, bad[1]), (0, bad[2]), (1, bad[3]), (1, bad[4]), (1, bad[5])):
            assert idx[b, r] == 0
            assert torch.equal(zr[b, r].cpu(), cb[0])


@pytest.mark.parametrize("D", [64, 128])
def test_vq_tensor_dot_error_bound(cuda_dev, D):
    """Measures the error of the tensor-core scores a_j = z.e_j - ee_j/2 (bf16x3 products + a bf16x3 image of -ee_j/2,
    fp32 accumulation in TMEM) against fp64 and checks it sits well inside the budget the kernel's candidate window
    assumes (vq_tc_impl.cuh header: 2^-15 |z||e| (1.5 x 2^-15 at D = 128) for the dot product, 2^-23 zz + 2^-21 ee for
    the roundings): half of the budget or less."""
    from fdm_b200 import lib
    gen = torch.Generator(device="cpu").manual_seed(5)
    B, L, n = 2, 2048, 256
    z = torch.randn(B, L, D, generator=gen) * torch.logspace(-3, 3, L)[None, :, None]
    cb = torch.randn(n, D, generator=gen) * torch.logspace(-2, 2, n)[:, None]
    acc = torch.empty(B * L, n, device=cuda_dev)
    lib.vq_quantize(z.to(cuda_dev), cb.to(cuda_dev), n, want_bdl=False, algo=lib.VQ_TENSOR, dbg_acc=acc)
    torch.cuda.synchronize()
    zd, ed = z.view(-1, D).double(), cb.double()
    ee32 = torch.zeros(n)
    for k in range(D):  # the fp32 chain the kernel and the oracle use
        ee32 = torch.addcmul(ee32, cb[:, k], cb[:, k])
    exact = zd @ ed.t() - 0.5 * ee32.double()[None]
    err = (acc.cpu().double() - exact).abs()
    budget_dot = zd.norm(dim=1, keepdim=True) * ed.norm(dim=1)[None]
    budget_rnd = (zd ** 2).sum(1, keepdim=True) + ee32.double()[None]
    c_dot = 2.0 ** -16 * (1.0 if D == 64 else 1.5)
    assert (err <= c_dot * budget_dot + 2.0 ** -22 * ee32.double()[None]).all(), (err / budget_dot).max().item()


def test_misc_kernels(cuda_dev):
    from fdm_b200 import lib
    import torch.nn.functional as F
    g = torch.Generator(device="cpu").manual_seed(1)
    src = torch.randn(2, 70, 45, generator=g).to(cuda_dev)
    dst = torch.empty(2, 45, 70, device=cuda_dev, dtype=torch.bfloat16)
    lib.transpose_bcl_to_blc(src, dst)
    assert torch.equal(dst, src.transpose(1, 2).bfloat16())
    # LeakyReLU + InstanceNorm over time
    x = torch.randn(3, 50, 96, generator=g).to(cuda_dev)
    o = torch.empty_like(x)
    lib.leaky_instnorm(x, o, 3, 50, 50, 96)
    ref = F.instance_norm(F.leaky_relu(x, 0.2).transpose(1, 2)).transpose(1, 2)
    assert _rel(o, ref) < 1e-5
    # HuBERT conv0 + LN + GELU
    audio = torch.randn(2, 4000, generator=g).to(cuda_dev)
    w, b = torch.randn(512, 1, 10, generator=g).to(cuda_dev), torch.randn(512, generator=g).to(cuda_dev)
    lg, lb = torch.randn(512, generator=g).to(cuda_dev), torch.randn(512, generator=g).to(cuda_dev)
    Lout = (4000 - 10) // 5 + 1
    out = torch.empty(2, 800, 512, device=cuda_dev)
    lib.hubert_conv0(audio, w.view(512, 10).contiguous(), b, lg, lb, out, Lout, 800, 512)
    ref = F.gelu(F.layer_norm(F.conv1d(audio[:, None], w, b, stride=5).transpose(1, 2), (512,), lg, lb))
    torch.cuda.synchronize()
    assert _rel(out[:, :Lout], ref) < 1e-5
    assert (out[:, Lout:] == 0).all()
    # bf16 output (polynomial erf-GELU), a per-clip stride that is not a multiple of the 128 frames a CTA covers, and the raw
    # variant without LayerNorm (wav2vec2-base: GroupNorm over time follows in another kernel)
    for stride in (800, 803):
        ob = torch.full((2, stride, 512), float("nan"), device=cuda_dev, dtype=torch.bfloat16)
        lib.hubert_conv0(audio, w.view(512, 10).contiguous(), b, lg, lb, ob, Lout, stride, 512)
        torch.cuda.synchronize()
        assert _rel(ob[:, :Lout], ref) < 4e-3
        assert (ob[:, :Lout].float() - ref).abs().max().item() < 2e-4 + 2.0 ** -8 * ref.abs().max().item()
        assert (ob[:, Lout:] == 0).all()
    raw = torch.full((2, 800, 512), float("nan"), device=cuda_dev)
    lib.hubert_conv0(audio, w.view(512, 10).contiguous(), None, None, None, raw, Lout, 800, 512)
    torch.cuda.synchronize()
    assert _rel(raw[:, :Lout], F.conv1d(audio[:, None], w, None, stride=5).transpose(1, 2)) < 1e-5
    assert (raw[:, Lout:] == 0).all()


@pytest.mark.parametrize("orig_sr", [44100, 48000, 22050, 8000])
def test_resample_poly_vs_scipy(cuda_dev, orig_sr):
    """Device polyphase resampler (the demos' librosa.load(sr=16000) step) against
    scipy.signal.resample_poly on the same float64 filter design; two clips of different content in one launch."""
    from scipy import signal
    import numpy as np
    from fdm_b200 import frontend
    rs = np.random.RandomState(orig_sr)
    L = orig_sr // 2 + 37
    x = rs.randn(2, L).astype(np.float32)
    x[1] = np.sin(2 * np.pi * 440.0 * np.arange(L) / orig_sr).astype(np.float32)
    got = frontend.resample(torch.from_numpy(x).to(cuda_dev), orig_sr, 16000)
    taps, up, down, pre = frontend.resample_filter(orig_sr, 16000)
    ref = np.stack([signal.resample_poly(x[i].astype(np.float64), up, down) for i in range(2)])
    assert got.shape == ref.shape
    assert np.abs(got.cpu().numpy() - ref).max() < 2e-5 * max(1.0, np.abs(ref).max())
    # the whole front-end: resample + normalise + 1 s of zeros
    full = frontend.prepare_audio(torch.from_numpy(x).to(cuda_dev), sample_rate=orig_sr)
    assert full.shape == (2, ref.shape[1] + 16000)
    want = (ref - ref.mean(1, keepdims=True)) / np.sqrt(ref.var(1, keepdims=True) + 1e-7)
    assert np.abs(full[:, :ref.shape[1]].cpu().numpy() - want).max() < 1e-4
    assert float(full[:, ref.shape[1]:].abs().max()) == 0.0
