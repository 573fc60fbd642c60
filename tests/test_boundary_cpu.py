"""CPU: the drop-in boundary — state_dict layout, schedule bits, C-ABI symbols, loud failure without a GPU."""
import json
import os
import re

import numpy as np
import pytest
import torch

from helpers import GOLDEN, golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build_meta(preset):
    os.environ["FDM_B200_RANDOM_AUDIO_ENCODER"] = "1"  # no checkpoint here: the architecture is what is under test
    with torch.device("meta"):
        if preset == "vocaset":
            from models.fdm_vocaset import FDM
            from models.vq_vae_vocaset import VQAutoEncoder
            from models.utils.config import vocaset_vq_vae_args as vargs
            from video_diffusion_pytorch.diffusion_BIWI_encoder_decoder import GaussianDiffusion
            fdm = FDM(feature_dim=1024)
        elif preset == "mead":
            from models.fdm_vqvae_mead import FDM
            from models.vq_vae_emotion import VQAutoEncoder
            from utiles.args import vq_vae_args as vargs
            from video_diffusion_pytorch.diffusion_mead_encoder_decoder import GaussianDiffusion
            fdm = FDM(feature_dim=512, vertice_dim=5023 * 3, struct="Dec")
        else:
            from models.fdm import FDM
            from models.vq_vae import VQAutoEncoder
            from models.utils.config import biwi_vq_vae_args as vargs
            from video_diffusion_pytorch.diffusion_BIWI_encoder_decoder import GaussianDiffusion
            fdm = FDM(feature_dim=1024, struct="Dec")
        return GaussianDiffusion(fdm, timesteps=1000, loss_type="l2"), VQAutoEncoder(vargs())


def test_audio_encoder_random_init_is_opt_in(monkeypatch):
    """a missing / mistyped checkpoint path must raise like HF's from_pretrained, not silently build a random
    audio encoder; random initialisation is an explicit request (keyword or FDM_B200_RANDOM_AUDIO_ENCODER=1)."""
    import warnings
    from models.hubert import HubertModel
    monkeypatch.delenv("FDM_B200_RANDOM_AUDIO_ENCODER", raising=False)
    monkeypatch.setenv("HF_HUB_OFFLINE", "1")
    with pytest.raises(Exception):
        HubertModel.from_pretrained("/nonexistent/hubert-large-ls960-ft")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        with torch.device("meta"):
            m = HubertModel.from_pretrained("/nonexistent/hubert-large-ls960-ft", random_init=True)
    assert m.config.hidden_size == 1024 and m.config.num_hidden_layers == 24


@pytest.mark.parametrize("preset", ["vocaset", "mead", "biwi"])
def test_state_dict_layout_matches_reference(preset):
    """Reference checkpoints must load unchanged: identical keys and shapes (golden: the real reference modules)."""
    import warnings
    warnings.simplefilter("ignore")
    with open(os.path.join(GOLDEN, "state_dict_layout.json")) as f:
        layout = json.load(f)[preset]
    diff, ae = _build_meta(preset)
    for mod, want in ((diff, layout["diffusion"]), (ae, layout["vqvae"])):
        got = {k: list(v.shape) for k, v in mod.state_dict().items()}
        assert set(got) == set(want), (sorted(set(want) - set(got))[:5], sorted(set(got) - set(want))[:5])
        for k in want:
            assert got[k] == want[k], k


def test_product_schedule_bit_exact():
    from fdm_b200.modules import cosine_tables
    g = golden("schedule")
    for name, val in cosine_tables(1000):
        assert np.array_equal(val.to(torch.float32).numpy(), g[name]), name


def test_c_abi_exports_every_declared_symbol():
    from fdm_b200 import lib
    hdr = open(os.path.join(ROOT, "include", "fdm_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(fdm_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    handle = lib.load()
    for name in declared:
        assert hasattr(handle, name), f"{name} declared in include/fdm_b200.h but not exported"
    assert declared == set(lib.EXPORTS), declared ^ set(lib.EXPORTS)
    assert handle.fdm_abi_version() == 4


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    from fdm_b200 import lib
    with pytest.raises(lib.FdmError):
        lib.require_device()
    with pytest.raises((lib.FdmError, AssertionError)):
        lib.gemm(torch.zeros(4, 8), torch.zeros(4, 8), torch.zeros(4, 4))


def test_ctypes_structs_mirror_the_header():
    """The ctypes argument structs of fdm_b200/lib.py must list the fields of include/fdm_b200.h's structs in the same
    order with matching widths, and sizeof must agree with what the compiler laid out (a tiny C program prints it)."""
    import ctypes as C
    import subprocess
    import tempfile
    from fdm_b200 import lib
    hdr_path = os.path.join(ROOT, "include", "fdm_b200.h")
    hdr = re.sub(r"/\*.*?\*/", "", open(hdr_path).read(), flags=re.S)
    pairs = {"fdm_gemm_args": lib.GemmArgs, "fdm_norm_args": lib.NormArgs, "fdm_attn_args": lib.AttnArgs,
             "fdm_ddpm_args": lib.DdpmArgs, "fdm_ddim_args": lib.DdimArgs}
    width = {"int64_t": 8, "uint64_t": 8, "int32_t": 4, "float": 4}
    for cname, cls in pairs.items():
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (cname, cname), hdr, flags=re.S).group(1)
        fields = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            m = re.match(r"(const\s+)?(\w+)\s*(\*?)\s*(.*)", decl)
            ctype, star, names = m.group(2), m.group(3), m.group(4)
            for nm in names.split(","):
                nm = nm.strip()
                ptr = bool(star) or nm.startswith("*")
                fields.append((nm.lstrip("* "), 8 if ptr else width[ctype]))
        got = [(n, C.sizeof(t)) for n, t in cls._fields_]
        assert got == fields, (cname, [a for a, b in zip(got, fields) if a != b][:3], len(got), len(fields))
    with tempfile.TemporaryDirectory() as td:
        src = os.path.join(td, "s.c")
        with open(src, "w") as f:
            f.write('#include <stdio.h>\n#include "%s"\nint main(void){printf("%%zu %%zu %%zu %%zu %%zu\\n", sizeof(fdm_gemm_args), '
                    'sizeof(fdm_norm_args), sizeof(fdm_attn_args), sizeof(fdm_ddpm_args), sizeof(fdm_ddim_args));return 0;}\n' % hdr_path)
        exe = os.path.join(td, "s")
        subprocess.check_call(["gcc", "-o", exe, src])
        sizes = [int(x) for x in subprocess.check_output([exe]).split()]
    assert sizes == [C.sizeof(c) for c in pairs.values()], sizes


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the real reference modules from oracle/_ref on the host cores, or the oracle port when
    that copy is absent; no GPU involved) must print one JSON line with the keys the driver reads."""
    import json
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--ref-sample-steps", "1", "--seconds", "1"], capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "frames/s" and line["higher_is_better"] is True
    from oracle import ref_runner
    want_kind = "reference" if ref_runner.available() else "port"  # oracle/_ref: the real reference, copied by build()
    assert line["value"] > 0 and line["cpu_baseline"]["kind"] == want_kind and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["ms_per_step"] > 0 and line["config"]["workload"].startswith("VOCASET LG-LDM sampling, batch 64 clips x 1 s")
    sys.path.insert(0, ROOT)
    import bench
    args = bench.parse.__globals__["argparse"].Namespace(preset="vocaset", clips=64, seconds=1.0, ddpm_steps=1000, no_cfg=False)
    assert line["config"] == bench.workload_config(args, 1)  # the two arms of the driver's ratio describe the same workload


def test_resample_filter_matches_scipy_design():
    """Host-side filter design of the device resampler == scipy.signal.resample_poly's own (firwin, Kaiser beta 5)."""
    import numpy as np
    from scipy import signal
    from fdm_b200.frontend import resample_filter
    for orig in (44100, 48000, 22050, 8000):
        taps, up, down, pre = resample_filter(orig, 16000)
        max_rate = max(up, down)
        half_len = 10 * max_rate
        h = signal.firwin(2 * half_len + 1, 1.0 / max_rate, window=("kaiser", 5.0)) * up
        n_pre_pad = down - half_len % down
        assert np.abs(taps.numpy() - np.concatenate([np.zeros(n_pre_pad), h])).max() < 1e-14
        assert pre == (half_len + n_pre_pad) // down


def test_binding_refuses_a_library_with_another_abi_version(monkeypatch):
    """A stale libfdm_b200.so would read the ctypes argument structs with another layout: load() must refuse it."""
    from fdm_b200 import lib
    monkeypatch.setattr(lib, "_lib", None)
    monkeypatch.setattr(lib, "ABI_VERSION", lib.ABI_VERSION + 1)
    with pytest.raises(lib.FdmError, match="ABI version"):
        lib.load()
    monkeypatch.setattr(lib, "ABI_VERSION", lib.ABI_VERSION - 1)
    assert lib.load().fdm_abi_version() == lib.ABI_VERSION


def test_polynomial_erf_gelu_constants():
    """The MUFU-free erf-GELU of the bf16 epilogues (csrc/common.cuh: act_gelu_erf_poly), restated in float32 numpy from the
    constants in the source: |error| < 1.3e-4 against the exact GELU everywhere, exactly x (1 + O(1e-6)) / x O(1e-6) beyond the clamp."""
    import re
    import numpy as np
    from scipy.special import erf
    src = open(os.path.join(ROOT, "face-diffusion-model_b200", "csrc", "common.cuh")).read()
    body = src[src.index("float act_gelu_erf_poly(float x)"):]
    body = body[:body.index("}") + 1]
    scale = np.float32(re.search(r"fmaf\(x, ([0-9.eE+-]+)f, 0\.5f\)", body).group(1))
    first = re.search(r"fmaf\((-?[0-9.]+)f, s, (-?[0-9.]+)f\)", body)
    chain = [np.float32(first.group(1)), np.float32(first.group(2))] + [np.float32(c) for c in re.findall(r"p = fmaf\(p, s, (-?[0-9.]+)f\)", body)]
    assert len(chain) == 8, chain
    x = np.linspace(-12, 12, 480001).astype(np.float32)
    w = (np.clip(x * scale + np.float32(0.5), 0, 1) - np.float32(0.5)).astype(np.float32)
    s2 = (w * w).astype(np.float32)
    p = np.full_like(s2, chain[0])
    for c in chain[1:]:
        p = (p * s2 + c).astype(np.float32)
    y = (x * (np.float32(0.5) * (w * p).astype(np.float32) + np.float32(0.5))).astype(np.float32)
    ref = 0.5 * x.astype(np.float64) * (1 + erf(x.astype(np.float64) / np.sqrt(2)))
    assert np.abs(y - ref).max() < 1.3e-4
    far = np.abs(x) > 4.1
    assert np.abs(y[far] - ref[far]).max() < 2e-5 * 12
    assert abs(float(scale) - 1 / (2 * 2.85 * np.sqrt(2))) < 1e-6
