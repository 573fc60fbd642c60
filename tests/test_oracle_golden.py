"""CPU: the oracle restatement against the golden vectors produced by the real reference (oracle/gen_golden.py)."""
import numpy as np
import pytest
import torch

from helpers import build_product, golden, hf_audio_model, oracle_inputs

PRESETS = ["vocaset", "mead", "biwi"]


def test_schedule_tables_bit_exact():
    from oracle import reference_ops as R
    g = golden("schedule")
    tabs = R.diffusion_tables(1000)
    assert set(tabs) == set(g.files)
    for k in g.files:
        assert np.array_equal(tabs[k].numpy(), g[k]), k


@pytest.mark.parametrize("preset", PRESETS)
def test_oracle_fdm_and_chain(preset):
    from oracle import reference_ops as R
    from oracle.weights import host_noise
    g = golden(preset)
    fdm, ae, diff = build_product(preset)
    sd, audio, idh, emo = oracle_inputs(preset, fdm)
    hidden = R.audio_encode(hf_audio_model(preset, sd), audio)
    assert np.allclose(hidden.numpy(), g["audio_hidden"], atol=1e-5)
    x = torch.from_numpy(g["x_T"])
    P = R.PRESETS[preset]
    for t in (999, 500, 0):
        y = R.fdm_forward(sd, preset, hidden, t, x, idh, emo)
        assert np.abs(y.numpy() - g[f"x0_t{t}"]).max() < 2e-5
    steps = g["chain_steps"].tolist()
    tabs = R.diffusion_tables(1000)
    out = R.p_sample_loop(tabs, lambda z, t: R.fdm_forward(sd, preset, hidden, t, z, idh, emo), x,
                          lambda t: host_noise(99, 0, t, (1,) + tuple(x.shape))[0], steps=steps)
    assert np.abs(out.numpy() - g["chain_out"]).max() < 2e-5


@pytest.mark.parametrize("preset", ["vocaset", "biwi"])
def test_oracle_ddim(preset):
    from oracle import reference_ops as R
    g = golden(preset)
    fdm, ae, diff = build_product(preset)
    sd, audio, idh, emo = oracle_inputs(preset, fdm)
    hidden = torch.from_numpy(g["audio_hidden"])
    x = torch.from_numpy(g["x_T"])
    out = R.ddim_sample(R.diffusion_tables(1000), lambda z, t: R.fdm_forward(sd, preset, hidden, t, z, idh, emo), x, int(g["ddim_steps"]))
    assert np.abs(out.numpy() - g["ddim_out"]).max() < 2e-5


@pytest.mark.parametrize("preset", PRESETS)
@pytest.mark.parametrize("codebook", ["reference", "normal"])
def test_oracle_vq_and_decode(preset, codebook):
    from oracle import reference_ops as R
    g = golden(preset)
    _, ae, _ = build_product(preset, codebook=codebook)
    P = R.PRESETS[preset]
    sd = {k: v.detach() for k, v in ae.state_dict().items()}
    z = torch.from_numpy(g["chain_out"])
    emo_pos = 4 if P["emotion"] else None
    idx, zq, margin = R.vq_quantize(z, sd["quantize.embedding.weight"], emo_pos)
    ref_idx = g[f"vq_idx_{codebook}"]
    robust = g[f"vq_margin_{codebook}"] > 1e-5
    assert np.array_equal(idx.numpy()[robust], ref_idx[robust])
    assert (idx.numpy() != ref_idx).mean() < 0.01
    verts = R.vq_decode(sd, preset, zq)
    cols = np.arange(0, verts.shape[-1], 16)
    assert np.abs(verts.numpy()[:, cols] - g[f"verts_cols_{codebook}"]).max() < 1e-4


def test_vq_oracle_tie_break_and_order():
    """Lowest index wins exact ties; the distance is the defined fp32 fmaf-chain expression."""
    from oracle import reference_ops as R
    cb = torch.zeros(256, 64)
    cb[5] = 1.0
    cb[9] = 1.0  # duplicate code: exact tie with 5
    z = torch.ones(3, 64)
    idx, zq, margin = R.vq_quantize(z, cb)
    assert idx.tolist() == [5, 5, 5] and float(margin.max()) == 0.0
    rng = np.random.default_rng(0)
    zz = rng.standard_normal((4, 64)).astype(np.float32)
    ee = rng.standard_normal((256, 64)).astype(np.float32)
    idx, _, _ = R.vq_quantize(torch.from_numpy(zz), torch.from_numpy(ee))
    d = []
    for r in range(4):
        row = []
        for j in range(256):
            a = np.float32(0); b = np.float32(0); c = np.float32(0)
            for k in range(64):
                a = np.float32(np.float64(zz[r, k]) * np.float64(zz[r, k]) + np.float64(a))
                b = np.float32(np.float64(ee[j, k]) * np.float64(ee[j, k]) + np.float64(b))
                c = np.float32(np.float64(zz[r, k]) * np.float64(ee[j, k]) + np.float64(c))
            row.append(np.float32(np.float32(a + b) - np.float32(np.float32(2) * c)))
        d.append(int(np.argmin(np.array(row))))
    assert idx.tolist() == d


def test_cfg_formula():
    from oracle import reference_ops as R
    g = golden("cfg")
    c, u = torch.from_numpy(g["cond"])[0], torch.from_numpy(g["uncond"])[0]
    out = R.cfg_forward(lambda oh: c if oh.abs().sum() > 0 else u, torch.ones(1, 3), 2.5)
    assert np.array_equal(out.numpy(), g["out"][0])


def test_philox_reference_statistics():
    from oracle.philox_ref import philox_normal, philox4x32_10
    # known-answer vector of Philox4x32-10 (Random123 kat_vectors: counter = key = 0)
    r = philox4x32_10([0], [0], [0], [0], 0, 0)
    assert [int(v[0]) for v in r] == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    z = philox_normal(1234, 3, 17, 1 << 16)
    assert abs(z.mean()) < 0.02 and abs(z.std() - 1) < 0.02


@pytest.mark.parametrize("preset", PRESETS)
def test_oracle_vq_encode(preset):
    """The encoder restatement against the outputs of the reference's VQAutoEncoder.encode (oracle/gen_golden_encode.py)."""
    import numpy as np
    from helpers import encoder_case
    from oracle import reference_ops as R
    sd, x, emo, want = encoder_case(preset)
    got = R.vq_encode(sd, preset, x, emo)
    assert got.shape == want.shape
    assert (got - want).abs().max().item() < 2e-5 * max(1.0, want.abs().max().item())


def test_audio_frontend_formula_is_the_processors():
    """Pins the statement the device front-end is tested against (tests/test_parity_gpu.py::test_audio_frontend_...):
    (x - mean) / sqrt(var + 1e-7) is what the installed Wav2Vec2 feature extractor - the arithmetic behind the
    Wav2Vec2Processor call of demo/demo_3d_mead.py:85-89 - produces for a mono clip."""
    import numpy as np
    from transformers import Wav2Vec2FeatureExtractor
    rng = np.random.default_rng(0)
    x = (rng.standard_normal(48000) * 0.05 + 0.01).astype(np.float32)
    fe = Wav2Vec2FeatureExtractor(feature_size=1, sampling_rate=16000, padding_value=0.0, do_normalize=True,
                                  return_attention_mask=False)  # wav2vec2-base-960h's preprocessor_config.json
    got = np.squeeze(fe(x, sampling_rate=16000).input_values)
    want = (x - x.mean()) / np.sqrt(x.var() + 1e-7)
    assert got.shape == want.shape and np.abs(got - want).max() < 1e-5
