"""Pin the CPU oracle against outputs of the reference itself (tests/golden, made by
oracle/make_golden.py in the build container).  CPU only."""
import numpy as np
import pytest
import torch

from oracle import conformer_oracle as O
from helpers import err_stats, load_case

CASES = ["tiny_ragged", "tiny_k7_nomask", "full2_ragged", "full12_s1", "full12_ragged"]


@pytest.mark.parametrize("name", CASES)
def test_oracle_fp64_matches_reference_fp64(name):
    c = load_case(name)
    stages = []
    out = O.encoder_forward(c["sd"], c["xs"].double(), c["lengths"] if c["masked"] else None,
                            c["cfg"]["n_heads"], stages)
    ref = torch.from_numpy(c["z"]["out_f64"])
    small = c["cfg"]["d_model"] <= 128
    # full-size fixtures hold the fp64 reference rounded to fp32 => 1 fp32 ulp of slack
    tol = 1e-10 if small else 5e-7
    mx, rms = err_stats(out, ref)
    assert mx < tol, (name, mx, rms)
    if "stage0" in c["z"]:
        for i, s in enumerate(stages):
            g = torch.from_numpy(c["z"][f"stage{i}"])
            mx, _ = err_stats(s, g)
            scale = max(1.0, g.abs().max().item())
            assert mx < (1e-10 if small else 5e-7) * scale * (1 if small else 10), (name, i, mx)
    if "out_f64_checksum" in c["z"]:
        cs = c["z"]["out_f64_checksum"]
        got = np.array([out.sum().item(), out.abs().sum().item(), (out ** 2).sum().item()])
        assert np.allclose(got, cs, rtol=1e-9, atol=1e-7)


@pytest.mark.parametrize("name", CASES)
def test_oracle_fp32_matches_reference_fp32(name):
    c = load_case(name)
    out = O.encoder_forward(c["sd"], c["xs"].float(), c["lengths"] if c["masked"] else None,
                            c["cfg"]["n_heads"])
    ref32 = torch.from_numpy(c["z"]["out_f32"])
    mx, rms = err_stats(out, ref32)
    # two fp32 evaluations with different op order: both are within ~7e-6 of fp64
    assert mx < 3e-5 and rms < 3e-6, (name, mx, rms)


def test_attention_probabilities_tiny():
    c = load_case("tiny_ragged")
    sd = {k: v.double() for k, v in c["sd"].items() if v.is_floating_point()}
    xs = c["xs"].double()
    B, T, D = xs.shape
    x = xs * D ** 0.5
    pfx = "encoders.0."
    x = x + 0.5 * O.feed_forward(O.layer_norm(x, sd[pfx + "norm_ff_macaron.weight"], sd[pfx + "norm_ff_macaron.bias"]),
                                 sd, pfx + "feed_forward_macaron.")
    xn = O.layer_norm(x, sd[pfx + "norm_mha.weight"], sd[pfx + "norm_mha.bias"])
    _, attn = O.rel_mha(xn, O.rel_sinusoid_table(T, D, torch.float64), torch.tensor(c["lengths"]), sd,
                        pfx + "self_attn.", c["cfg"]["n_heads"], return_attn=True)
    ref = torch.from_numpy(c["z"]["attn0"])
    assert err_stats(attn, ref)[0] < 1e-12
    # masked keys get probability exactly 0, rows renormalise to 1, padded query rows are non-zero
    for b, L in enumerate(c["lengths"]):
        assert (attn[b, :, :, L:] == 0).all()
        assert torch.allclose(attn[b].sum(-1), torch.ones_like(attn[b].sum(-1)), atol=1e-12)
        if L < T:
            assert attn[b, :, L:, :L].abs().sum() > 0


def test_rel_shift_closed_form():
    """rel_shift(x)[i,j] == x[i, j-i+T-1] -- restated with the pad/view trick of attention.py:131-151."""
    T = 7
    x = torch.arange(T * (2 * T - 1), dtype=torch.float64).view(1, 1, T, 2 * T - 1)
    zero_pad = torch.zeros(1, 1, T, 1, dtype=x.dtype)
    xp = torch.cat([zero_pad, x], dim=-1).view(1, 1, 2 * T, T)[:, :, 1:].reshape(1, 1, T, 2 * T - 1)[..., :T]
    i = torch.arange(T)[:, None]
    j = torch.arange(T)[None, :]
    assert torch.equal(xp[0, 0], x[0, 0].gather(1, j - i + T - 1))


def test_sinusoid_table_orientation():
    T, D = 9, 16
    pe = O.rel_sinusoid_table(T, D, torch.float64)
    assert pe.shape == (2 * T - 1, D)
    assert torch.all(pe[T - 1, 0::2] == 0) and torch.all(pe[T - 1, 1::2] == 1)     # rel = 0
    w0 = 1.0
    assert abs(pe[0, 0].item() - np.sin(np.float32(T - 1) * w0)) < 1e-6              # rel = +(T-1) first
    assert abs(pe[-1, 0].item() + np.sin(np.float32(T - 1) * w0)) < 1e-6             # rel = -(T-1) last


def test_pad_mask_known_answer(golden_dir):
    z = np.load(golden_dir + "/pad_mask_532.npz")
    m = O.non_pad_mask([5, 3, 2])
    assert np.array_equal(m.numpy(), z["non_pad_mask"])
    # docstring table nets_utils.py:82-90
    assert (~m).int().tolist() == [[0, 0, 0, 0, 0], [0, 0, 0, 1, 1], [0, 0, 1, 1, 1]]


def test_zero_length_row_is_zero_attention():
    """All keys masked -> attention output is linear_out(0) = bias (attention.py:72-77)."""
    c = load_case("tiny_ragged")
    sd = {k: v.double() for k, v in c["sd"].items() if v.is_floating_point()}
    xs = c["xs"].double()
    B, T, D = xs.shape
    out = O.rel_mha(xs, O.rel_sinusoid_table(T, D, torch.float64), torch.tensor([T, 0, 5]), sd,
                    "encoders.0.self_attn.", 2)
    bias = sd["encoders.0.self_attn.linear_out.bias"]
    assert torch.allclose(out[1], bias.expand(T, D), atol=1e-12)


@pytest.mark.parametrize("name", ["tiny_ragged", "tiny_k7_nomask", "full2_ragged"])
def test_ref_copy_reproduces_golden(name):
    """oracle/_ref (verbatim copy of the reference encoder modules made by oracle/build_ref.py; what bench.py's
    reference arm and cpu_baseline leg time) reproduces the committed fixtures, i.e. it IS the code that generated them."""
    from oracle import build_ref
    build_ref.build()
    if not build_ref.available():
        pytest.skip("oracle/_ref not built (no /root/reference in this environment)")
    Ref, make_non_pad_mask = build_ref.import_reference_encoder()
    c = load_case(name)
    cfg = c["cfg"]
    enc = Ref(attention_dim=cfg["d_model"], attention_heads=cfg["n_heads"], linear_units=cfg["linear_units"],
              num_blocks=cfg["num_blocks"], cnn_module_kernel=cfg["cnn_kernel"])
    enc.load_state_dict(c["sd"], strict=True)
    enc = enc.double().eval()
    mask = make_non_pad_mask(c["lengths"]).unsqueeze(-2) if c["masked"] else None
    with torch.no_grad():
        out = enc(c["xs"].double(), mask)[0]
    mx, _ = err_stats(out, torch.from_numpy(c["z"]["out_f64"]))
    # the full-size fixtures keep the fp64 output rounded to fp32 (oracle/make_golden.py): <= 6e-8 relative
    assert mx < (1e-12 if cfg["d_model"] <= 128 else 5e-7), mx
