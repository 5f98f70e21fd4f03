"""Pin the CPU oracle of the steps either side of the encoder (proj_encoder, CTC head: SURVEY.md section 8f #1)
against outputs of the reference's own modules (tests/golden/head_*.npz, made by oracle/make_golden_head.py in the
build container).  CPU only."""
import pytest
import torch

from oracle import conformer_oracle as O
from oracle import head_oracle as HO
from helpers import err_stats, load_head_case

CASES = ["head_tiny", "head_full"]


@pytest.mark.parametrize("name", CASES)
def test_head_oracle_fp64_matches_reference_fp64(name):
    c = load_head_case(name)
    z, lengths = c["z"], c["lengths"]
    feats = c["feats"].double()
    x = HO.proj_encoder(feats, c["head_sd"])
    assert err_stats(x, torch.from_numpy(z["proj_f64"]))[0] < 1e-12
    hs = O.encoder_forward(c["enc_sd"], x, lengths if c["masked"] else None, c["cfg"]["n_heads"])
    assert err_stats(hs, torch.from_numpy(z["enc_f64"]))[0] < 1e-9
    logp = HO.ctc_log_softmax(hs, c["head_sd"])
    assert err_stats(logp, torch.from_numpy(z["logp_f64"]))[0] < 1e-9
    # the whole chain in one call, and the other two public methods of the reference's CTC module
    chain = HO.features_to_log_probs(c["head_sd"], c["enc_sd"], feats, lengths if c["masked"] else None,
                                     c["cfg"]["n_heads"])
    assert torch.equal(chain, logp)
    assert err_stats(HO.ctc_softmax(hs, c["head_sd"]).sum(-1), torch.from_numpy(z["prob_rowsum_f64"]))[0] < 1e-12
    assert torch.equal(HO.ctc_argmax(hs, c["head_sd"]), torch.from_numpy(z["argmax_f64"]))


@pytest.mark.parametrize("name", CASES)
def test_head_oracle_fp32_matches_reference_fp32(name):
    c = load_head_case(name)
    logp = HO.features_to_log_probs(c["head_sd"], c["enc_sd"], c["feats"].float(),
                                    c["lengths"] if c["masked"] else None, c["cfg"]["n_heads"])
    mx, rms = err_stats(logp, torch.from_numpy(c["z"]["logp_f32"]))
    assert mx < 5e-5 and rms < 5e-6, (name, mx, rms)


def test_log_probs_are_normalised_and_padding_rows_are_data():
    c = load_head_case("head_tiny")
    logp = HO.features_to_log_probs(c["head_sd"], c["enc_sd"], c["feats"].double(), c["lengths"], c["cfg"]["n_heads"])
    assert torch.allclose(logp.exp().sum(-1), torch.ones(logp.shape[:2], dtype=torch.float64), atol=1e-12)
    # frames beyond an utterance's length still get log-probs (the reference applies no length mask here either)
    assert torch.isfinite(logp).all()
