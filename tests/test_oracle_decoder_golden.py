"""Pin the CPU oracle of the attention-decoder scoring path, the CTC prefix scorer and the beam loop (SURVEY.md section
8f #3) against outputs of the reference's own TransformerDecoder / CTCPrefixScoreTH / BatchBeamSearch
(tests/golden/decoder_*.npz, made by oracle/make_golden_decoder.py in the build container).  CPU only."""
import pytest
import torch

from oracle import decoder_oracle as DO
from oracle import head_oracle as HO
from helpers import err_stats, load_decoder_case

CASES = ["decoder_tiny", "decoder_full"]


def _ctc_logp(c, dtype):
    return HO.ctc_log_softmax(c["memory"].to(dtype), c["head_sd"])


@pytest.mark.parametrize("name", CASES)
def test_decoder_oracle_matches_reference_batch_score(name):
    c = load_decoder_case(name)
    z, cfg = c["z"], c["cfg"]
    cols = torch.from_numpy(z["cols"])
    for dtype, tag, tol in ((torch.float64, "f64", 1e-10), (torch.float32, "f32", 2e-4)):
        for step in range(cfg["steps"]):
            logp = DO.decoder_logp(c["dec_sd"], c["prefixes"][step], c["memory"].to(dtype), cfg["n_heads"])
            mx, _ = err_stats(logp[:, cols], torch.from_numpy(z[f"dec_logp_{tag}_{step}"]))
            assert mx < tol, (name, tag, step, mx)
            if tag == "f64":
                assert torch.equal(torch.topk(logp, 8, dim=-1)[1], torch.from_numpy(z[f"dec_top_f64_{step}"]))
                assert err_stats((logp ** 2).sum(-1), torch.from_numpy(z[f"dec_sumsq_f64_{step}"]))[0] < 1e-6
                assert torch.allclose(logp.exp().sum(-1), torch.ones(cfg["n_hyp"], dtype=dtype), atol=1e-12)


@pytest.mark.parametrize("name", CASES)
def test_ctc_prefix_oracle_matches_reference(name):
    c = load_decoder_case(name)
    z, cfg = c["z"], c["cfg"]
    eos = cfg["odim"] - 1
    logp = _ctc_logp(c, torch.float64)
    if "ctc_logp_f64" in z.files:
        assert err_stats(logp, torch.from_numpy(z["ctc_logp_f64"]))[0] < 1e-10
    r0, s0 = DO.ctc_initial_state(logp)
    r_prev, s_prev = r0.expand(-1, -1, cfg["n_hyp"]).clone(), s0.expand(cfg["n_hyp"]).clone()
    for step in range(cfg["steps"]):
        ys = c["prefixes"][step]
        cand = torch.from_numpy(z[f"ctc_cand_{step}"])
        local, r, log_psi = DO.ctc_prefix_scores(logp, step, ys[:, -1].tolist(), r_prev, s_prev, cand, 0, eos)
        got = torch.gather(local, 1, cand)
        assert err_stats(got, torch.from_numpy(z[f"ctc_local_f64_{step}"]))[0] < 1e-9, (name, step)
        assert err_stats(local[:, eos], torch.from_numpy(z[f"ctc_eos_f64_{step}"]))[0] < 1e-9
        off = torch.ones_like(local, dtype=torch.bool).scatter_(1, cand, False)
        off[:, eos] = False
        live = s_prev > DO.LOGZERO / 2        # (a prefix extended by blank / eos carries s_prev = logzero itself)
        assert (local[live][off[live]] <= DO.LOGZERO / 2).all()      # off the candidate list: logzero
        if f"ctc_keep_{step}" in z.files:
            keep = torch.from_numpy(z[f"ctc_keep_{step}"])
            pos = (cand == keep[:, None]).int().argmax(1)
            r_prev = torch.stack([r[:, :, i, int(pos[i])] for i in range(cfg["n_hyp"])], dim=2)
            s_prev = log_psi[torch.arange(cfg["n_hyp"]), keep]


@pytest.mark.parametrize("name", CASES)
def test_beam_search_oracle_matches_reference_nbest(name):
    c = load_decoder_case(name)
    z, cfg = c["z"], c["cfg"]
    mem = c["memory"].double()
    nbest = DO.beam_search(lambda ys: DO.decoder_logp(c["dec_sd"], ys, mem, cfg["n_heads"]), _ctc_logp(c, torch.float64),
                           cfg["odim"], cfg["beam"], maxlen=cfg["T"])
    assert len(nbest) == int(z["nbest_count_f64"])
    for i in range(len(z["nbest_len_f64"])):
        L = int(z["nbest_len_f64"][i])
        assert nbest[i]["yseq"] == z["nbest_yseq_f64"][i, :L].tolist(), (name, i)
        assert abs(nbest[i]["score"] - float(z["nbest_score_f64"][i])) < 1e-8
        assert abs(nbest[i]["scores"]["decoder"] - float(z["nbest_dec_f64"][i])) < 1e-8
        assert abs(nbest[i]["scores"]["ctc"] - float(z["nbest_ctc_f64"][i])) < 1e-8


def test_end_detect_restatement():
    ended = [dict(yseq=[0] * 5, score=-1.0), dict(yseq=[0] * 9, score=-30.0), dict(yseq=[0] * 8, score=-31.0),
             dict(yseq=[0] * 7, score=-40.0)]
    assert DO.end_detect(ended, 9)                 # lengths 9, 8, 7 all more than 10 below the best
    assert not DO.end_detect(ended, 8)             # no ended hypothesis of length 6
    assert not DO.end_detect([], 3)
