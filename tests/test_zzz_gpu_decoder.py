"""GPU parity of row 8f #3 (attention-decoder scoring, CTC prefix scorer, beam loop) -- through the C ABI, against the
fixtures generated from the reference's own TransformerDecoder / CTCPrefixScoreTH / BatchBeamSearch
(tests/golden/decoder_*.npz) and the fp64 oracle.

STATUS: written in the last session of round 2 with no GPU minutes left: the per-element functors, the launch
schedule, the slot / ancestor addressing and all host logic are checked on the CPU against the same fixtures through
tests/emu (test_decoder_emu_cpu.py, test_decoder_dropin_cpu.py); the GEMM / LayerNorm / log-softmax launchers the
schedule calls are the encoder path's GPU-verified kernels.  What has NOT run on a B200 is the composition on the
device, hence ``xfail(strict=False)`` with that reason: a pass shows up as XPASS, a failure does not mask the other 92
GPU tests.  The bounds are provisional (to be pinned at <= 4x observed like the rest)."""
import os
import subprocess
import sys

import pytest
import torch

from helpers import err_stats, load_decoder_case, record
from oracle import decoder_oracle as DO
from oracle import head_oracle as HO

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HERE = os.path.dirname(os.path.abspath(__file__))
UNVERIFIED = os.environ.get("AVSR_B200_DECODER_VERIFIED", "0") != "1"

pytestmark = [pytest.mark.gpu,
              pytest.mark.xfail(UNVERIFIED, reason="row 8f#3 device composition has not run on a B200 yet (no GPU minutes left "
                                                   "in round 2); CPU-verified through tests/emu", strict=False)]

# provisional (max-abs on log-probabilities of O(5) magnitude)
TOL_LOGP = {"fp32": 2e-4, "tf32": 5e-2, "f16": 5e-2}


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "run with gpurun"
    import __graft_entry__ as g
    g.build()
    return torch.device("cuda:0")


def _decoder(c, dev, prec):
    from auto_avsr_b200 import TransformerDecoder
    cfg = c["cfg"]
    dec = TransformerDecoder(cfg["odim"], cfg["d_model"], cfg["n_heads"], cfg["linear_units"], cfg["num_blocks"])
    dec.load_state_dict(c["dec_sd"], strict=True)
    dec = dec.to(dev).eval()
    dec.precision = prec
    return dec


def _ctc(c, dev, prec):
    from auto_avsr_b200 import CTC
    cfg = c["cfg"]
    ctc = CTC(cfg["odim"], cfg["d_model"], 0.1)
    ctc.load_state_dict({"ctc_lo.weight": c["head_sd"]["ctc.ctc_lo.weight"], "ctc_lo.bias": c["head_sd"]["ctc.ctc_lo.bias"]})
    ctc = ctc.to(dev).eval()
    ctc.precision = prec
    return ctc


@pytest.mark.parametrize("prec", ["fp32", "tf32", "f16"])
@pytest.mark.parametrize("name", ["decoder_tiny", "decoder_full"])
def test_decoder_batch_score_matches_reference_fixture(dev, name, prec):
    """TransformerDecoder.batch_score driven the way the beam search drives it (states handed back), vs the reference's
    batch_score outputs."""
    c = load_decoder_case(name)
    z, cfg = c["z"], c["cfg"]
    dec = _decoder(c, dev, prec)
    cols = torch.from_numpy(z["cols"])
    mem = c["memory"].to(dev)
    states = [None] * cfg["n_hyp"]
    worst = 0.0
    for step in range(cfg["steps"]):
        ys = c["prefixes"][step].to(dev)
        logp, states = dec.batch_score(ys, states, mem.unsqueeze(0).expand(cfg["n_hyp"], -1, -1))
        torch.cuda.synchronize()
        mx, _ = err_stats(logp.cpu()[:, cols], torch.from_numpy(z[f"dec_logp_f64_{step}"]))
        worst = max(worst, mx)
        assert torch.allclose(logp.exp().sum(-1).cpu(), torch.ones(cfg["n_hyp"]), atol=1e-4)
        if prec == "fp32":
            assert torch.equal(torch.topk(logp.cpu(), 3, dim=-1)[1], torch.from_numpy(z[f"dec_top_f64_{step}"])[:, :3])
    record("decoder_batch_score", (name, prec), worst, TOL_LOGP[prec])
    assert worst < TOL_LOGP[prec], (name, prec, worst)


@pytest.mark.parametrize("prec", ["fp32", "f16"])
def test_decoder_follows_reordered_and_forked_beams(dev, prec):
    c = load_decoder_case("decoder_tiny")
    cfg = c["cfg"]
    dec = _decoder(c, dev, prec)
    eng = dec.engine()
    odim, sos = cfg["odim"], cfg["odim"] - 1
    mem = c["memory"]
    eng.begin(dec, mem.to(dev), max_hyps=5, max_steps=6, precision=prec)
    g = torch.Generator().manual_seed(5)
    prefixes, chains = [[sos]], [[]]
    eng.step(torch.tensor([sos], dtype=torch.int32, device=dev), None, 0)
    worst = 0.0
    for step in range(1, 5):
        n_prev = len(prefixes)
        n = min(5, n_prev + 2)
        parents = torch.randint(0, n_prev, (n,), generator=g).tolist()
        toks = torch.randint(1, odim - 1, (n,), generator=g).tolist()
        prefixes = [prefixes[p] + [t] for p, t in zip(parents, toks)]
        chains = [chains[p] + [p] for p in parents]
        anc = torch.tensor(chains, dtype=torch.int32).T.contiguous().to(dev)
        logp = eng.step(torch.tensor(toks, dtype=torch.int32, device=dev), anc, step)
        ref = DO.decoder_logp(c["dec_sd"], torch.tensor(prefixes), mem.double(), cfg["n_heads"])
        worst = max(worst, err_stats(logp.cpu(), ref)[0])
    record("decoder_forked_beam", prec, worst, TOL_LOGP[prec])
    assert worst < TOL_LOGP[prec], worst


@pytest.mark.parametrize("name", ["decoder_tiny", "decoder_full"])
def test_ctc_prefix_scorer_matches_reference_fixture(dev, name):
    from auto_avsr_b200.decoder import CtcPrefixEngine
    c = load_decoder_case(name)
    z, cfg = c["z"], c["cfg"]
    eos, n = cfg["odim"] - 1, cfg["n_hyp"]
    logp = HO.ctc_log_softmax(c["memory"].float(), c["head_sd"]).to(dev)
    eng = CtcPrefixEngine(logp, 0, eos)
    r_prev, s_prev = eng.initial(n)
    worst = 0.0
    for step in range(cfg["steps"]):
        ys = c["prefixes"][step]
        cand = torch.from_numpy(z[f"ctc_cand_{step}"]).to(torch.int32).to(dev)
        local, r, log_psi = eng.score(step, ys[:, -1].to(torch.int32).to(dev), r_prev, s_prev, cand)
        got = torch.gather(local, 1, cand.long()).cpu()
        want = torch.from_numpy(z[f"ctc_local_f64_{step}"])
        live = (want > DO.LOGZERO / 2) & (want < -DO.LOGZERO / 2)
        assert (((got <= DO.LOGZERO / 2) | (got >= -DO.LOGZERO / 2)) == ~live).all()
        worst = max(worst, err_stats(got[live], want[live])[0] / max(1.0, float(want[live].abs().max()) / 50))
        if f"ctc_keep_{step}" in z.files:
            keep = torch.from_numpy(z[f"ctc_keep_{step}"]).to(torch.int32).to(dev)
            r_prev, s_prev = eng.select(r, log_psi, cand, torch.arange(n, dtype=torch.int32, device=dev), keep)
            pos = (cand == keep[:, None]).int().argmax(1)
            assert torch.equal(r_prev, torch.stack([r[:, :, i, int(pos[i])] for i in range(n)], dim=2))
    record("ctc_prefix_scorer", name, worst, 2e-3)
    assert worst < 2e-3, worst


def _compare_nbest(nbest, z, tag, tol):
    """best hypothesis identical to the reference's; the listed n-best scores within tol (order may swap inside tol)"""
    assert nbest, "no hypothesis ended"
    L = int(z[f"nbest_len_{tag}"][0])
    assert nbest[0]["yseq"] == z[f"nbest_yseq_{tag}"][0, :L].tolist()
    k = min(len(nbest), len(z[f"nbest_score_{tag}"]))
    got = torch.tensor([h["score"] for h in nbest[:k]])
    want = torch.from_numpy(z[f"nbest_score_{tag}"][:k]).float()
    return float((got - want).abs().max())


@pytest.mark.parametrize("name", ["decoder_tiny", "decoder_full"])
def test_device_beam_search_matches_reference_nbest(dev, name):
    from auto_avsr_b200.beam_search import DeviceBeamSearch
    c = load_decoder_case(name)
    cfg = c["cfg"]
    dec, ctc = _decoder(c, dev, "fp32"), _ctc(c, dev, "fp32")
    bs = DeviceBeamSearch(dec, ctc, beam_size=cfg["beam"], vocab_size=cfg["odim"])
    nbest = [h.asdict() for h in bs(c["memory"].to(dev))]
    err = _compare_nbest(nbest, c["z"], "f32", 5e-3)
    record("device_beam_search_fp32", name, err, 5e-3)
    assert err < 5e-3 and len(nbest) == int(c["z"]["nbest_count_f32"])
    # the product precision: same best hypothesis on these fixtures, scores within the f16 operand noise
    dec16, ctc16 = _decoder(c, dev, "f16"), _ctc(c, dev, "f16")
    nb16 = [h.asdict() for h in DeviceBeamSearch(dec16, ctc16, beam_size=cfg["beam"], vocab_size=cfg["odim"])(c["memory"].to(dev))]
    gap = abs(nb16[0]["score"] - float(c["z"]["nbest_score_f64"][0]))
    record("device_beam_search_f16_best_score", name, gap, 0.5)
    assert gap < 0.5 and int(nb16[0]["yseq"][-1]) == cfg["odim"] - 1


@pytest.mark.skipif(not os.path.isdir(os.path.join(ROOT, "oracle", "_ref", "espnet")), reason="oracle/_ref did not travel")
def test_reference_batch_beam_search_drives_the_dropins_on_the_gpu():
    code = f"""
import sys
sys.path.insert(0, {ROOT!r}); sys.path.insert(0, {HERE!r})
from oracle import build_ref
ref = build_ref.import_reference_search()
import torch
from helpers import load_decoder_case
from test_zzz_gpu_decoder import _decoder, _ctc, _compare_nbest
from auto_avsr_b200 import CTCPrefixScorer
from auto_avsr_b200.espnet_dropin import scorer_interface
assert scorer_interface.rebind()
dev = torch.device("cuda:0")
c = load_decoder_case("decoder_tiny")
cfg = c["cfg"]
dec = _decoder(c, dev, "fp32")
ctc = CTCPrefixScorer(_ctc(c, dev, "fp32"), cfg["odim"] - 1)
token_list = [str(i) for i in range(cfg["odim"])]
scorers = dict(decoder=dec, ctc=ctc, lm=None, length_bonus=ref["LengthBonus"](len(token_list)))
weights = dict(decoder=0.9, ctc=0.1, lm=0.0, length_bonus=0)
bs = ref["BatchBeamSearch"](beam_size=cfg["beam"], vocab_size=len(token_list), weights=weights, scorers=scorers,
                            sos=cfg["odim"] - 1, eos=cfg["odim"] - 1, token_list=token_list, pre_beam_score_key="decoder")
with torch.no_grad():
    nbest = [h.asdict() for h in bs(c["memory"].to(dev))]
err = _compare_nbest(nbest, c["z"], "f32", 5e-3)
assert err < 5e-3 and len(nbest) == int(c["z"]["nbest_count_f32"]), err
print("CHILD-OK", err)
"""
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0 and "CHILD-OK" in r.stdout, r.stdout[-1500:] + r.stderr[-2500:]


def test_encoder_to_nbest_chain_on_the_dropins(dev):
    """lightning.py:69-75 after the front-end, on the drop-ins end to end: proj_encoder -> encoder(x, None) -> beam search
    (decoder + CTC prefix scorer): every hypothesis is <eos>-closed, scores finite, n-best sorted."""
    from auto_avsr_b200 import CTC, ConformerEncoder, TransformerDecoder
    from auto_avsr_b200.beam_search import DeviceBeamSearch
    from auto_avsr_b200.synthetic import decoder_state_dict, encoder_input, encoder_state_dict, head_state_dict
    enc = ConformerEncoder(num_blocks=2)
    enc.load_state_dict(encoder_state_dict(3, num_blocks=2), strict=True)
    dec = TransformerDecoder(odim=5049, attention_dim=768, attention_heads=12, linear_units=3072, num_blocks=6)
    dec.load_state_dict(decoder_state_dict(4), strict=True)
    hsd = head_state_dict(4)
    ctc = CTC(5049, 768, 0.1)
    ctc.load_state_dict({"ctc_lo.weight": hsd["ctc.ctc_lo.weight"], "ctc_lo.bias": hsd["ctc.ctc_lo.bias"]})
    enc, dec, ctc = enc.to(dev).eval(), dec.to(dev).eval(), ctc.to(dev).eval()
    x = encoder_input([40], 768, 9).to(dev)
    with torch.no_grad():
        feat, _ = enc(x, None)
    nbest = DeviceBeamSearch(dec, ctc, beam_size=10)(feat.squeeze(0))
    assert nbest and all(int(h.yseq[0]) == 5048 and int(h.yseq[-1]) == 5048 for h in nbest)
    scores = [h.score for h in nbest]
    assert all(s == s and abs(s) < 1e6 for s in scores) and scores == sorted(scores, reverse=True)
