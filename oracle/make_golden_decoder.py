#!/usr/bin/env python3
"""Generate tests/golden/decoder_*.npz by running the UNMODIFIED reference decoder, CTC prefix scorer and batch beam
search.  TEST INFRASTRUCTURE ONLY; build container only:

    PYTHONPATH=/root/reference python oracle/make_golden_decoder.py

Per case it stores what the reference computed, in float64 and float32:
  * ``batch_score`` of the reference ``TransformerDecoder`` (transformer_decoder.py:302-334) driven for a few steps with
    its own layer-output cache, on fixed prefixes;
  * ``CTCPrefixScoreTH.__call__`` (ctc_prefix_score.py:72-200) for the same steps, with the state selection of
    ``CTCPrefixScorer.select_state`` (scorers/ctc.py:37-60);
  * the n-best list of ``BatchBeamSearch`` built exactly like ``get_beam_search_decoder`` (lightning.py:126-157).
Weights and inputs regenerate from seeds (``auto_avsr_b200.synthetic``); only outputs are stored.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from espnet.nets.batch_beam_search import BatchBeamSearch  # noqa: E402
from espnet.nets.ctc_prefix_score import CTCPrefixScoreTH  # noqa: E402
from espnet.nets.pytorch_backend.ctc import CTC  # noqa: E402
from espnet.nets.pytorch_backend.decoder.transformer_decoder import TransformerDecoder  # noqa: E402
from espnet.nets.scorers.ctc import CTCPrefixScorer  # noqa: E402
from espnet.nets.scorers.length_bonus import LengthBonus  # noqa: E402

from auto_avsr_b200.synthetic import decoder_state_dict, encoder_input, head_state_dict  # noqa: E402

CASES = [
    dict(name="decoder_tiny", odim=37, d_model=128, n_heads=2, linear_units=256, num_blocks=2, T=23, beam=5,
         wseed=51, xseed=61, steps=4, n_hyp=3),
    dict(name="decoder_full", odim=5049, d_model=768, n_heads=12, linear_units=3072, num_blocks=6, T=30, beam=40,
         wseed=52, xseed=62, steps=3, n_hyp=4),
]
ROWS_KEPT = 64        # of the (n, 5049) score rows of the full case only a column subset + row statistics are stored


def memory_of(case, dtype):
    """(T, d) stand-in for the encoder output: LayerNorm-like O(1) rows."""
    return encoder_input([case["T"]], case["d_model"], case["xseed"]).to(dtype)[0]


def prefixes_of(case, step):
    """fixed token prefixes (n_hyp, step + 1): <sos> then seeded tokens (never blank / eos)"""
    g = torch.Generator().manual_seed(case["xseed"] * 7 + 1)
    body = torch.randint(1, case["odim"] - 1, (case["n_hyp"], case["steps"]), generator=g)
    sos = torch.full((case["n_hyp"], 1), case["odim"] - 1, dtype=torch.long)
    return torch.cat([sos, body[:, :step]], dim=1)


def build(case, dtype):
    dec = TransformerDecoder(odim=case["odim"], attention_dim=case["d_model"], attention_heads=case["n_heads"],
                             linear_units=case["linear_units"], num_blocks=case["num_blocks"])
    dec.load_state_dict(decoder_state_dict(case["wseed"], case["odim"], case["d_model"], case["n_heads"],
                                           case["linear_units"], case["num_blocks"]), strict=True)
    hsd = head_state_dict(case["wseed"], 64, case["d_model"], case["odim"])
    ctc = CTC(case["odim"], case["d_model"], 0.1, reduce=True)
    ctc.load_state_dict({"ctc_lo.weight": hsd["ctc.ctc_lo.weight"], "ctc_lo.bias": hsd["ctc.ctc_lo.bias"]}, strict=True)
    return dec.to(dtype).eval(), ctc.to(dtype).eval()


def run_reference(case, dtype):
    dec, ctc = build(case, dtype)
    mem = memory_of(case, dtype)
    eos = case["odim"] - 1
    out = {}
    with torch.no_grad():
        # --- decoder: batch_score step by step with the reference's own cache --------------------------------------
        states = [None] * case["n_hyp"]
        for step in range(case["steps"]):
            ys = prefixes_of(case, step)
            logp, states = dec.batch_score(ys, states, mem.unsqueeze(0).expand(case["n_hyp"], -1, -1))
            out[f"dec_logp_{step}"] = logp
        # --- CTC prefix scorer: candidates = top pre-beam tokens of the decoder scores of that step ----------------
        ctc_logp = ctc.log_softmax(mem.unsqueeze(0))
        out["ctc_logp"] = ctc_logp[0]
        impl = CTCPrefixScoreTH(ctc_logp.clone(), torch.tensor([case["T"]]), 0, eos)
        scorer = CTCPrefixScorer(ctc, eos)
        pre = int(1.5 * case["beam"])
        state = None
        for step in range(case["steps"]):
            ys = prefixes_of(case, step)
            cand = torch.topk(out[f"dec_logp_{step}"], pre, dim=-1)[1]
            if state is not None:                     # what batch_score_partial stacks (scorers/ctc.py:118-128)
                state = (torch.stack([s[0] for s in state], dim=2), torch.stack([s[1] for s in state]), state[0][2], state[0][3])
            local, new_state = impl(ys, state, cand)
            out[f"ctc_local_{step}"] = local
            out[f"ctc_cand_{step}"] = cand
            nxt = prefixes_of(case, step + 1)[:, -1] if step + 1 < case["steps"] else None
            if nxt is not None:
                # keep hypothesis i extended by its NEXT fixed token when that token is a candidate, else by its best candidate
                keep = []
                for i in range(case["n_hyp"]):
                    tok = int(nxt[i]) if int(nxt[i]) in cand[i].tolist() else int(cand[i, 0])
                    keep.append(tok)
                out[f"ctc_keep_{step}"] = torch.tensor(keep)
                state = [scorer.select_state(new_state, i, keep[i]) for i in range(case["n_hyp"])]
        # --- the whole search ------------------------------------------------------------------------------------
        token_list = [str(i) for i in range(case["odim"])]
        scorers = dict(decoder=dec, ctc=CTCPrefixScorer(ctc, eos), lm=None, length_bonus=LengthBonus(len(token_list)))
        weights = dict(decoder=0.9, ctc=0.1, lm=0.0, length_bonus=0)
        bs = BatchBeamSearch(beam_size=case["beam"], vocab_size=len(token_list), weights=weights, scorers=scorers,
                             sos=eos, eos=eos, token_list=token_list, pre_beam_score_key="decoder")
        nbest = [h.asdict() for h in bs(mem)]
        out["nbest"] = nbest
    return out


def main():
    out_dir = os.path.join(ROOT, "tests", "golden")
    for case in CASES:
        r64 = run_reference(case, torch.float64)
        r32 = run_reference(case, torch.float32)
        full = case["odim"] > 1000
        g = torch.Generator().manual_seed(99)
        cols = torch.sort(torch.randperm(case["odim"], generator=g)[:ROWS_KEPT])[0] if full else torch.arange(case["odim"])
        store = {"cols": cols.numpy()}
        for step in range(case["steps"]):
            lp64, lp32 = r64[f"dec_logp_{step}"], r32[f"dec_logp_{step}"]
            store[f"dec_logp_f64_{step}"] = lp64[:, cols].numpy()
            store[f"dec_logp_f32_{step}"] = lp32[:, cols].numpy()
            store[f"dec_top_f64_{step}"] = torch.topk(lp64, 8, dim=-1)[1].numpy()
            store[f"dec_topv_f64_{step}"] = torch.topk(lp64, 8, dim=-1)[0].numpy()
            store[f"dec_sumsq_f64_{step}"] = (lp64 ** 2).sum(-1).numpy()
            store[f"ctc_cand_{step}"] = r64[f"ctc_cand_{step}"].numpy()
            # local scores at the candidate positions (+ eos) -- everything else is logzero - s_prev
            cand = r64[f"ctc_cand_{step}"]
            store[f"ctc_local_f64_{step}"] = torch.gather(r64[f"ctc_local_{step}"], 1, cand).numpy()
            store[f"ctc_local_f32_{step}"] = torch.gather(r32[f"ctc_local_{step}"], 1, r32[f"ctc_cand_{step}"]).numpy()
            store[f"ctc_cand_f32_{step}"] = r32[f"ctc_cand_{step}"].numpy()
            store[f"ctc_eos_f64_{step}"] = r64[f"ctc_local_{step}"][:, case["odim"] - 1].numpy()
            if f"ctc_keep_{step}" in r64:
                store[f"ctc_keep_{step}"] = r64[f"ctc_keep_{step}"].numpy()
        if not full:
            store["ctc_logp_f64"] = r64["ctc_logp"].numpy()
        for tag, r in (("f64", r64), ("f32", r32)):
            nb = r["nbest"][:10]
            store[f"nbest_len_{tag}"] = np.asarray([len(h["yseq"]) for h in nb])
            width = max(len(h["yseq"]) for h in nb)
            store[f"nbest_yseq_{tag}"] = np.asarray([h["yseq"] + [-1] * (width - len(h["yseq"])) for h in nb])
            store[f"nbest_score_{tag}"] = np.asarray([h["score"] for h in nb])
            store[f"nbest_dec_{tag}"] = np.asarray([h["scores"]["decoder"] for h in nb])
            store[f"nbest_ctc_{tag}"] = np.asarray([h["scores"]["ctc"] for h in nb])
            store[f"nbest_count_{tag}"] = np.asarray(len(r["nbest"]))
        meta = {k: v for k, v in case.items() if k != "name"}
        np.savez_compressed(os.path.join(out_dir, case["name"] + ".npz"),
                            **store, **{"cfg_" + k: np.asarray(v) for k, v in meta.items()})
        print(case["name"], "nbest", len(r64["nbest"]), "best len", len(r64["nbest"][0]["yseq"]), "score",
              r64["nbest"][0]["score"], "f32", r32["nbest"][0]["score"],
              "same best yseq f32/f64:", r64["nbest"][0]["yseq"] == r32["nbest"][0]["yseq"])


if __name__ == "__main__":
    main()
