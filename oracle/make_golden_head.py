#!/usr/bin/env python3
"""Generate tests/golden/head_*.npz by running the UNMODIFIED reference modules of the steps either side of the
encoder.  TEST INFRASTRUCTURE ONLY; build container only:

    PYTHONPATH=/root/reference python oracle/make_golden_head.py

``proj_encoder`` is a plain ``torch.nn.Linear(idim, d)`` in the reference (e2e_asr_conformer.py:31), the CTC head is
``espnet.nets.pytorch_backend.ctc.CTC`` (ctc.py:9-93), the encoder is the reference ``ConformerEncoder``.  The chain is
the one ``E2E.forward`` runs at e2e_asr_conformer.py:70-71 followed by ``ctc.log_softmax`` / ``ctc.argmax``.
Weights and inputs regenerate from seeds (``auto_avsr_b200.synthetic``); only outputs are stored.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from espnet.nets.pytorch_backend.ctc import CTC  # noqa: E402
from espnet.nets.pytorch_backend.encoder.conformer_encoder import ConformerEncoder  # noqa: E402
from espnet.nets.pytorch_backend.nets_utils import make_non_pad_mask  # noqa: E402

from auto_avsr_b200.synthetic import encoder_state_dict, frontend_features, head_state_dict  # noqa: E402

CASES = [
    dict(name="head_tiny", idim=64, d_model=128, n_heads=2, linear_units=256, num_blocks=1, cnn_kernel=31, odim=37,
         lengths=[19, 12, 7], masked=True, wseed=31, xseed=41),
    dict(name="head_full", idim=512, d_model=768, n_heads=12, linear_units=3072, num_blocks=2, cnn_kernel=31,
         odim=5049, lengths=[11, 7], masked=True, wseed=32, xseed=42),
]


def run_reference(case, dtype):
    hsd = head_state_dict(case["wseed"], case["idim"], case["d_model"], case["odim"])
    proj = torch.nn.Linear(case["idim"], case["d_model"])                       # e2e_asr_conformer.py:31
    proj.load_state_dict({"weight": hsd["proj_encoder.weight"], "bias": hsd["proj_encoder.bias"]}, strict=True)
    ctc = CTC(case["odim"], case["d_model"], 0.1, reduce=True)                   # e2e_asr_conformer.py:56
    ctc.load_state_dict({"ctc_lo.weight": hsd["ctc.ctc_lo.weight"], "ctc_lo.bias": hsd["ctc.ctc_lo.bias"]}, strict=True)
    enc = ConformerEncoder(attention_dim=case["d_model"], attention_heads=case["n_heads"],
                           linear_units=case["linear_units"], num_blocks=case["num_blocks"],
                           cnn_module_kernel=case["cnn_kernel"])
    enc.load_state_dict(encoder_state_dict(case["wseed"], case["d_model"], case["n_heads"], case["linear_units"],
                                           case["num_blocks"], case["cnn_kernel"]), strict=True)
    proj, ctc, enc = proj.to(dtype).eval(), ctc.to(dtype).eval(), enc.to(dtype).eval()
    feats = frontend_features(case["lengths"], case["idim"], case["xseed"]).to(dtype)
    masks = make_non_pad_mask(case["lengths"]).unsqueeze(-2) if case["masked"] else None
    with torch.no_grad():
        x = proj(feats)
        hs, _ = enc(x, masks)
        logp = ctc.log_softmax(hs)
        prob = ctc.softmax(hs)
        best = ctc.argmax(hs)
    return x, hs, logp, prob, best


def main():
    out_dir = os.path.join(ROOT, "tests", "golden")
    for case in CASES:
        x64, hs64, lp64, pr64, am64 = run_reference(case, torch.float64)
        x32, hs32, lp32, pr32, am32 = run_reference(case, torch.float32)
        meta = {k: v for k, v in case.items() if k not in ("name", "lengths")}
        np.savez_compressed(
            os.path.join(out_dir, case["name"] + ".npz"),
            lengths=np.asarray(case["lengths"], dtype=np.int64),
            proj_f64=x64.numpy(), enc_f64=hs64.numpy(),
            logp_f64=lp64.numpy(), logp_f32=lp32.numpy(),
            prob_rowsum_f64=pr64.sum(-1).numpy(), argmax_f64=am64.numpy(), argmax_f32=am32.numpy(),
            **{"cfg_" + k: np.asarray(v) for k, v in meta.items()})
        print(case["name"], "logp", tuple(lp64.shape), "max|f32-f64|", float((lp32.double() - lp64).abs().max()),
              "argmax agree", bool((am32 == am64).all()))


if __name__ == "__main__":
    main()
