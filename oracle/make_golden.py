#!/usr/bin/env python3
"""Generate tests/golden/*.npz by running the UNMODIFIED reference encoder.

TEST INFRASTRUCTURE ONLY.  Run in the build container (the GPU box has no
/root/reference):

    PYTHONPATH=/root/reference python oracle/make_golden.py

For every case it builds the reference ``ConformerEncoder``
(espnet/nets/pytorch_backend/encoder/conformer_encoder.py:186), loads the
deterministic synthetic weights of ``auto_avsr_b200.synthetic`` with
``load_state_dict(strict=True)``, runs ``forward(xs, masks)`` in eval mode in
float64 and float32 with ``masks = make_non_pad_mask(lengths).unsqueeze(-2)``
(e2e_asr_conformer.py:67) -- or ``None`` -- and stores the outputs.  Weights and
inputs are NOT stored (they regenerate from the seeds in the file); layer-0 stage
intermediates come from forward hooks on the reference sub-modules.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from espnet.nets.pytorch_backend.encoder.conformer_encoder import ConformerEncoder  # noqa: E402
from espnet.nets.pytorch_backend.nets_utils import make_non_pad_mask, make_pad_mask  # noqa: E402

from auto_avsr_b200.synthetic import encoder_input, encoder_state_dict  # noqa: E402

CASES = [
    # name, config, lengths (None => mask None, B=1), wseed, xseed
    dict(name="tiny_ragged", d_model=128, n_heads=2, linear_units=256, num_blocks=2, cnn_kernel=31,
         lengths=[37, 29, 18], masked=True, wseed=11, xseed=21),
    dict(name="tiny_k7_nomask", d_model=128, n_heads=2, linear_units=256, num_blocks=1, cnn_kernel=7,
         lengths=[23], masked=False, wseed=12, xseed=22),
    dict(name="full2_ragged", d_model=768, n_heads=12, linear_units=3072, num_blocks=2, cnn_kernel=31,
         lengths=[50, 33], masked=True, wseed=13, xseed=23),
    dict(name="full12_s1", d_model=768, n_heads=12, linear_units=3072, num_blocks=12, cnn_kernel=31,
         lengths=[100], masked=False, wseed=0, xseed=1234),
    dict(name="full12_ragged", d_model=768, n_heads=12, linear_units=3072, num_blocks=12, cnn_kernel=31,
         lengths=[60, 45, 30], masked=True, wseed=0, xseed=1235),
]


def run_reference(case, dtype):
    enc = ConformerEncoder(attention_dim=case["d_model"], attention_heads=case["n_heads"],
                           linear_units=case["linear_units"], num_blocks=case["num_blocks"],
                           cnn_module_kernel=case["cnn_kernel"])
    sd = encoder_state_dict(case["wseed"], case["d_model"], case["n_heads"], case["linear_units"],
                            case["num_blocks"], case["cnn_kernel"])
    enc.load_state_dict(sd, strict=True)
    enc = enc.to(dtype).eval()
    xs = encoder_input(case["lengths"], case["d_model"], case["xseed"]).to(dtype)
    masks = make_non_pad_mask(case["lengths"]).unsqueeze(-2) if case["masked"] else None

    stages = []
    l0 = enc.encoders[0]
    # stage outputs of layer 0 = inputs of the next LayerNorm (pre-hooks), then norm_final's output
    hooks = [
        l0.norm_mha.register_forward_pre_hook(lambda m, a: stages.append(a[0].detach().clone())),
        l0.norm_conv.register_forward_pre_hook(lambda m, a: stages.append(a[0].detach().clone())),
        l0.norm_ff.register_forward_pre_hook(lambda m, a: stages.append(a[0].detach().clone())),
        l0.norm_final.register_forward_pre_hook(lambda m, a: stages.append(a[0].detach().clone())),
        l0.norm_final.register_forward_hook(lambda m, a, o: stages.append(o.detach().clone())),
    ]
    with torch.no_grad():
        out, out_masks = enc(xs, masks)
    for h in hooks:
        h.remove()
    assert out_masks is masks
    attn0 = l0.self_attn.attn.detach().clone()      # (B,H,T,T) side effect, attention.py:75
    return out, stages, attn0


def main():
    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
    torch.manual_seed(0)
    for case in CASES:
        out64, st64, attn64 = run_reference(case, torch.float64)
        out32, _, _ = run_reference(case, torch.float32)
        d = dict(
            config=np.array([case["d_model"], case["n_heads"], case["linear_units"],
                             case["num_blocks"], case["cnn_kernel"]], dtype=np.int64),
            lengths=np.array(case["lengths"], dtype=np.int64),
            masked=np.array(int(case["masked"])), wseed=np.array(case["wseed"]), xseed=np.array(case["xseed"]),
            out_f64=out64.numpy(), out_f32=out32.numpy(),
        )
        small = case["d_model"] <= 128
        if small or case["name"] == "full2_ragged":
            for i, s in enumerate(st64):
                # the full-size case keeps fp32 copies of the fp64 stages to bound fixture size
                d[f"stage{i}"] = s.numpy() if small else s.float().numpy()
        if small:
            d["attn0"] = attn64.numpy()
        if not small:
            d["out_f64"] = out64.numpy().astype(np.float32)   # fp64 reference rounded to fp32 (|err| <= 6e-8 rel)
            d["out_f64_checksum"] = np.array([out64.double().sum().item(), out64.double().abs().sum().item(),
                                              (out64.double() ** 2).sum().item()])
        path = os.path.join(ROOT, "tests", "golden", case["name"] + ".npz")
        np.savez_compressed(path, **d)
        print(case["name"], "out", tuple(out64.shape), "f32-vs-f64 max-abs",
              (out32.double() - out64).abs().max().item(), os.path.getsize(path) // 1024, "KiB")

    # the state-dict contract: key -> shape of the reference's default-config encoder
    enc = ConformerEncoder()
    with open(os.path.join(ROOT, "tests", "golden", "encoder_state_keys.txt"), "w") as f:
        for k, v in enc.state_dict().items():
            f.write(f"{k} {' '.join(str(d) for d in v.shape)}\n")

    # known-answer table from the reference docstring (nets_utils.py:82-90): lengths [5,3,2]
    m = make_pad_mask([5, 3, 2])
    np.savez(os.path.join(ROOT, "tests", "golden", "pad_mask_532.npz"),
             pad_mask=m.numpy(), non_pad_mask=make_non_pad_mask([5, 3, 2]).numpy())
    print("pad_mask", m.int().tolist())


if __name__ == "__main__":
    main()
