"""Opt-in kernel variants that have not run on a B200 yet (written after round 1's GPU minutes were spent).
Each test runs in a child process with the variant's environment switch and is xfail(strict=False): the default
path and the verified parity suite are unaffected whatever happens here.

AVSR_B200_ATTN=x4 -> attention_f16x.cu (16 softmax warps, four threads per query row).
AVSR_B200_PREB=1  -> two-SM GEMM fetches the weight halves of its first stages before griddepcontrol.wait."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HERE = os.path.dirname(os.path.abspath(__file__))

pytestmark = [pytest.mark.gpu,
              pytest.mark.xfail(strict=False, reason="experimental variant: first B200 run pending")]

VARIANT_CHECK = f"""
import sys, time, torch
sys.path.insert(0, {ROOT!r}); sys.path.insert(0, {HERE!r})
from helpers import err_stats, load_case
from auto_avsr_b200 import ConformerEncoder
from auto_avsr_b200.synthetic import SHAPES, encoder_input, encoder_state_dict
from oracle import conformer_oracle as O
dev = torch.device("cuda:0")
for name in ("full2_ragged", "full12_ragged"):
    c = load_case(name)
    enc = ConformerEncoder(num_blocks=c["cfg"]["num_blocks"]); enc.load_state_dict(c["sd"]); enc = enc.to(dev).eval()
    enc.precision = "f16"
    mask = O.non_pad_mask(c["lengths"]).unsqueeze(1).to(dev)
    out = enc(c["xs"].to(dev), mask)[0].cpu()
    mx, rms = err_stats(out, torch.from_numpy(c["z"]["out_f64"]))
    assert mx < 2e-2 and rms < 3e-3, (name, mx, rms)
    assert torch.equal(out, enc(c["xs"].to(dev), mask)[0].cpu())
# full size, ragged: every masking branch of the kernel
lengths = list(SHAPES["S2r"]); sd = encoder_state_dict(0); xs = encoder_input(lengths, 768, 1234)
enc = ConformerEncoder(); enc.load_state_dict(sd); enc = enc.to(dev).eval(); enc.precision = "f16"
mask = O.non_pad_mask(lengths).unsqueeze(1).to(dev)
out = enc(xs.to(dev), mask)[0]
ref = O.encoder_forward(sd, xs.float(), lengths, 12)
mx, rms = err_stats(out.cpu(), ref)
assert mx < 2e-2 and rms < 3e-3, ("S2r", mx, rms)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(20): enc(xs.to(dev), mask)
torch.cuda.synchronize(); print("variant ms/forward (S2r, incl. H2D):", (time.perf_counter() - t0) / 20 * 1e3)
print("CHILD-OK")
"""


@pytest.mark.parametrize("switch", ["AVSR_B200_ATTN=x4", "AVSR_B200_PREB=1"])
def test_variant_matches_golden_and_oracle(switch):
    name, value = switch.split("=")
    r = subprocess.run([sys.executable, "-c", VARIANT_CHECK], capture_output=True, text=True, timeout=300, cwd=ROOT,
                       env=dict(os.environ, **{name: value}))
    assert r.returncode == 0 and "CHILD-OK" in r.stdout, r.stdout[-1500:] + r.stderr[-2500:]
