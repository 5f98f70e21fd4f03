#!/usr/bin/env python3
"""Where does a launch spend its cycles?  Runs the S2 forward through the phase-trace build (scripts/build_trace.py)
and prints, per kernel kind, the median cycles between consecutive phase marks over all CTAs.

    python scripts/build_trace.py && python scripts/phase_probe.py [steps] [graph|direct]

Kernel ids: 200 + epilogue mode = two-SM GEMM (aux = pair-tile width | k-blocks << 16), 300 = fp16 attention
(aux = key tiles | query tile << 8).  Mark meanings are listed next to the AVSR_TRACE_OPEN calls in the kernels."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
LIB = os.path.join(ROOT, "auto_avsr_b200", "csrc", "libavsr_b200_trace.so")
os.environ["AVSR_B200_LIB"] = LIB

import numpy as np  # noqa: E402
import torch  # noqa: E402
from auto_avsr_b200 import ConformerEncoder, _cabi  # noqa: E402
from auto_avsr_b200.synthetic import SHAPES, encoder_input, encoder_state_dict  # noqa: E402

WORDS = 16
GEMM_MARKS = ["prologue", "dep-wait", "1st TMA", "TMA issue", "1st full", "main loop", "acc visible", "epilogue", "drain"]
ATT_MARKS = ["prologue", "dep-wait", "1st S/G issue", "S/G(0) ready", "softmax(0)", "P.V(0)", "rest of tiles", "drain"]


def main() -> None:
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    graph = (sys.argv[2] if len(sys.argv) > 2 else "graph") == "graph"
    dev = torch.device("cuda:0")
    lengths = list(SHAPES["S2"])
    enc = ConformerEncoder()
    enc.load_state_dict(encoder_state_dict(0))
    enc = enc.to(dev).eval()
    enc.use_graph = graph
    enc.assume_frozen = True
    xs = encoder_input(lengths).to(dev)
    mask = (torch.arange(max(lengths))[None, :] < torch.tensor(lengths)[:, None]).unsqueeze(1).to(dev)
    with torch.no_grad():
        for _ in range(3):                      # plan creation, warm-up: untraced
            enc(xs, mask)
    torch.cuda.synchronize()
    cap = 200_000
    buf = torch.zeros(2 + cap * WORDS, dtype=torch.int64, device=dev)
    fn = _cabi.lib.avsr_trace_set
    fn.restype, fn.argtypes = C.c_int, [C.c_void_p, C.c_size_t, C.c_void_p]
    _cabi.check(fn(buf.data_ptr(), buf.numel(), None))
    with torch.no_grad():
        for _ in range(steps):
            enc(xs, mask)
    torch.cuda.synchronize()
    _cabi.check(fn(None, 0, None))
    host = buf.cpu().numpy().astype(np.uint64)
    n = int(min(host[0], host[1]))
    rec = host[2:2 + n * WORDS].reshape(n, WORDS)
    print(f"{n} CTA records over {steps} forward(s), graph={graph}")
    keys = sorted({(int(r[0]), int(r[1] >> np.uint64(32))) for r in rec})
    for kid, aux in keys:
        sel = rec[(rec[:, 0] == kid) & ((rec[:, 1] >> np.uint64(32)) == aux)]
        marks = sel[:, 4:].astype(np.int64)
        names = ATT_MARKS if kid == 300 else GEMM_MARKS
        if kid == 300:
            label = f"attention_f16 key_tiles={aux & 0xff} qtile={aux >> 8}"
        else:
            label = f"gemm_tc2 mode={kid - 200} pair_tile=256x{aux & 0xffff} k_blocks={aux >> 16}"
        print(f"\n{label}: {len(sel)} CTAs")
        prev = None
        for s, name in enumerate(names):
            col = marks[:, s]
            ok = col > 0
            if prev is not None:
                both = ok & (marks[:, prev] > 0)
                if both.any():
                    d = (col[both] - marks[both, prev])
                    print(f"  {names[prev]:>14s} -> {name:<14s} median {int(np.median(d)):>7d}  p90 {int(np.percentile(d, 90)):>7d} cycles"
                          f"  ({both.sum()} CTAs)")
            if ok.any():
                prev = s
        first = np.where(marks > 0, marks, np.iinfo(np.int64).max).min(axis=1)
        last = marks.max(axis=1)
        print(f"  {'first -> last mark':>32s} median {int(np.median(last - first)):>7d} cycles")
        # launch-level: spread of the CTA open times (globaltimer, ns)
        gt = sel[:, 3].astype(np.int64)
        print(f"  {'CTA start spread (all launches)':>32s} {int(gt.max() - gt.min())} ns over {len(sel)} CTAs")


if __name__ == "__main__":
    main()
