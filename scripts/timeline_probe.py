#!/usr/bin/env python3
"""In-situ launch timeline of one encoder forward (CUDA-graph replay, PDL on) from the phase-trace build.

    python scripts/build_trace.py && python scripts/timeline_probe.py [workload] [steps] [out.txt]

Every CTA of every kernel records %globaltimer when it opens, when its dependency resolved (after
griddepcontrol.wait) and when it is done (common.cuh, "phase trace").  Per launch: start = first CTA open,
dep = first dependency-resolved stamp, end = last CTA done.  `exposed` = end - previous launch's end = what the
launch adds to the critical path of the forward (the kernels are serialised by data dependencies; with PDL a
launch's prologue overlaps its predecessor).  Summed over the launches it is the forward time, so the per-kind totals
are an attribution of the step that ncu's serialised, cold-cache launch list cannot give.
Also prints the intra-CTA phase medians (clock64) for the two-SM GEMM and the attention kernel."""
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
if os.environ.get("AVSR_TRACE_EPI"):      # trace library built with AVSR_TRACE_EPI=1: slots 2..5 belong to the first epilogue warp
    GEMM_MARKS = ["prologue", "dep-wait", "chunk0 in regs", "chunk0 staged", "store0 issued", "last store issued",
                  "acc visible", "epilogue", "drain"]
ATT_MARKS = ["prologue", "dep-wait", "1st S/G issue", "S/G(0) ready", "softmax(0)", "P.V(0)", "rest of tiles", "drain"]
EPI = {0: "linear", 1: "qkv", 2: "vt", 3: "glu", 4: "pos"}


def kind_name(kid, aux):
    if kid == 100:
        return f"layernorm<{aux}>"
    if kid == 110:
        return f"layernorm2<{aux}>"
    if kid == 120:
        return "dwconv_bn_silu"
    if kid == 130:
        return "embed_scale"
    if kid == 140:
        return "sinusoid"
    if 200 <= kid < 300:
        return f"gemm_tc2<{EPI.get(kid - 200, kid - 200)},{aux & 0xffff},kb={aux >> 16}>"
    if kid == 300:
        return "attention_f16"
    if 400 <= kid < 500:
        return f"gemm_tc<{EPI.get(kid - 400, kid - 400)},{aux & 0xffff}x{aux >> 16}>"
    return f"kernel{kid}"


def main() -> None:
    workload = sys.argv[1] if len(sys.argv) > 1 else "S2"
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    out_path = sys.argv[3] if len(sys.argv) > 3 and sys.argv[3] != "-" else None
    brief = len(sys.argv) > 4 and sys.argv[4] == "brief"      # per-kind summary only
    dev = torch.device("cuda:0")
    lengths = list(SHAPES[workload])
    enc = ConformerEncoder()
    enc.load_state_dict(encoder_state_dict(0))
    enc = enc.to(dev).eval()
    enc.assume_frozen = True
    xs = encoder_input(lengths).to(dev)
    mask = (torch.arange(max(lengths))[None, :] < torch.tensor(lengths)[:, None]).unsqueeze(1).to(dev)
    cap = 400_000
    buf = torch.zeros(2 + cap * WORDS, dtype=torch.int64, device=dev)
    fn = _cabi.lib.avsr_trace_set
    fn.restype, fn.argtypes = C.c_int, [C.c_void_p, C.c_size_t, C.c_void_p]
    _cabi.check(fn(buf.data_ptr(), buf.numel(), None))
    with torch.no_grad():
        for _ in range(5):                      # plan creation + warm-up (traced, then discarded)
            enc(xs, mask)
        torch.cuda.synchronize()
        buf[0] = 0                              # records used
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            enc(xs, mask)
        e1.record()
    torch.cuda.synchronize()
    ms_traced = e0.elapsed_time(e1) / steps
    host = buf.cpu().numpy().astype(np.uint64)
    _cabi.check(fn(None, 0, None))
    n = int(min(host[0], host[1]))
    rec = host[2:2 + n * WORDS].reshape(n, WORDS).astype(np.int64)
    lines = []
    P = lines.append
    P(f"workload {workload} lengths={lengths}: {n} CTA records over {steps} traced forward(s); "
      f"{ms_traced:.3f} ms/forward WITH tracing on (the untraced product build is faster)")

    # ---- split the records of every kernel kind into launches: CTAs of one launch open before the next launch of
    # the same kind (same-kind launches are >= one other kernel apart), so sort by open time and cut by grid size
    kinds = {}
    for r in rec:
        kinds.setdefault((int(r[0]), int(r[1] >> 32) if r[0] != 300 else 0), []).append(r)
    launches = []   # (start, dep, end, name, nctas)
    for (kid, aux), rows in kinds.items():
        rows = np.stack(rows)
        rows = rows[np.argsort(rows[:, 3], kind="stable")]
        if kid in (300, 120):                     # 3-D grids: one launch per layer
            nl = 12 * steps
        else:
            nl = int((rows[:, 1] & 0xffffffff == 0).sum())
        if nl == 0 or len(rows) % nl:
            P(f"!! {kind_name(kid, aux)}: {len(rows)} records do not split into {nl} launches")
            continue
        g = len(rows) // nl
        for i in range(nl):
            ch = rows[i * g:(i + 1) * g]
            dep = ch[:, 14][ch[:, 14] > 0]
            end = ch[:, 15][ch[:, 15] > 0]
            launches.append((int(ch[:, 3].min()), int(dep.min()) if len(dep) else int(ch[:, 3].min()),
                             int(end.max()) if len(end) else int(ch[:, 3].max()), kind_name(kid, aux), g))
    launches.sort(key=lambda t: t[2])
    t0 = launches[0][0]
    per_kind = {}
    P("")
    P("timeline of the LAST traced forward (us, relative to its first CTA):")
    P(f"{'kernel':<44s} {'ctas':>5s} {'open':>8s} {'dep':>8s} {'end':>8s} {'dep->end':>9s} {'exposed':>8s}")
    fw_start_idx = [i for i, l in enumerate(launches) if l[3].startswith("embed_scale")]
    last_fw = fw_start_idx[-1] if fw_start_idx else 0
    prev_end = None
    for i, (st, dep, end, name, g) in enumerate(launches):
        exposed = (end - prev_end) if prev_end is not None else (end - st)
        if name.startswith("embed_scale"):
            exposed = end - st                   # first kernel of a forward: the gap before it is host time
        prev_end = end
        d = per_kind.setdefault(name, [0, 0.0, 0.0])
        d[0] += 1; d[1] += (end - dep) / 1e3; d[2] += exposed / 1e3
        if i >= last_fw:
            base = launches[last_fw][0]
            P(f"{name:<44s} {g:>5d} {(st - base) / 1e3:>8.2f} {(dep - base) / 1e3:>8.2f} {(end - base) / 1e3:>8.2f} "
              f"{(end - dep) / 1e3:>9.2f} {exposed / 1e3:>8.2f}")
    P("")
    P(f"per kernel kind, per forward (mean over {steps} forwards): launches, dep->end sum, exposed sum (us)")
    tot = 0.0
    for name, (cnt, de, ex) in sorted(per_kind.items(), key=lambda kv: -kv[1][2]):
        P(f"  {name:<44s} n={cnt / steps:>5.1f}  dep->end {de / steps:>8.1f}  exposed {ex / steps:>8.1f}  "
          f"({ex / cnt:>6.2f} us / launch)")
        tot += ex / steps
    P(f"  {'sum of exposed':<44s} {tot:>8.1f} us / forward")

    # ---- per-CTA residency of the attention kernel: how long a CTA lives and how many share an SM at a time
    att = rec[rec[:, 0] == 300]
    if len(att):
        life = (att[:, 15] - att[:, 3]) / 1e3
        work = (att[:, 15] - att[:, 14]) / 1e3
        P("")
        P(f"attention_f16 CTAs: open->done median {np.median(life):.2f} us (p90 {np.percentile(life, 90):.2f}), "
          f"dep->done median {np.median(work):.2f} us (p90 {np.percentile(work, 90):.2f})")
        g = len(att) // (12 * steps)
        one = att[np.argsort(att[:, 3], kind="stable")][:g]          # first launch
        conc = []
        for r in one:
            same = one[one[:, 2] == r[2]]
            conc.append(int(((same[:, 3] < r[15]) & (same[:, 15] > r[3])).sum()))
        P(f"  first launch: {g} CTAs on {len(set(one[:, 2].tolist()))} SMs; CTAs alive on the same SM during a CTA's life: "
          f"median {int(np.median(conc))}, max {max(conc)}")

    # ---- intra-CTA phases (clock64 deltas)
    keys = sorted({(int(r[0]), int(r[1] >> 32)) for r in rec if r[0] in (300,) or 200 <= r[0] < 300})
    for kid, aux in keys:
        sel = rec[(rec[:, 0] == kid) & ((rec[:, 1] >> 32) == aux)]
        marks = sel[:, 4:14]
        names = ATT_MARKS if kid == 300 else GEMM_MARKS
        label = (f"attention_f16 key_tiles={aux & 0xff} qtile={aux >> 8}" if kid == 300 else kind_name(kid, aux))
        P(f"\n{label}: {len(sel)} CTAs")
        prev = None
        for s, name in enumerate(names):
            col = marks[:, s]
            ok = col > 0
            if prev is not None:
                both = ok & (marks[:, prev] > 0)
                if both.any():
                    d = col[both] - marks[both, prev]
                    P(f"  {names[prev]:>14s} -> {name:<14s} median {int(np.median(d)):>7d}  p90 {int(np.percentile(d, 90)):>7d} cycles"
                      f"  ({both.sum()} CTAs)")
            if ok.any():
                prev = s
    if brief:
        a = next(i for i, l in enumerate(lines) if l.startswith("per kernel kind"))
        b = next(i for i, l in enumerate(lines) if "sum of exposed" in l)
        lines = lines[:1] + lines[a:b + 1]
    text = "\n".join(lines)
    print(text)
    if out_path:
        os.makedirs(os.path.dirname(os.path.abspath(out_path)), exist_ok=True)
        with open(out_path, "w") as f:
            f.write(text + "\n")


if __name__ == "__main__":
    main()
