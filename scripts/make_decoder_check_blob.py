#!/usr/bin/env python3
"""Inputs + expected outputs for scripts/decoder_gpu_check.cu (a torch-free GPU check of the row 8f#3 kernels that
starts in seconds on a fresh box).  TEST INFRASTRUCTURE: uses the oracle; run in the build container:

    python scripts/make_decoder_check_blob.py        -> scripts/bin/decoder_check.blob

Case "tiny": the decoder_tiny fixture weights (in the blob), a forked / re-ordered beam over 5 steps, fp64 oracle logp.
Case "full": d 768 / 6 layers / odim 5049 / 40 hypotheses / T 50, weights from a 32-bit multiplicative hash that the
binary regenerates itself (only the expected outputs travel).  Case "ctc": prefix scorer steps on the tiny posteriors."""
import os
import struct
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from auto_avsr_b200._cabi import DECODER_LAYER_FIELDS  # noqa: E402
from helpers import load_decoder_case  # noqa: E402
from oracle import decoder_oracle as DO  # noqa: E402
from oracle import head_oracle as HO  # noqa: E402

OUT = os.path.join(ROOT, "scripts", "bin", "decoder_check.blob")


def hash_uniform(count, seed, scale, offset=0.0):
    """what decoder_gpu_check.cu's fill() produces: ((i * 2654435761 + seed * 40503) mod 2^32 >> 8) / 2^24 - 0.5, * scale + offset"""
    i = np.arange(count, dtype=np.uint64)
    u = ((i * np.uint64(2654435761) + np.uint64(seed * 40503)) & np.uint64(0xFFFFFFFF)) >> np.uint64(8)
    v = u.astype(np.float32) * np.float32(1.0 / 16777216.0) - np.float32(0.5)
    return v * np.float32(scale) + np.float32(offset)


def hashed_decoder_sd(odim, d, ff, L):
    """parameter order == the order decoder_gpu_check.cu fills them in (seed = running tensor index)"""
    sd, seed = {}, [1]

    def put(key, shape, scale, offset=0.0):
        n = int(np.prod(shape))
        sd[key] = torch.from_numpy(hash_uniform(n, seed[0], scale, offset).reshape(shape).copy())
        seed[0] += 1

    put("embed.0.weight", (odim, d), 2.0 / np.sqrt(d))
    for l in range(L):
        for _, suffix in DECODER_LAYER_FIELDS:
            key = f"decoders.{l}.{suffix}"
            if suffix.startswith("norm"):
                put(key, (d,), 1.0, 1.0) if suffix.endswith("weight") else put(key, (d,), 0.2)
            elif suffix.endswith("bias"):
                put(key, (ff if "w_1" in suffix else d,), 0.2)
            else:
                o, i = (ff, d) if "w_1" in suffix else ((d, ff) if "w_2" in suffix else (d, d))
                put(key, (o, i), 2.0 / np.sqrt(i))
    put("after_norm.weight", (d,), 1.0, 1.0)
    put("after_norm.bias", (d,), 0.2)
    put("output_layer.weight", (odim, d), 2.0 / np.sqrt(d))
    put("output_layer.bias", (odim,), 0.2)
    return sd, seed[0]


class Blob:
    def __init__(self):
        self.items = []

    def add(self, name, arr):
        arr = np.ascontiguousarray(arr)
        assert arr.dtype in (np.float32, np.int32), (name, arr.dtype)
        self.items.append((name, arr))

    def write(self, path):
        os.makedirs(os.path.dirname(path), exist_ok=True)
        with open(path, "wb") as f:
            f.write(struct.pack("<I", len(self.items)))
            for name, arr in self.items:
                nb = name.encode()
                f.write(struct.pack("<I", len(nb)) + nb)
                f.write(struct.pack("<II", 0 if arr.dtype == np.float32 else 1, arr.ndim))
                f.write(struct.pack("<" + "q" * arr.ndim, *arr.shape))
                f.write(arr.tobytes())


def beam_scenario(sd, memory, odim, n_heads, steps, max_n, seed, blob, tag):
    """a re-ordered, forking beam; per step: tokens, anc (step, n), expected logp (fp64 oracle)"""
    g = torch.Generator().manual_seed(seed)
    sos = odim - 1
    prefixes, chains = [[sos]], [[]]
    for step in range(steps):
        if step > 0:
            n_prev = len(prefixes)
            n = min(max_n, n_prev * 3)
            parents = torch.randint(0, n_prev, (n,), generator=g).tolist()
            toks = torch.randint(1, odim - 1, (n,), generator=g).tolist()
            chains = [chains[p] + [p] for p in parents]
            prefixes = [prefixes[p] + [t] for p, t in zip(parents, toks)]
        n = len(prefixes)
        ref = DO.decoder_logp(sd, torch.tensor(prefixes), memory.double(), n_heads)
        blob.add(f"{tag}.tokens{step}", np.asarray([p[-1] for p in prefixes], dtype=np.int32))
        blob.add(f"{tag}.anc{step}", np.asarray(chains, dtype=np.int32).T.reshape(step, n))
        blob.add(f"{tag}.logp{step}", ref.float().numpy())
        print(tag, "step", step, "n", n, "logp range", float(ref.min()), float(ref.max()))


def main():
    blob = Blob()
    # ---- tiny: fixture weights --------------------------------------------------------------------------------------
    c = load_decoder_case("decoder_tiny")
    cfg = c["cfg"]
    blob.add("tiny.cfg", np.asarray([cfg["d_model"], cfg["n_heads"], cfg["linear_units"], cfg["num_blocks"], cfg["odim"],
                                     cfg["T"], 5, 5], dtype=np.int32))            # ..., T, steps, max_hyps
    sd = c["dec_sd"]
    blob.add("tiny.embed", sd["embed.0.weight"].numpy())
    for l in range(cfg["num_blocks"]):
        for field, suffix in DECODER_LAYER_FIELDS:
            blob.add(f"tiny.l{l}.{field}", sd[f"decoders.{l}.{suffix}"].numpy())
    for k, name in (("after_norm.weight", "after_w"), ("after_norm.bias", "after_b"), ("output_layer.weight", "out_w"),
                    ("output_layer.bias", "out_b")):
        blob.add("tiny." + name, sd[k].numpy())
    blob.add("tiny.memory", c["memory"].numpy())
    beam_scenario(sd, c["memory"], cfg["odim"], cfg["n_heads"], 5, 5, 5, blob, "tiny")
    # ---- ctc: tiny posteriors, 3 chained steps -------------------------------------------------------------------------
    logp = HO.ctc_log_softmax(c["memory"].double(), c["head_sd"])
    T, O = logp.shape
    n, S = 4, 7
    g = torch.Generator().manual_seed(3)
    blob.add("ctc.cfg", np.asarray([T, O, n, S, 3], dtype=np.int32))
    blob.add("ctc.logp", logp.float().numpy())
    r, s = DO.ctc_initial_state(logp)
    r, s = r.expand(-1, -1, n).clone(), s.expand(n).clone()
    last = [O - 1] * n
    for step in range(3):
        cand = torch.stack([torch.randperm(O, generator=g)[:S] for _ in range(n)])
        local, r_new, psi = DO.ctc_prefix_scores(logp, step, last, r, s, cand, 0, O - 1)
        keep = [[t for t in cand[i].tolist() if t not in (0, O - 1)][0] for i in range(n)]
        blob.add(f"ctc.last{step}", np.asarray(last, dtype=np.int32))
        blob.add(f"ctc.cand{step}", cand.numpy().astype(np.int32))
        blob.add(f"ctc.local{step}", local.float().numpy())
        blob.add(f"ctc.keep{step}", np.asarray(keep, dtype=np.int32))
        pos = [cand[i].tolist().index(keep[i]) for i in range(n)]
        r = torch.stack([r_new[:, :, i, pos[i]] for i in range(n)], dim=2)
        s = psi[torch.arange(n), torch.tensor(keep)]
        last = keep
    # ---- full size: hashed weights (regenerated by the binary) ---------------------------------------------------------
    odim, d, H, ff, L, T, n = 5049, 768, 12, 3072, 6, 50, 40
    sd, next_seed = hashed_decoder_sd(odim, d, ff, L)
    memory = torch.from_numpy(hash_uniform(T * d, next_seed, 3.0).reshape(T, d).copy())
    blob.add("full.cfg", np.asarray([d, H, ff, L, odim, T, 5, n], dtype=np.int32))
    beam_scenario(sd, memory, odim, H, 5, n, 9, blob, "full")
    blob.write(OUT)
    print("wrote", OUT, os.path.getsize(OUT) // 1024, "KiB")


if __name__ == "__main__":
    main()
