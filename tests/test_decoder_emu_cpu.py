"""CPU checks of the decoder / CTC-prefix path without a GPU: the launch schedule, buffer layouts,
slot / ancestor addressing and every per-element functor of csrc/decoder_body.cuh are compiled against the host backend
of tests/emu (same source, plain loops, naive fp32 GEMM / LayerNorm / log-softmax in place of the GPU-verified
launchers) and driven through the package's own host code (auto_avsr_b200.decoder) with the SAME ctypes signatures.
Compared with the reference-generated fixtures and the oracle.  The GPU parity proper is tests/test_zzz_gpu_decoder.py."""
import shutil

import pytest
import torch

from helpers import err_stats, load_decoder_case
from oracle import decoder_oracle as DO
from oracle import head_oracle as HO

pytestmark = pytest.mark.skipif(shutil.which("nvcc") is None and not __import__("os").path.exists("/usr/local/cuda/bin/nvcc"),
                                reason="nvcc is needed to build the host replay")


@pytest.fixture(scope="module")
def emu():
    from emu import build
    return build.load()


class _Params(torch.nn.Module):
    """a bag of parameters under the reference decoder's names"""

    def __init__(self, sd):
        super().__init__()
        self._names = {}
        for k, v in sd.items():
            name = k.replace(".", "__")
            self.register_parameter(name, torch.nn.Parameter(v.clone(), requires_grad=False))
            self._names[name] = k

    def named_parameters(self, *a, **kw):
        for name, p in super().named_parameters(*a, **kw):
            yield self._names[name], p


def _engine(c, emu):
    from auto_avsr_b200.decoder import DecoderEngine
    cfg = c["cfg"]
    eng = DecoderEngine(cfg["odim"], cfg["d_model"], cfg["n_heads"], cfg["linear_units"], cfg["num_blocks"], _lib=emu)
    return eng, _Params(c["dec_sd"])


def test_product_engine_refuses_cpu_tensors():
    from auto_avsr_b200.decoder import CtcPrefixEngine, DecoderEngine
    c = load_decoder_case("decoder_tiny")
    cfg = c["cfg"]
    eng = DecoderEngine(cfg["odim"], cfg["d_model"], cfg["n_heads"], cfg["linear_units"], cfg["num_blocks"])
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        eng.begin(_Params(c["dec_sd"]), c["memory"], 4, precision="fp32")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        CtcPrefixEngine(torch.zeros(5, 7), 0, 6)


@pytest.mark.parametrize("name", ["decoder_tiny", "decoder_full"])
def test_emu_decoder_steps_match_reference_fixture(emu, name):
    c = load_decoder_case(name)
    z, cfg = c["z"], c["cfg"]
    eng, params = _engine(c, emu)
    n = cfg["n_hyp"]
    eng.begin(params, c["memory"], max_hyps=n, precision="fp32")
    cols = torch.from_numpy(z["cols"])
    for step in range(cfg["steps"]):
        ys = c["prefixes"][step]
        anc = None if step == 0 else torch.arange(n, dtype=torch.int32).repeat(step, 1)       # every hypothesis its own lane
        logp = eng.step(ys[:, -1].to(torch.int32), anc, step)
        mx, _ = err_stats(logp[:, cols], torch.from_numpy(z[f"dec_logp_f32_{step}"]))
        assert mx < 2e-4, (step, mx)
        mx64, _ = err_stats(logp[:, cols], torch.from_numpy(z[f"dec_logp_f64_{step}"]))
        assert mx64 < 2e-4, (step, mx64)
    assert eng.stats == {"begin": 1, "step": cfg["steps"], "prepare": 1}


def test_emu_decoder_follows_reordered_and_forked_beams(emu):
    """Hypotheses change lanes and fork between steps (what the beam search does): the slot chains must address the
    right ancestors' K/V.  Oracle = full recomputation of every prefix."""
    c = load_decoder_case("decoder_tiny")
    cfg = c["cfg"]
    eng, params = _engine(c, emu)
    odim, sos = cfg["odim"], cfg["odim"] - 1
    mem = c["memory"]
    eng.begin(params, mem, max_hyps=5, max_steps=6, precision="fp32")
    g = torch.Generator().manual_seed(5)
    # step 0: the single <sos> hypothesis in lane 0
    prefixes = [[sos]]
    chains = [[]]                       # slots of positions < step
    eng.step(torch.tensor([sos], dtype=torch.int32), None, 0)
    for step in range(1, 5):
        n_prev = len(prefixes)
        n = min(5, n_prev + 2)
        parents = torch.randint(0, n_prev, (n,), generator=g).tolist()           # forks: several children of one parent
        toks = torch.randint(1, odim - 1, (n,), generator=g).tolist()
        new_prefixes = [prefixes[p] + [t] for p, t in zip(parents, toks)]
        new_chains = [chains[p] + [p] for p in parents]                            # the parent sat in lane p at step - 1
        anc = torch.tensor(new_chains, dtype=torch.int32).T.contiguous()           # (step, n)
        logp = eng.step(torch.tensor(toks, dtype=torch.int32), anc, step)
        ref = DO.decoder_logp(c["dec_sd"], torch.tensor(new_prefixes), mem.double(), cfg["n_heads"])
        mx, _ = err_stats(logp, ref)
        assert mx < 2e-4, (step, mx)
        prefixes, chains = new_prefixes, new_chains


def test_emu_decoder_argument_errors(emu):
    c = load_decoder_case("decoder_tiny")
    eng, params = _engine(c, emu)
    with pytest.raises(RuntimeError, match="before begin"):
        eng.step(torch.zeros(1, dtype=torch.int32), None, 0)
    eng.begin(params, c["memory"], max_hyps=2, max_steps=3, precision="fp32")
    with pytest.raises(ValueError, match="beam slots"):
        eng.step(torch.zeros(3, dtype=torch.int32), None, 0)
    with pytest.raises(ValueError, match="ancestor table"):
        eng.step(torch.zeros(2, dtype=torch.int32), None, 1)
    with pytest.raises(ValueError, match="positions"):
        eng.step(torch.zeros(2, dtype=torch.int32), torch.zeros(3, 2, dtype=torch.int32), 3)
    with pytest.raises(TypeError):
        eng.step(torch.zeros(2, dtype=torch.int64), None, 0)
    with pytest.raises(ValueError, match="memory must be"):
        eng.begin(params, c["memory"][:, :5], max_hyps=2, precision="fp32")


@pytest.mark.parametrize("name", ["decoder_tiny", "decoder_full"])
def test_emu_ctc_prefix_matches_reference_fixture(emu, name):
    from auto_avsr_b200.decoder import CtcPrefixEngine
    c = load_decoder_case(name)
    z, cfg = c["z"], c["cfg"]
    eos, n = cfg["odim"] - 1, cfg["n_hyp"]
    logp = HO.ctc_log_softmax(c["memory"].float(), c["head_sd"])
    eng = CtcPrefixEngine(logp, 0, eos, _lib=emu)
    r_prev, s_prev = eng.initial(n)
    for step in range(cfg["steps"]):
        ys = c["prefixes"][step]
        cand = torch.from_numpy(z[f"ctc_cand_{step}"]).to(torch.int32)
        local, r, log_psi = eng.score(step, ys[:, -1].to(torch.int32), r_prev, s_prev, cand)
        got = torch.gather(local, 1, cand.long())
        want = torch.from_numpy(z[f"ctc_local_f64_{step}"])
        live = (want > DO.LOGZERO / 2) & (want < -DO.LOGZERO / 2)
        assert (((got <= DO.LOGZERO / 2) | (got >= -DO.LOGZERO / 2)) == ~live).all()
        mx, _ = err_stats(got[live], want[live])
        assert mx < 2e-3 * max(1.0, float(want[live].abs().max()) / 50), (name, step, mx)
        ok = s_prev > DO.LOGZERO / 2        # (a prefix the fixture extended by blank / eos carries s_prev = logzero: 1e10-scale fp32)
        mx_eos, _ = err_stats(local[ok, eos], torch.from_numpy(z[f"ctc_eos_f64_{step}"])[ok])
        assert mx_eos < 2e-3, (name, step, mx_eos)
        assert (local[:, 0] - (DO.LOGZERO - s_prev) == 0).all()                         # blank
        if f"ctc_keep_{step}" in z.files:
            keep = torch.from_numpy(z[f"ctc_keep_{step}"]).to(torch.int32)
            r_prev, s_prev = eng.select(r, log_psi, cand, torch.arange(n, dtype=torch.int32), keep)
            # the oracle's selection of the same states
            pos = (cand == keep[:, None]).int().argmax(1)
            want_r = torch.stack([r[:, :, i, int(pos[i])] for i in range(n)], dim=2)
            assert torch.equal(r_prev, want_r)
            assert torch.equal(s_prev, log_psi[torch.arange(n), keep.long()])


def test_emu_ctc_prefix_against_oracle_random(emu):
    """random candidates incl. blank / eos / repeats of the last label, T = 1 and out_len > 0, vs the fp64 oracle"""
    from auto_avsr_b200.decoder import CtcPrefixEngine
    g = torch.Generator().manual_seed(11)
    for T, O, n, S in ((1, 9, 2, 4), (7, 12, 3, 12), (19, 30, 4, 7)):
        logp = torch.log_softmax(torch.randn(T, O, generator=g, dtype=torch.float64) * 2, -1)
        eng = CtcPrefixEngine(logp.float(), 0, O - 1, _lib=emu)
        r64, s64 = DO.ctc_initial_state(logp)
        r64, s64 = r64.expand(-1, -1, n).clone(), s64.expand(n).clone()
        r32, s32 = eng.initial(n)
        last = [O - 1] * n
        for out_len in range(0, min(4, T)):          # (a prefix longer than T frames has probability logzero: degenerate)
            cand = torch.stack([torch.randperm(O, generator=g)[:S] for _ in range(n)])
            cand[0, 0] = last[0] if out_len else cand[0, 0]                       # a repeat of the last label
            if len(set(cand[0].tolist())) < S:
                cand[0] = torch.randperm(O, generator=g)[:S]
            want, r_new, psi = DO.ctc_prefix_scores(logp, out_len, last, r64, s64, cand, 0, O - 1)
            got, r, log_psi = eng.score(out_len, torch.tensor(last, dtype=torch.int32), r32, s32, cand.to(torch.int32))
            live = want > DO.LOGZERO / 2
            assert ((got > DO.LOGZERO / 2) == live).all(), (T, out_len)
            assert err_stats(got[live], want[live])[0] < 1e-4, (T, out_len)
            # continue with each hypothesis' first non-special candidate
            keep = []
            for i in range(n):
                ok = [t for t in cand[i].tolist() if t not in (0, O - 1)]
                keep.append(ok[0])
            pos = [cand[i].tolist().index(keep[i]) for i in range(n)]
            r64 = torch.stack([r_new[:, :, i, pos[i]] for i in range(n)], dim=2)
            s64 = psi[torch.arange(n), torch.tensor(keep)]
            r32, s32 = eng.select(r, log_psi, cand.to(torch.int32), torch.arange(n, dtype=torch.int32),
                                  torch.tensor(keep, dtype=torch.int32))
            last = keep


def test_emu_decoder_long_memory_is_projected_in_row_chunks(emu):
    """T > 192: avsr_decoder_begin projects the source K|V in row chunks; every frame must land in its row."""
    c = load_decoder_case("decoder_tiny")
    cfg = c["cfg"]
    eng, params = _engine(c, emu)
    g = torch.Generator().manual_seed(21)
    mem = torch.randn(401, cfg["d_model"], generator=g)
    eng.begin(params, mem, max_hyps=2, max_steps=2, precision="fp32")
    sos = cfg["odim"] - 1
    logp = eng.step(torch.tensor([sos, sos], dtype=torch.int32), None, 0)
    ref = DO.decoder_logp(c["dec_sd"], torch.tensor([[sos], [sos]]), mem.double(), cfg["n_heads"])
    assert err_stats(logp, ref)[0] < 2e-4


def test_emu_ctc_select_of_a_token_off_the_candidate_list(emu):
    """A kept token that was not a candidate (only possible when fewer than `beam` candidates are live) carries logzero
    forward variables and the logzero prefix score, like the reference's select_state on log_psi."""
    from auto_avsr_b200.decoder import CtcPrefixEngine
    g = torch.Generator().manual_seed(2)
    logp = torch.log_softmax(torch.randn(6, 9, generator=g), -1)
    eng = CtcPrefixEngine(logp, 0, 8, _lib=emu)
    r0, s0 = eng.initial(2)
    cand = torch.tensor([[1, 2, 3], [4, 5, 6]], dtype=torch.int32)
    local, r, log_psi = eng.score(0, torch.tensor([8, 8], dtype=torch.int32), r0, s0, cand)
    r_next, s_next = eng.select(r, log_psi, cand, torch.tensor([0, 1, 1], dtype=torch.int32), torch.tensor([2, 7, 5], dtype=torch.int32))
    assert torch.equal(r_next[:, :, 0], r[:, :, 0, 1]) and torch.equal(r_next[:, :, 2], r[:, :, 1, 1])
    assert (r_next[:, :, 1] == DO.LOGZERO).all() and float(s_next[1]) == DO.LOGZERO
    assert float(s_next[0]) == float(log_psi[0, 2]) and float(s_next[2]) == float(log_psi[1, 5])


@pytest.mark.parametrize("d,H,ff,L,odim", [(128, 4, 192, 1, 19), (192, 3, 64, 3, 70), (64, 1, 128, 2, 5)])
def test_emu_decoder_other_geometries_against_oracle(emu, d, H, ff, L, odim):
    """head sizes other than 64 (32, 64, 64 with 3 heads), odd layer counts, tiny vocabularies padded to 64 output rows"""
    from auto_avsr_b200.decoder import DecoderEngine
    from auto_avsr_b200.synthetic import decoder_state_dict, encoder_input
    sd = decoder_state_dict(7, odim, d, H, ff, L)
    mem = encoder_input([11], d, 8)[0]
    eng = DecoderEngine(odim, d, H, ff, L, _lib=emu)
    params = _Params(sd)
    eng.begin(params, mem, max_hyps=3, max_steps=4, precision="fp32")
    sos = odim - 1
    prefixes, chains = [[sos], [sos], [sos]], [[], [], []]
    eng.step(torch.tensor([sos] * 3, dtype=torch.int32), None, 0)
    for step in range(1, 4):
        toks = [(step * 3 + i) % (odim - 1) for i in range(3)]
        parents = [(i + step) % 3 for i in range(3)]                   # the hypotheses change lanes every step
        prefixes = [prefixes[p] + [t] for p, t in zip(parents, toks)]
        chains = [chains[p] + [p] for p in parents]
        anc = torch.tensor(chains, dtype=torch.int32).T.contiguous()
        logp = eng.step(torch.tensor(toks, dtype=torch.int32), anc, step)
        ref = DO.decoder_logp(sd, torch.tensor(prefixes), mem.double(), H)
        assert err_stats(logp, ref)[0] < 2e-4, (step, d, H)
