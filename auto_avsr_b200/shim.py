"""Lightning-free replay of the reference's callers of the encoder (SURVEY.md D4, 8d "config 1"): the two code paths
that drive ``ConformerEncoder`` in mpc001/auto_avsr, restated line by line WITHOUT pytorch_lightning / the front-ends /
the beam search (which stay reference PyTorch and are not on the GPU box), so that the drop-in is exercised exactly the
way the reference exercises it:

* ``E2EShell``       -- the members ``E2E.__init__`` builds that those paths touch (e2e_asr_conformer.py:31-56), built
                        from the drop-in classes the way ``auto_avsr_b200.install()`` / ``install_head()`` arrange it;
* ``e2e_forward_encoder`` -- ``E2E.forward`` up to the encoder output (e2e_asr_conformer.py:63-71): padding mask from
                        ``lengths`` via make_non_pad_mask, proj_encoder, ``encoder(x, padding_mask)``;
* ``test_step_encoder``   -- ``ModelModule.test_step`` / ``forward`` up to the beam search (lightning.py:58,69-72): one
                        utterance, B = 1, ``encoder(x, None)``, a new T for every call;
* ``get_beam_search_decoder`` / ``test_step_decode`` -- the rest of ``test_step`` (lightning.py:73-76, :126-157):
                        the search over ``model.scorers()`` (decoder 0.9 + CTC prefix scorer 0.1, beam 40) and the predicted
                        token ids.
Inputs are front-end features ((T, 512) per utterance): what ``self.model.frontend`` returns."""
from __future__ import annotations

from typing import Sequence

import torch

from .espnet_dropin import CTC, ConformerEncoder, CTCPrefixScorer, ProjEncoder, TransformerDecoder


class E2EShell(torch.nn.Module):
    def __init__(self, odim: int = 5049, idim: int = 512, adim: int = 768, aheads: int = 12, eunits: int = 3072, elayers: int = 12,
                 dropout_rate: float = 0.1, cnn_module_kernel: int = 31):
        super().__init__()
        self.proj_encoder = ProjEncoder(idim, adim)                                   # e2e_asr_conformer.py:31
        self.encoder = ConformerEncoder(attention_dim=adim, attention_heads=aheads, linear_units=eunits,
                                        num_blocks=elayers, cnn_module_kernel=cnn_module_kernel)   # :33-39 (5 kwargs)
        self.decoder = TransformerDecoder(odim=odim, attention_dim=adim, attention_heads=aheads, linear_units=eunits,
                                          num_blocks=6)                                # :41-47
        self.sos = self.eos = odim - 1                                                 # :50-51
        self.odim = odim
        self.ctc = CTC(odim, adim, dropout_rate, reduce=True)                          # :56

    def scorers(self):
        """e2e_asr_conformer.py:58-59"""
        return dict(decoder=self.decoder, ctc=CTCPrefixScorer(self.ctc, self.eos))


def make_non_pad_mask(lengths: Sequence[int], device) -> torch.Tensor:
    """nets_utils.py:183-269 for a 1-D length list: True = valid frame."""
    ln = torch.as_tensor(list(lengths), device=device)
    return torch.arange(int(ln.max()), device=device)[None, :] < ln[:, None]


def e2e_forward_encoder(model: E2EShell, feats: torch.Tensor, lengths: Sequence[int]):
    """e2e_asr_conformer.py:67-71 (after the front-end): returns (encoder features (B,T,adim), padding mask)."""
    padding_mask = make_non_pad_mask(lengths, feats.device).unsqueeze(-2)
    x = model.proj_encoder(feats)
    x, _ = model.encoder(x, padding_mask)
    return x, padding_mask


def test_step_encoder(model: E2EShell, feats_one: torch.Tensor) -> torch.Tensor:
    """lightning.py:70-73 (after the front-end, before the beam search): (T, idim) -> enc_feat (T, adim)."""
    x = model.proj_encoder(feats_one.unsqueeze(0))
    enc_feat, _ = model.encoder(x, None)
    return enc_feat.squeeze(0)


def get_beam_search_decoder(model, beam_size: int = 40, ctc_weight: float = 0.1, reference_loop=None):
    """lightning.py:126-157 (no LM, length penalty 0).  Default: the device-resident loop (``DeviceBeamSearch``) over
    ``model.decoder`` / ``model.ctc``; pass the reference's ``BatchBeamSearch`` class as ``reference_loop`` to have the
    reference's own loop drive ``model.scorers()`` exactly as upstream does."""
    if reference_loop is None:
        from .beam_search import DeviceBeamSearch
        return DeviceBeamSearch(model.decoder, model.ctc, beam_size=beam_size, vocab_size=model.odim, sos=model.sos, eos=model.eos,
                                ctc_weight=ctc_weight)
    from .espnet_dropin import scorer_interface
    scorer_interface.rebind()
    scorers = model.scorers()
    scorers["lm"] = None
    weights = {"decoder": 1.0 - ctc_weight, "ctc": ctc_weight, "lm": 0.0, "length_bonus": 0}
    token_list = [str(i) for i in range(model.odim)]
    return reference_loop(beam_size=beam_size, vocab_size=model.odim, weights=weights, scorers=scorers, sos=model.sos, eos=model.eos,
                          token_list=token_list, pre_beam_score_key=None if ctc_weight == 1.0 else "decoder")


def test_step_decode(model: E2EShell, feats_one: torch.Tensor, beam_search) -> torch.Tensor:
    """lightning.py:70-76 after the front-end: (T, idim) features -> predicted token ids (``yseq[1:]`` of the best hypothesis:
    what ``text_transform.post_process`` receives)."""
    enc_feat = test_step_encoder(model, feats_one)
    nbest_hyps = beam_search(enc_feat)
    nbest_hyps = [h.asdict() for h in nbest_hyps[: min(len(nbest_hyps), 1)]]
    return torch.tensor(list(map(int, nbest_hyps[0]["yseq"][1:])))
