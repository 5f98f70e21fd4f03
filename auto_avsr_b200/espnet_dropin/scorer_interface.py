"""The scorer base classes the reference's beam search classifies its scorers by (espnet/nets/scorer_interface.py:
``isinstance(v, ScorerInterface)`` / ``PartialScorerInterface`` in BeamSearch.__init__, beam_search.py:84-96).

When the reference's ``espnet`` package is importable the drop-ins derive from ITS classes, so that its unmodified
``BatchBeamSearch`` accepts them; otherwise minimal stand-ins with the same method names are used (the package's own
``DeviceBeamSearch`` does not need them).  ``rebind()`` re-bases the drop-ins after the reference became importable."""
from typing import Any


def _reference_bases():
    try:
        from espnet.nets.scorer_interface import BatchPartialScorerInterface, BatchScorerInterface
        return BatchScorerInterface, BatchPartialScorerInterface
    except Exception:            # the reference is not on the path
        return None


class _StandInScorer:
    """the three defaults of ScorerInterface the beam search relies on (scorer_interface.py:27-80)"""

    def init_state(self, x) -> Any:
        return None

    def select_state(self, state: Any, i: int, new_id: int = None) -> Any:
        return None if state is None else state[i]

    def final_score(self, state: Any) -> float:
        return 0.0

    def batch_init_state(self, x) -> Any:
        return self.init_state(x)


class _StandInBatchScorer(_StandInScorer):
    pass


class _StandInBatchPartialScorer(_StandInScorer):
    pass


_ref = _reference_bases()
BatchScorerInterface, BatchPartialScorerInterface = _ref if _ref else (_StandInBatchScorer, _StandInBatchPartialScorer)


def rebind() -> bool:
    """Re-base TransformerDecoder / CTCPrefixScorer onto the reference's interfaces once ``espnet`` is importable
    (``install_decoder`` calls this).  -> True when the drop-ins now derive from the reference's classes."""
    global BatchScorerInterface, BatchPartialScorerInterface
    ref = _reference_bases()
    if ref is None:
        return False
    from . import scorers_ctc, transformer_decoder
    for cls, old, new in ((transformer_decoder.TransformerDecoder, BatchScorerInterface, ref[0]),
                          (scorers_ctc.CTCPrefixScorer, BatchPartialScorerInterface, ref[1])):
        if new not in cls.__mro__:
            cls.__bases__ = tuple(new if b is old else b for b in cls.__bases__)
    BatchScorerInterface, BatchPartialScorerInterface = ref
    return True
