"""Translator (mirror of the reference's utils/translation.py)."""
from ..models.captioning.gnmt import BeamSearchScorer

__all__ = ['BeamSearchTranslator']


class BeamSearchTranslator(object):
    """Beam Search Translator (reference utils/translation.py:28-82).

    The reference wraps gluonnlp's BeamSearchSampler around `log_softmax(model.decode_step(...))`; here the whole
    sampler loop (scorer, top-k, back-pointers, state selection, termination) runs on the device in
    tn_gnmt_beam_search and returns the same triple."""

    def __init__(self, model, beam_size=1, scorer=None, max_length=100):
        self._model = model
        self._beam_size = beam_size
        self._scorer = scorer if scorer is not None else BeamSearchScorer()
        self._max_length = max_length
        vocab = model.tgt_vocab
        self._eos_id = vocab.token_to_idx[vocab.eos_token]
        self._bos_id = vocab.token_to_idx[vocab.bos_token]

    def translate(self, src_seq, src_valid_length):
        """-> samples (B,beam,L) int32, scores (B,beam) descending, valid_length (B,beam) int32."""
        encoder_outputs, _ = self._model.encode(src_seq, valid_length=src_valid_length)
        decoder_states = self._model.decoder.init_state_from_encoder(encoder_outputs, src_valid_length)
        return self._model.beam_search(decoder_states, self._beam_size, self._max_length, self._scorer._alpha,
                                       self._scorer._K, self._bos_id, self._eos_id)
