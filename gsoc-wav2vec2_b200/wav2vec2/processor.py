"""``Wav2Vec2Processor`` with the reference's surface (src/wav2vec2/processor.py:10-106): one object is EITHER the feature
extractor (per-utterance normalisation of raw speech) OR the character tokenizer with greedy-CTC decoding.

Pre/post-processing around the hot path: text handling is host-side Python (pure functions below, the class only holds the
vocabulary), CUDA tensors are normalised by the ``w2v2_normalize_utterances`` kernel, host arrays on the host like the
reference.  No network access: the default vocabulary is built in, a missing ``vocab_path`` file raises instead of being downloaded.
"""
import json
import os
import re
from itertools import groupby
from typing import Dict, Iterable, List

import torch

# The character vocabulary of the public wav2vec2 CTC checkpoints (what the reference ships as data/vocab.json): four special
# tokens, the word delimiter, then the letters by corpus frequency and the apostrophe.  Built in; ``vocab_path`` overrides it.
_BUILTIN_VOCAB = ("<pad>", "<s>", "</s>", "<unk>", "|") + tuple("ETAONIHSRDLUMWCFGYPBVK") + ("'",) + tuple("XJQZ")
_NOT_IN_ALPHABET = re.compile(r"[^A-Z' ]")
PAD_TOKEN, UNK_TOKEN, WORD_DELIMITER = "<pad>", "<unk>", "|"


# ------------------------------------------------------------------------------------------ text <-> ids (pure functions)
def text_to_characters(text: str) -> List[str]:
    """processor.py:91-94: hyphens become spaces, upper-case, drop everything outside ``A-Z' ``, spaces -> ``|``."""
    cleaned = _NOT_IN_ALPHABET.sub("", text.replace("-", " ").upper())
    return [WORD_DELIMITER if ch == " " else ch for ch in cleaned]


def ids_to_text(ids: Iterable[int], id_to_token: Dict[int, str], drop_ids=(), collapse_repeats=True) -> str:
    """Greedy CTC decoding (processor.py:71-89): collapse runs of equal ids, drop ``drop_ids`` (the pad / blank id),
    map ids to characters, ``|`` -> space, strip."""
    ids = list(ids)
    if collapse_repeats:
        ids = [k for k, _run in groupby(ids)]
    chars = (id_to_token.get(k, UNK_TOKEN) for k in ids if k not in drop_ids)
    return "".join(" " if c == WORD_DELIMITER else c for c in chars).strip()


def normalize_speech(x):
    """processor.py:101-106: ``(x - mean) / sqrt(var + 1e-5)`` with the biased variance, per utterance, BEFORE padding."""
    if torch.is_tensor(x) and x.is_cuda:
        from . import ops                       # device path: one utterance [L] or a batch [B, L] of full-length utterances
        return torch.squeeze(ops.normalize_utterances(x.reshape(-1, x.shape[-1])).reshape(x.shape))
    x = torch.as_tensor(x, dtype=torch.float32)
    centred = x - x.mean(dim=-1, keepdim=True)
    return torch.squeeze(centred / torch.sqrt(x.var(dim=-1, unbiased=False, keepdim=True) + 1e-5))


# ------------------------------------------------------------------------------------------ the reference's class surface
class Wav2Vec2Processor:
    def __init__(self, is_tokenizer, do_normalize=True, vocab_path=None):
        self.is_tokenizer, self.do_normalize, self.vocab_path = is_tokenizer, do_normalize, vocab_path
        if not is_tokenizer:
            return
        if vocab_path is not None and not os.path.isfile(vocab_path):
            raise ValueError(f"Couldn't find `vocab.json` at {vocab_path} (no network download here)")
        self.token_to_id_mapping = self.get_vocab()
        self.id_to_token_mapping = {i: tok for tok, i in self.token_to_id_mapping.items()}
        self.unk_token, self.dimiliter_token = UNK_TOKEN, WORD_DELIMITER            # (sic) the reference's attribute names
        self.unk_id = self.token_to_id_mapping[UNK_TOKEN]
        self.dimiliter_id = self.token_to_id_mapping[WORD_DELIMITER]
        self.special_ids = [self.token_to_id_mapping[PAD_TOKEN]]

    def __call__(self, input_values):
        """Tokenizer: str -> list of ids.  Feature extractor: speech -> normalised speech (or unchanged)."""
        if self.is_tokenizer:
            return [self.token_to_id_mapping.get(ch, self.unk_id) for ch in self._tokenize(input_values)]
        return self._normalize(input_values) if self.do_normalize else input_values

    def decode(self, input_ids: list, skip_special_tokens=True, group_tokens=True):
        if torch.is_tensor(input_ids):
            input_ids = input_ids.tolist()
        return ids_to_text(input_ids, self.id_to_token_mapping, self.special_ids if skip_special_tokens else (), group_tokens)

    def get_vocab(self):
        """token -> id: the JSON file at ``vocab_path`` (same format as the reference's data/vocab.json) or the built-in set."""
        if self.vocab_path is None:
            return {tok: i for i, tok in enumerate(_BUILTIN_VOCAB)}
        with open(self.vocab_path, "r") as fh:
            return json.load(fh)

    _tokenize = staticmethod(text_to_characters)
    _normalize = staticmethod(normalize_speech)
