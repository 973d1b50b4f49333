"""``Wav2Vec2Processor`` with the reference's surface (src/wav2vec2/processor.py:10-106): either a
feature extractor (per-utterance normalisation) or a character tokenizer with greedy-CTC decode.
Host-side pre/post-processing (CUDA tensors are normalised by the ``w2v2_normalize_utterances`` kernel); no network
access (a missing vocab raises instead of downloading)."""
import json
import os
import re
from itertools import groupby

import torch

_DEFAULT_VOCAB = os.path.join(os.path.dirname(os.path.abspath(__file__)), "vocab.json")


class Wav2Vec2Processor:
    def __init__(self, is_tokenizer, do_normalize=True, vocab_path=_DEFAULT_VOCAB):
        self.is_tokenizer = is_tokenizer
        self.do_normalize = do_normalize
        self.vocab_path = vocab_path
        if self.is_tokenizer:
            if not os.path.isfile(self.vocab_path):
                raise ValueError(f"Couldn't find `vocab.json` at {self.vocab_path} (no network download here)")
            self.token_to_id_mapping = self.get_vocab()
            self.id_to_token_mapping = {v: k for k, v in self.token_to_id_mapping.items()}
            self.unk_token = "<unk>"
            self.unk_id = self.token_to_id_mapping[self.unk_token]
            self.dimiliter_token = "|"
            self.dimiliter_id = self.token_to_id_mapping[self.dimiliter_token]
            self.special_ids = [self.token_to_id_mapping[k] for k in ["<pad>"]]

    def __call__(self, input_values):
        if self.is_tokenizer:
            return [self.token_to_id_mapping.get(k, self.unk_id) for k in self._tokenize(input_values)]
        return self._normalize(input_values) if self.do_normalize else input_values

    def decode(self, input_ids: list, skip_special_tokens=True, group_tokens=True):
        """processor.py:71-89: collapse repeats, drop <pad>, '|' -> ' '."""
        if torch.is_tensor(input_ids):
            input_ids = input_ids.tolist()
        if group_tokens:
            input_ids = [k for k, _ in groupby(input_ids)]
        if skip_special_tokens:
            input_ids = [k for k in input_ids if k not in self.special_ids]
        tokens = [self.id_to_token_mapping.get(k, self.unk_token) for k in input_ids]
        return "".join(" " if t == self.dimiliter_token else t for t in tokens).strip()

    def _tokenize(self, string: str):
        string = re.sub("-", " ", string)
        string = re.sub("[^A-Z' ]", "", string.upper())
        return list(string.replace(" ", self.dimiliter_token))

    def get_vocab(self):
        with open(self.vocab_path, "r") as fh:
            return json.load(fh)

    def _normalize(self, x):
        """(x - mean) / sqrt(var + 1e-5), biased variance, per utterance, before padding
        (processor.py:101-106)."""
        if torch.is_tensor(x) and x.is_cuda:
            # device path (w2v2_normalize_utterances): one utterance [L] or a batch [B, L] of full-length utterances
            from . import ops
            return torch.squeeze(ops.normalize_utterances(x.reshape(-1, x.shape[-1])).reshape(x.shape))
        x = torch.as_tensor(x, dtype=torch.float32)
        mean = x.mean(dim=-1, keepdim=True)
        var = x.var(dim=-1, unbiased=False, keepdim=True)
        return torch.squeeze((x - mean) / torch.sqrt(var + 1e-5))
