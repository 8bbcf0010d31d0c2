"""Whitespace tokenizer shim with the CLIP tokenizer's call surface.

The CLIP vocabulary is not available offline; the reference only needs `tokenizer(text)['input_ids']` (with BOS/EOS)
and `tokenizer.decode(id)` (reference `run.py:81-91`, `pipeline_guided_attention.py:214, 1105-1106`).  Token ids are a
stable hash of the lower-cased word so every process of a seed sweep agrees without sharing state.
"""
from __future__ import annotations

import zlib
from typing import List, Union

import torch


class _Encoding(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:  # pragma: no cover
            raise AttributeError(k) from e


class WhitespaceTokenizer:
    bos_token_id = 49406
    eos_token_id = 49407
    model_max_length = 77

    def __init__(self):
        self._words = {self.bos_token_id: "<|startoftext|>", self.eos_token_id: "<|endoftext|>"}

    def _word_id(self, w: str) -> int:
        wid = 1000 + zlib.crc32(w.encode("utf-8")) % 48000
        self._words.setdefault(wid, w)
        return wid

    def _encode_one(self, text: str) -> List[int]:
        return [self.bos_token_id] + [self._word_id(w) for w in text.lower().split()] + [self.eos_token_id]

    def __call__(self, text: Union[str, List[str]], padding=None, max_length=None, truncation=False,
                 return_tensors=None, **_):
        single = isinstance(text, str)
        ids = [self._encode_one(t) for t in ([text] if single else text)]
        if truncation and max_length is not None:
            ids = [s[: max_length - 1] + [self.eos_token_id] if len(s) > max_length else s for s in ids]
        if padding == "max_length":
            tgt = max_length or self.model_max_length
            ids = [s + [self.eos_token_id] * (tgt - len(s)) for s in ids]
        elif padding == "longest":
            tgt = max(len(s) for s in ids)
            ids = [s + [self.eos_token_id] * (tgt - len(s)) for s in ids]
        mask = [[1] * len(s) for s in ids]
        if return_tensors == "pt":
            return _Encoding(input_ids=torch.tensor(ids, dtype=torch.long),
                             attention_mask=torch.tensor(mask, dtype=torch.long))
        if single:
            return _Encoding(input_ids=ids[0], attention_mask=mask[0])
        return _Encoding(input_ids=ids, attention_mask=mask)

    def encode(self, text: str) -> List[int]:
        return self._encode_one(text)

    def decode(self, token_id) -> str:
        if isinstance(token_id, (list, tuple)):
            return " ".join(self.decode(t) for t in token_id)
        return self._words.get(int(token_id), f"<{int(token_id)}>")

    def batch_decode(self, ids):
        return [self.decode(list(map(int, row))) for row in ids]
