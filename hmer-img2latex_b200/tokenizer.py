"""Id-level subset of the reference tokenizer (img2latex/data/tokenizer.py): the hot path
needs only the special ids, ``vocab_size`` and ``decode`` (ids -> LaTeX string)."""
from __future__ import annotations

from typing import Dict, Iterable, List, Optional

_DEFAULT_TOKENS = ["+", "-", "=", "(", ")", "[", "]", "{", "}", "\\frac", "\\sum", "\\int", "a", "b", "c", "x", "y",
                   "z", "0", "1", "2", "3", "4", "5", "6", "7", "8", "9", "\\alpha", "\\beta", "\\gamma", "\\delta",
                   "\\theta", "\\pi", "\\sigma", "\\mathbf", "\\mathrm", "\\mathcal", "\\limits", "_", "^", "\\infty"]


class LaTeXTokenizer:
    def __init__(self, special_tokens: Optional[Dict[str, str]] = None, max_sequence_length: Optional[int] = None):
        # tokenizer.py:35-47: PAD/START/END/UNK take ids 0..3 in dict order
        self.special_tokens = special_tokens or {"PAD": "<PAD>", "START": "<START>", "END": "<END>", "UNK": "<UNK>"}
        self.max_sequence_length = 150 if max_sequence_length is None else max_sequence_length
        self._init_special_tokens()

    def _init_special_tokens(self) -> None:                      # tokenizer.py:62-78
        self.token_to_id = {tok: i for i, tok in enumerate(self.special_tokens.values())}
        self.id_to_token = {i: tok for tok, i in self.token_to_id.items()}
        self.vocab_size = len(self.token_to_id)
        self.pad_token_id = self.token_to_id[self.special_tokens["PAD"]]
        self.start_token_id = self.token_to_id[self.special_tokens["START"]]
        self.end_token_id = self.token_to_id[self.special_tokens["END"]]
        self.unk_token_id = self.token_to_id[self.special_tokens["UNK"]]

    def add_tokens(self, tokens: Iterable[str]) -> None:
        for t in tokens:
            if t not in self.token_to_id:
                self.token_to_id[t] = self.vocab_size
                self.id_to_token[self.vocab_size] = t
                self.vocab_size += 1

    def default_init(self) -> None:                              # tokenizer.py:323-385 (46 tokens)
        self._init_special_tokens()
        self.add_tokens(_DEFAULT_TOKENS)

    @classmethod
    def from_config(cls, cfg: dict) -> "LaTeXTokenizer":
        """Rebuild from a checkpoint's tokenizer_config (training/predictor.py:86-105): the special ids are
        looked up in the stored ``token_to_id``."""
        t = cls(cfg.get("special_tokens"), cfg.get("max_sequence_length", 141))
        t.token_to_id = dict(cfg.get("token_to_id", {}))
        t.id_to_token = {int(i): tok for tok, i in t.token_to_id.items()}
        t.vocab_size = len(t.token_to_id)
        t.pad_token_id = t.token_to_id[t.special_tokens["PAD"]]
        t.start_token_id = t.token_to_id[t.special_tokens["START"]]
        t.end_token_id = t.token_to_id[t.special_tokens["END"]]
        t.unk_token_id = t.token_to_id[t.special_tokens["UNK"]]
        return t

    def encode(self, text: str, add_special_tokens: bool = False) -> List[int]:  # tokenizer.py:143-164
        if add_special_tokens:
            text = f"{self.special_tokens['START']} {text} {self.special_tokens['END']}"
        return [self.token_to_id.get(tok, self.unk_token_id) for tok in text.split()]

    def decode(self, ids: List[int], skip_special_tokens: bool = True) -> str:   # tokenizer.py:166-194
        special = set(self.token_to_id[t] for t in self.special_tokens.values()) if skip_special_tokens else set()
        toks = [self.id_to_token.get(i, self.special_tokens["UNK"]) for i in ids if i not in special]
        return " ".join(toks)
