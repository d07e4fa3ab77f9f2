"""Text conditioning: T5 encoder wrapper (reference: sam_audio/model/text_encoder.py:11-37).

T5-base is a third-party model (HF ``transformers``) that runs once per
separate() call on a handful of tokens; it stays a PyTorch module feeding
``sab_prepare`` (SURVEY.md §8a a5 / §8f-3 "next" row).  When no ``t5-base``
checkpoint/tokenizer is on disk (this sandbox has no network) the same module
graph is built from the t5-base *shape* with seeded random weights and a
deterministic hash tokenizer, so that benchmarks pay the real encoder cost on
synthetic text.
"""
from __future__ import annotations

from typing import List, Tuple

import torch

from .config import T5EncoderConfig


class _HashTokenizer:
    """Whitespace tokens -> stable ids in [3, vocab); id 1 = </s>, id 0 = pad (T5 conventions)."""

    def __init__(self, vocab_size: int = 32128):
        self.vocab_size = vocab_size

    def __call__(self, texts: List[str], truncation=True, max_length=512, padding="longest", return_tensors="pt"):
        rows = []
        for t in texts:
            ids = []
            for tok in t.split():
                h = 0
                for ch in tok:
                    h = (h * 131 + ord(ch)) % 1000003
                ids.append(3 + h % (self.vocab_size - 3))
            ids = ids[: (max_length or 512) - 1] + [1]
            rows.append(ids)
        width = max(len(r) for r in rows)
        input_ids = torch.zeros(len(rows), width, dtype=torch.long)
        mask = torch.zeros(len(rows), width, dtype=torch.long)
        for i, r in enumerate(rows):
            input_ids[i, : len(r)] = torch.tensor(r)
            mask[i, : len(r)] = 1
        return {"input_ids": input_ids, "attention_mask": mask}


class T5TextEncoder(torch.nn.Module):
    def __init__(self, cfg: T5EncoderConfig, allow_random_init: bool = False, seed: int = 0):
        super().__init__()
        import transformers
        self.pad_mode = cfg.pad_mode
        self.max_length = cfg.max_length
        self.random_init = False
        try:
            self.model = transformers.T5EncoderModel.from_pretrained(cfg.name, local_files_only=True)
            self.tokenizer = transformers.AutoTokenizer.from_pretrained(cfg.name, local_files_only=True)
        except Exception as exc:  # no checkpoint on disk
            if not allow_random_init:
                raise RuntimeError(
                    f"text encoder '{cfg.name}' is not available locally ({type(exc).__name__}); "
                    "pass allow_random_init=True for synthetic benchmarking") from exc
            t5 = transformers.T5Config(vocab_size=32128, d_model=cfg.dim, d_kv=64, d_ff=3072, num_layers=12,
                                       num_heads=12, relative_attention_num_buckets=32, dropout_rate=0.0,
                                       feed_forward_proj="relu")
            with torch.random.fork_rng(devices=[]):
                torch.manual_seed(seed)
                self.model = transformers.T5EncoderModel(t5)
            self.tokenizer = _HashTokenizer(t5.vocab_size)
            self.random_init = True
        self.model.eval()

    @torch.inference_mode()
    def forward(self, texts: List[str]) -> Tuple[torch.Tensor, torch.Tensor]:
        device = next(self.model.parameters()).device
        enc = self.tokenizer(texts, truncation=True, max_length=self.max_length, padding=self.pad_mode,
                             return_tensors="pt")
        input_ids = enc["input_ids"].to(device)
        attention_mask = enc["attention_mask"].to(device)
        hidden = self.model(input_ids=input_ids, attention_mask=attention_mask)["last_hidden_state"]
        return hidden, attention_mask.bool()


class SyntheticTextEncoder(torch.nn.Module):
    """Deterministic stand-in used by parity tests: the same features the oracle sees
    (sam_audio_b200.synthetic.synthetic_text_features), moved to the model's device."""

    def __init__(self, dim: int = 768):
        super().__init__()
        self.dim = dim
        self.register_buffer("_anchor", torch.zeros(1), persistent=False)

    def forward(self, texts: List[str]):
        from .synthetic import synthetic_text_features
        f, m = synthetic_text_features(texts, self.dim)
        return f.to(self._anchor.device), m.to(self._anchor.device)
