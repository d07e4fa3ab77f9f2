"""Reranker call-site surface (reference: sam_audio/ranking/ranker.py:9-36, used at sam_audio/model/model.py:306-330).

The concrete rankers of the reference (CLAP, ImageBind, Judge) are third-party scoring models outside the
``separate()`` arithmetic and are not built here; what IS mirrored is the contract ``separate()`` relies on — a ranker is
a callable returning scores ``[batch, candidates]`` — and the ensemble combinator, so that any ranker module written for
the reference can be attached as ``model.text_ranker`` / ``model.visual_ranker`` and candidate selection behaves as in
the reference (argmax over the candidates of each clip).
"""
from __future__ import annotations

from abc import ABCMeta, abstractmethod
from typing import List

import torch


class Ranker(torch.nn.Module, metaclass=ABCMeta):
    @abstractmethod
    def forward(self, **kwargs) -> torch.Tensor:
        """kwargs as passed by separate(): ``extracted_audio`` (list over clips of [candidates, samples]),
        ``sample_rate``, and either ``videos`` (visual ranker) or ``input_audio`` + ``descriptions`` (text ranker).
        Returns scores [batch, candidates]."""


class EnsembleRanker(Ranker):
    """Weighted sum of the member rankers' scores (reference ranking/ranker.py:22-36)."""

    def __init__(self, rankers: List[torch.nn.Module], weights: List[float]):
        super().__init__()
        assert len(rankers) == len(weights)
        self.rankers = torch.nn.ModuleList(rankers)
        self.weights = weights

    def forward(self, **kwargs) -> torch.Tensor:
        result = None
        for weight, ranker in zip(self.weights, self.rankers):
            result = weight * ranker(**kwargs) if result is None else result + weight * ranker(**kwargs)
        return result
