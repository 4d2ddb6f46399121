"""Helpers the evaluation drivers import from ``packages.utils`` (count_parameters: packages/utils.py:1-2, get_key: 5-8)."""
from __future__ import annotations


def count_parameters(model) -> int:
    """Number of trainable scalars of a ``torch.nn.Module`` (what the drivers print before enhancing)."""
    total = 0
    for tensor in model.parameters():
        if tensor.requires_grad:
            total += int(tensor.numel())
    return total


def get_key(mapping, wanted):
    """First key of ``mapping`` whose value equals ``wanted``; ``None`` when there is none (reverse lookup of label maps)."""
    return next((k for k, v in mapping.items() if v == wanted), None)
