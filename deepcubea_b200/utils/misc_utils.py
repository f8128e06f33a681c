"""Ragged-list plumbing of the reference's Python BWAS (utils/misc_utils.py:6-36)."""
from typing import Any, List, Tuple

import numpy as np


def flatten(data: List[List[Any]]) -> Tuple[List[Any], List[int]]:
    """Concatenate sub-lists; also return the split points that `unflatten` needs."""
    split_idxs = np.cumsum([len(x) for x in data])[:-1].tolist()
    return [item for sub in data for item in sub], split_idxs


def unflatten(data: List[Any], split_idxs: List[int]) -> List[List[Any]]:
    bounds = [0] + list(split_idxs) + [len(data)]
    return [data[bounds[i]:bounds[i + 1]] for i in range(len(bounds) - 1)]


def split_evenly(num_total: int, num_splits: int) -> List[int]:
    base, extra = divmod(num_total, num_splits)
    return [base + (1 if i < extra else 0) for i in range(num_splits)]
