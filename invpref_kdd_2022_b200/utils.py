"""The few reference utilities the hot path depends on (reference ``utils.py``)."""
from __future__ import annotations

import numpy as np


def mini_batch(batch_size: int, *tensors):
    """utils.py:12-19: sequential, UNSHUFFLED slices ``[s*B, (s+1)*B)``; the last batch is short.

    The fixed slicing is what lets the trainer build each batch's sort-segment plan once."""
    n = len(tensors[0])
    single = len(tensors) == 1
    for lo in range(0, n, batch_size):
        if single:
            yield tensors[0][lo:lo + batch_size]
        else:
            yield tuple(t[lo:lo + batch_size] for t in tensors)


def merge_dict(dict_list: list, merge_func, **func_args) -> dict:
    """utils.py:166-178: merge a list of equally-keyed dicts key by key."""
    keys = dict_list[0].keys()
    for d in dict_list:
        assert keys == d.keys()
    return {k: merge_func([d[k] for d in dict_list], **func_args) for k in keys}


def _mean_merge_dict_func(elements_list, **args):
    """utils.py:181-183."""
    return np.mean(elements_list)


def _show_me_a_list_func(elements_list, **args):
    return elements_list


def transfer_loss_dict_to_line_str(loss_dict: dict) -> str:
    return " ".join(f"{k}: {v}" for k, v in loss_dict.items())
