"""Mirror of bblackjax/base.py:76-103: the (init, step) pair returned by `mala(...)`."""
from typing import Callable, NamedTuple


class SamplingAlgorithm(NamedTuple):
    init: Callable
    step: Callable
