"""Type aliases (mirror of bblackjax/types.py)."""
from typing import Any, Iterable, Mapping, Union

import torch

Array = torch.Tensor
PyTree = Union[Array, Iterable[Array], Mapping[Any, Array]]
PRNGKey = torch.Tensor  # uint32[2] (or uint32[N,2] for a batch of chains)
