"""Masked layers — drop-in for sparse_caption/pruning/masked_layer.py (MaskMixin 20-114, MaskedLinear 118-135,
MaskedEmbedding 139-174).  Same constructor signatures, attribute names (``weight``, ``bias``,
``weight_pruning_mask``, ``mask_type``, ``mask_init_value``, ``mask_trainable``, ``mask_parameters``) and state-dict keys.

Forward runs on the GPU only: the mask (binarized / Philox-Bernoulli / raw) is applied inside the GEMM operand-load
prologue (sc_linear) or inside the embedding gather (sc_embed); the masked weight is never materialised.
There is no CPU fallback: calling forward on CPU tensors raises.
"""
import logging
from copy import deepcopy
from typing import List, Tuple, Union

import torch
from torch import Tensor, nn
from torch.nn import init
from torch.nn.parameter import Parameter

from . import kernels as K
from . import prune, sampler
from . import autograd_ops as A

logger = logging.getLogger(__name__)

# global numeric mode of the layer-level API: "bf16" (tcgen05 tensor cores, 2e-2 parity) or "fp32" (1e-5 parity)
_precision = {"mode": "bf16"}


def set_precision(mode: str) -> None:
    assert mode in ("bf16", "fp32")
    _precision["mode"] = mode


def get_precision() -> str:
    return _precision["mode"]


# noinspection PyAttributeOutsideInit
class MaskMixin:
    mask_type: str
    mask_init_value: float
    mask_trainable: bool
    training: bool

    def setup_masks(self, parameters: Union[str, List[str], Tuple[str, ...]], mask_type: str,
                    mask_init_value: float = 1.0, bypass_sigmoid_grad: bool = False) -> None:
        names = (parameters,) if isinstance(parameters, str) else tuple(parameters)
        assert all(isinstance(n, str) for n in names)
        assert mask_type in prune.VALID_MASKS, f"`mask_type` must be one of {prune.VALID_MASKS}, saw `{mask_type}`"
        self.mask_type = mask_type
        self.bypass_sigmoid_grad = bool(bypass_sigmoid_grad)
        self.mask_parameters = []
        for name in names:
            weight = getattr(self, name, None)
            assert weight is not None, f"Invalid weight attribute name: {name}"
            if not isinstance(weight, Parameter):
                logger.warning(f"{type(self).__name__}: `{name}` is a {type(weight)}, converting it into a Parameter.")
                weight = Parameter(weight)
            setattr(self, f"{name}_pruning_mask", deepcopy(weight))
            self.mask_parameters.append(getattr(self, f"{name}_pruning_mask"))
        if mask_type in prune.SUPER_MASKS:
            assert isinstance(mask_init_value, (float, int)), "`mask_init_value` must be provided as a float or int."
            self.mask_init_value = float(mask_init_value)
            self.mask_trainable = True
            self.mask_train_sample_fn = lambda x: sampler.bernoulli_sample_sigmoid(x, bypass_sigmoid_grad)
            self.mask_eval_sample_fn = lambda x: sampler.rounding_sigmoid(x, bypass_sigmoid_grad)
        else:
            if mask_init_value is not None:
                logger.info(f"{type(self).__name__}: `mask_init_value` is always 1.0 for mask_type = `{mask_type}`")
            self.mask_init_value = 1.0
            self.mask_train_sample_fn = self.mask_eval_sample_fn = None
            self.mask_trainable = mask_type == prune.SNIP
        for m in self.mask_parameters:
            m.requires_grad = self.mask_trainable
        self.reset_masks()

    def reset_masks(self) -> None:
        for m in self.mask_parameters:
            init.constant_(m, self.mask_init_value)

    def mask_mode(self) -> int:
        """Kernel mask mode for the current (mask_type, training) state (masked_layer.py:92-102)."""
        if self.mask_type in prune.SUPER_MASKS:
            return K.MASK_BERNOULLI if self.training else K.MASK_ROUND
        return K.MASK_RAW

    def get_masked_weight(self, weight_name: str) -> Tensor:
        """Materialised ``sampled_mask * weight`` (kept for API compatibility; the layer forwards do not call it)."""
        weight = getattr(self, weight_name, None)
        assert weight is not None, f"Invalid weight attribute name: {weight_name}"
        mask = getattr(self, f"{weight_name}_pruning_mask", None)
        assert mask is not None, f"Invalid weight attribute name: {weight_name}_pruning_mask"
        if self.mask_type in prune.SUPER_MASKS:
            fn = self.mask_train_sample_fn if self.training else self.mask_eval_sample_fn
            return fn(mask) * weight
        return mask * weight

    @staticmethod
    def assert_in_kwargs(key, kwargs):
        assert key in kwargs, f"{key} not found in provided keyword arguments: {kwargs}"


# noinspection PyAbstractClass
class MaskedLinear(MaskMixin, nn.Linear):
    r"""y = x (W (.) mask)^T + b  with the mask applied in the GEMM operand prologue."""
    __constants__ = nn.Linear.__constants__ + ["mask_type", "mask_init_value", "bypass_sigmoid_grad"]

    def __init__(self, in_features: int, out_features: int, mask_type: str, mask_init_value: float,
                 bypass_sigmoid_grad: bool = False, **kwargs) -> None:
        super().__init__(in_features, out_features, **kwargs)
        self.setup_masks("weight", mask_type, mask_init_value, bypass_sigmoid_grad)

    def forward(self, input: Tensor) -> Tensor:
        return A.masked_linear(input, self.weight, self.weight_pruning_mask, self.bias, self.mask_mode(),
                               self.bypass_sigmoid_grad or self.mask_type not in prune.SUPER_MASKS, get_precision())


# noinspection PyAbstractClass
class MaskedEmbedding(MaskMixin, nn.Embedding):
    r"""Lookup of (W (.) mask)[ids]; only the gathered rows are ever masked."""
    __constants__ = nn.Embedding.__constants__ + ["mask_type", "mask_init_value", "bypass_sigmoid_grad"]

    def __init__(self, num_embeddings: int, embedding_dim: int, mask_type: str, mask_init_value: float,
                 bypass_sigmoid_grad: bool = False, **kwargs) -> None:
        super().__init__(num_embeddings, embedding_dim, **kwargs)
        assert self.padding_idx is None and self.max_norm is None and not self.scale_grad_by_freq and not self.sparse, \
            "MaskedEmbedding: padding_idx / max_norm / scale_grad_by_freq / sparse are not used by the captioning path"
        self.setup_masks("weight", mask_type, mask_init_value, bypass_sigmoid_grad)

    def forward(self, input: Tensor) -> Tensor:
        return A.masked_embedding(input, self.weight, self.weight_pruning_mask, self.mask_mode(),
                                  self.bypass_sigmoid_grad or self.mask_type not in prune.SUPER_MASKS)

    @classmethod
    def from_pretrained(cls, *args, **kwargs):
        raise NotImplementedError
