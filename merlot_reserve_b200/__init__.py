"""merlot_reserve_b200 -- B200-native (sm_100a) encoder + contrastive hot path of MERLOT Reserve.

`modeling` mirrors `mreserve/modeling.py`, `pretrain_model` the forward + loss of
`pretrain/pretrain_model.py`; `_lib` binds the C ABI of include/mreserve_b200.h;
`build` compiles the CUDA sources in-tree.  Importing the package does not need a GPU; calling any
model method does (there is no CPU fallback).
"""
from .modeling import MerlotReserve, PretrainedMerlotReserve, one_hot_pool, unit_normalize  # noqa: F401
from .pretrain_model import MerlotReservePretrainer, gather_contrastive_targets, loss_fn_given_preds  # noqa: F401

__all__ = ["MerlotReserve", "PretrainedMerlotReserve", "MerlotReservePretrainer", "loss_fn_given_preds",
           "gather_contrastive_targets", "one_hot_pool", "unit_normalize"]
