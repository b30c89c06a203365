"""Checkpoint compatibility with the reference (/root/reference/mvuld/utils_multi.py:7-32,35-122,125-152).

* ``load_pretrained``: the weight surgery of ``utils_multi.load_pretrained`` for a SwinV2 checkpoint -- derived buffers
  (``relative_position_index``, ``relative_coords_table``, ``attn_mask``) are dropped because they are rebuilt from the
  geometry, a v1 ``relative_position_bias_table`` / ``absolute_pos_embed`` of another size is bicubically resized, a
  classifier head of another width is re-initialised to zero -- then ``load_state_dict(strict=False)``.
* ``save_checkpoint`` / ``load_checkpoint``: the reference's file layout ``{model, optimizer, lr_scheduler,
  max_accuracy, scaler, epoch, config}``.  ``optimizer`` holds the ``FusionTrainer`` state (flat AdamW moments, step
  count, hyper-parameters) instead of ``torch.optim.AdamW``'s per-parameter dict.

State-dict KEYS are the reference's own (tests load this package's weights into the reference classes with
``strict=True``), so reference checkpoints load into these models and vice versa.
"""
from __future__ import annotations

from typing import Optional

import torch

_DERIVED = ("relative_position_index", "relative_coords_table", "attn_mask")


def load_pretrained(model, checkpoint, logger=None):
    """utils_multi.py:35-122.  ``checkpoint``: a path, a ``{'model': state_dict}`` dict or a bare state dict."""
    if isinstance(checkpoint, str):
        checkpoint = torch.load(checkpoint, map_location="cpu", weights_only=False)
    state_dict = dict(checkpoint["model"] if "model" in checkpoint else checkpoint)
    for k in [k for k in state_dict if any(d in k for d in _DERIVED)]:
        del state_dict[k]                                          # always re-derived (utils_multi.py:40-53)
    current = model.state_dict()
    for k in [k for k in state_dict if "relative_position_bias_table" in k]:      # Swin v1 tables (utils_multi.py:55-71)
        if k not in current:
            continue
        pre, cur = state_dict[k], current[k]
        (L1, nH1), (L2, nH2) = pre.shape, cur.shape
        if nH1 != nH2:
            if logger:
                logger.warning(f"Error in loading {k}, passing......")
        elif L1 != L2:
            S1, S2 = int(L1 ** 0.5), int(L2 ** 0.5)
            r = torch.nn.functional.interpolate(pre.permute(1, 0).view(1, nH1, S1, S1), size=(S2, S2), mode="bicubic")
            state_dict[k] = r.view(nH2, L2).permute(1, 0)
    for k in [k for k in state_dict if "absolute_pos_embed" in k]:               # utils_multi.py:73-92
        if k not in current:
            continue
        pre, cur = state_dict[k], current[k]
        (_, L1, C1), (_, L2, C2) = pre.shape, cur.shape
        if C1 != C2:
            if logger:
                logger.warning(f"Error in loading {k}, passing......")
        elif L1 != L2:
            S1, S2 = int(L1 ** 0.5), int(L2 ** 0.5)
            r = torch.nn.functional.interpolate(pre.reshape(-1, S1, S1, C1).permute(0, 3, 1, 2), size=(S2, S2),
                                                mode="bicubic")
            state_dict[k] = r.permute(0, 2, 3, 1).flatten(1, 2)
    if "head.bias" in state_dict and hasattr(model, "head") and hasattr(model.head, "bias"):
        if state_dict["head.bias"].shape[0] != model.head.bias.shape[0]:         # utils_multi.py:94-113
            torch.nn.init.constant_(model.head.bias, 0.)
            torch.nn.init.constant_(model.head.weight, 0.)
            del state_dict["head.weight"], state_dict["head.bias"]
            if logger:
                logger.warning("Error in loading classifier head, re-init classifier head to 0")
    msg = model.load_state_dict(state_dict, strict=False)
    if hasattr(model, "invalidate"):
        model.invalidate()
    return msg


def save_checkpoint(path: str, epoch: int, model, trainer=None, max_accuracy: float = 0.0, lr_scheduler: Optional[dict] = None,
                    config=None):
    """utils_multi.py:125-137 file layout."""
    state = {"model": {k: v.detach().cpu().clone() for k, v in model.state_dict().items()},
             "optimizer": trainer.state_dict() if trainer is not None else None,
             "lr_scheduler": lr_scheduler or {}, "max_accuracy": max_accuracy, "scaler": {}, "epoch": epoch,
             "config": config.dump() if hasattr(config, "dump") else config}
    torch.save(state, path)
    return path


def load_checkpoint(path_or_dict, model, trainer=None, eval_mode: bool = False):
    """utils_multi.py:7-32 -> (max_accuracy, epoch).  Optimiser state is restored unless ``eval_mode``."""
    ck = torch.load(path_or_dict, map_location="cpu", weights_only=False) if isinstance(path_or_dict, str) else path_or_dict
    model.load_state_dict(ck["model"], strict=False)      # copies INTO the trainer's flat buffer when one owns the weights
    if hasattr(model, "invalidate"):
        model.invalidate()
    max_accuracy = 0.0
    if trainer is not None:
        if not eval_mode and ck.get("optimizer") is not None:
            trainer.load_state_dict(ck["optimizer"])
            max_accuracy = ck.get("max_accuracy", 0.0)
        trainer.refresh()
    return max_accuracy, ck.get("epoch", 0)
