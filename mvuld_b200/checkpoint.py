"""Checkpoint compatibility with the reference (/root/reference/mvuld/utils_multi.py:7-32,35-122,125-152).

* ``load_pretrained``: the weight surgery of ``utils_multi.load_pretrained`` for a SwinV2 checkpoint -- derived buffers
  (``relative_position_index``, ``relative_coords_table``, ``attn_mask``) are dropped because they are rebuilt from the
  geometry, a v1 ``relative_position_bias_table`` / ``absolute_pos_embed`` of another size is bicubically resized, a
  classifier head of another width is re-initialised to zero -- then ``load_state_dict(strict=False)``.
* ``save_checkpoint`` / ``load_checkpoint``: the reference's file layout ``{model, optimizer, lr_scheduler,
  max_accuracy, scaler, epoch, config}``.  ``optimizer`` is in ``torch.optim.AdamW.state_dict()``'s own layout
  (``FusionTrainer.state_dict`` converts its flat moments to per-parameter ``{step, exp_avg, exp_avg_sq}`` entries in
  ``build_optimizer``'s group order, optimizer.py:35-50, and back), ``scaler`` in ``torch.amp.GradScaler.state_dict()``'s
  (scale 1: the bf16 step needs no loss scaling), so a checkpoint written here resumes in ``utils_multi.load_checkpoint``
  and a reference checkpoint resumes here.  Files are read with ``weights_only=True`` first (tensors and plain
  containers); a reference checkpoint that pickles its yacs ``config`` node needs ``trusted=True``.

State-dict KEYS are the reference's own (tests load this package's weights into the reference classes with
``strict=True``), so reference checkpoints load into these models and vice versa.
"""
from __future__ import annotations

from typing import Optional

import torch

_DERIVED = ("relative_position_index", "relative_coords_table", "attn_mask")


def _read(path: str, trusted: bool = False):
    """``torch.load`` restricted to tensors / plain containers unless the caller vouches for the file (a pickled yacs
    node, as the reference's ``save_checkpoint`` writes under ``config``, executes code on load)."""
    try:
        return torch.load(path, map_location="cpu", weights_only=True)
    except Exception as e:                                  # noqa: BLE001 -- the unpickler's error types vary by version
        if not trusted:
            raise RuntimeError(f"{path} holds objects beyond tensors and plain containers ({type(e).__name__}: {e}); "
                               "pass trusted=True to unpickle it if you trust its origin") from e
        return torch.load(path, map_location="cpu", weights_only=False)


#: ``torch.amp.GradScaler().state_dict()`` keys; scale 1 = no loss scaling (bf16 step), loadable by the reference's scaler
SCALER_STATE = {"scale": 1.0, "growth_factor": 2.0, "backoff_factor": 0.5, "growth_interval": 2000, "_growth_tracker": 0}


def load_pretrained(model, checkpoint, logger=None, trusted: bool = False):
    """utils_multi.py:35-122.  ``checkpoint``: a path, a ``{'model': state_dict}`` dict or a bare state dict."""
    if isinstance(checkpoint, str):
        checkpoint = _read(checkpoint, trusted)
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
             "lr_scheduler": lr_scheduler or {}, "max_accuracy": max_accuracy, "scaler": dict(SCALER_STATE),
             "epoch": epoch,
             # the yaml text of the node: a pickled CfgNode would make the file unreadable under weights_only=True
             "config": config.dump() if hasattr(config, "dump") else config}
    torch.save(state, path)
    return path


def load_checkpoint(path_or_dict, model, trainer=None, eval_mode: bool = False, trusted: bool = False):
    """utils_multi.py:7-32 -> (max_accuracy, epoch).  Optimiser state (this package's or a reference
    ``torch.optim.AdamW`` one) is restored unless ``eval_mode``."""
    ck = _read(path_or_dict, trusted) if isinstance(path_or_dict, str) else path_or_dict
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
