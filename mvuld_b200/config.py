"""yacs-compatible configuration for the MVulD hot path.

Mirrors /root/reference/mvuld/config.py:5-400: same tree (``DATA``, ``MODEL.SWINV2``, ``TRAIN`` ...),
same YAML merge with recursive ``BASE``, same ``get_config(args)`` entry point, and the
``defrost()/freeze()/clone()/dump()/merge_from_file()/merge_from_list()`` calls the reference's callers
use (mvuld/main_bigvul.py:195-197,566).  yacs itself is not a dependency: ``CfgNode`` below is a small
attribute-dict with the same behaviour for those calls.  Only keys the model boundary reads carry
meaning here; the rest are kept so reference YAMLs and ``--opts`` lists merge without error.
"""
from __future__ import annotations

import copy
import os
from ast import literal_eval

import yaml


class CfgNode(dict):
    _FROZEN = "__frozen__"

    def __init__(self, init=None):
        super().__init__()
        object.__setattr__(self, CfgNode._FROZEN, False)
        for k, v in (init or {}).items():
            self[k] = CfgNode(v) if isinstance(v, dict) and not isinstance(v, CfgNode) else v

    def __getattr__(self, name):
        if name in self:
            return self[name]
        raise AttributeError(name)

    def __setattr__(self, name, value):
        if self.is_frozen():
            raise AttributeError(f"Attempted to set {name} to {value}, but CfgNode is immutable")
        self[name] = value

    def is_frozen(self):
        return object.__getattribute__(self, CfgNode._FROZEN)

    def _set_frozen(self, flag):
        object.__setattr__(self, CfgNode._FROZEN, flag)
        for v in self.values():
            if isinstance(v, CfgNode):
                v._set_frozen(flag)

    def freeze(self):
        self._set_frozen(True)

    def defrost(self):
        self._set_frozen(False)

    def clone(self):
        return copy.deepcopy(self)

    def __deepcopy__(self, memo):
        out = CfgNode()
        for k, v in self.items():
            dict.__setitem__(out, k, copy.deepcopy(v, memo))
        object.__setattr__(out, CfgNode._FROZEN, self.is_frozen())
        return out

    def _to_dict(self):
        return {k: (v._to_dict() if isinstance(v, CfgNode) else (list(v) if isinstance(v, tuple) else v))
                for k, v in self.items()}

    def dump(self, **kw):
        return yaml.safe_dump(self._to_dict(), **kw)

    def _merge(self, other: dict, path=""):
        for k, v in other.items():
            if k not in self:
                raise KeyError(f"Non-existent config key: {path + k}")
            if isinstance(self[k], CfgNode):
                if not isinstance(v, dict):
                    raise ValueError(f"Type mismatch for config key {path + k}")
                self[k]._merge(v, path + k + ".")
            else:
                old = self[k]
                if isinstance(old, tuple) and isinstance(v, list):
                    v = tuple(v)
                if isinstance(old, float) and isinstance(v, int) and not isinstance(v, bool):
                    v = float(v)
                dict.__setitem__(self, k, v)

    def merge_from_other_cfg(self, other):
        if self.is_frozen():
            raise AttributeError("CfgNode is immutable")
        self._merge(other)

    def merge_from_file(self, cfg_filename):
        if self.is_frozen():
            raise AttributeError("CfgNode is immutable")
        with open(cfg_filename, "r") as f:
            loaded = yaml.safe_load(f) or {}
        loaded.pop("BASE", None) if "BASE" not in self else None
        self._merge(loaded)

    def merge_from_list(self, cfg_list):
        if self.is_frozen():
            raise AttributeError("CfgNode is immutable")
        if len(cfg_list) % 2 != 0:
            raise ValueError("Override list has odd length")
        for full_key, v in zip(cfg_list[0::2], cfg_list[1::2]):
            node = self
            parts = full_key.split(".")
            for p in parts[:-1]:
                if p not in node:
                    raise KeyError(f"Non-existent key: {full_key}")
                node = node[p]
            if parts[-1] not in node:
                raise KeyError(f"Non-existent key: {full_key}")
            if isinstance(v, str):
                try:
                    v = literal_eval(v)
                except (ValueError, SyntaxError):
                    pass
            old = node[parts[-1]]
            if isinstance(old, tuple) and isinstance(v, list):
                v = tuple(v)
            if isinstance(old, float) and isinstance(v, int) and not isinstance(v, bool):
                v = float(v)
            dict.__setitem__(node, parts[-1], v)


CN = CfgNode


def _defaults() -> CfgNode:
    """Default tree of mvuld/config.py:5-322 (later duplicate assignments win, as in the reference)."""
    C = CN()
    C.BASE = [""]
    C.DATA = CN(dict(BATCH_SIZE=128, DATA_PATH="datasets", DATASET="imagenet", IMG_SIZE=384,
                     INTERPOLATION="bicubic", ZIP_MODE=False, CACHE_MODE="part", PIN_MEMORY=False,
                     NUM_WORKERS=8))
    C.MODEL = CN(dict(TYPE="swin2", NAME="swin_base_patch4_window7_224", PRETRAINED="", RESUME="",
                      NUM_CLASSES=2, DROP_RATE=0.0, DROP_PATH_RATE=0.1, LABEL_SMOOTHING=0.1))
    swin_common = dict(PATCH_SIZE=4, IN_CHANS=3, EMBED_DIM=96, DEPTHS=[2, 2, 6, 2], NUM_HEADS=[3, 6, 12, 24],
                       WINDOW_SIZE=7, MLP_RATIO=4.0, APE=False, PATCH_NORM=True)
    C.MODEL.SWIN = CN(dict(swin_common, QKV_BIAS=True, QK_SCALE=None))
    C.MODEL.SWINV2 = CN(dict(swin_common, QKV_BIAS=True, PRETRAINED_WINDOW_SIZES=[0, 0, 0, 0]))
    C.MODEL.SWIN_MOE = CN(dict(swin_common, QKV_BIAS=True, QK_SCALE=None, MLP_FC2_BIAS=True, INIT_STD=0.02,
                               PRETRAINED_WINDOW_SIZES=[0, 0, 0, 0], MOE_BLOCKS=[[-1], [-1], [-1], [-1]],
                               NUM_LOCAL_EXPERTS=1, TOP_VALUE=1, CAPACITY_FACTOR=1.25, COSINE_ROUTER=False,
                               NORMALIZE_GATE=False, USE_BPR=True, IS_GSHARD_LOSS=False, GATE_NOISE=1.0,
                               COSINE_ROUTER_DIM=256, COSINE_ROUTER_INIT_T=0.5, MOE_DROP=0.0,
                               AUX_LOSS_WEIGHT=0.01))
    C.MODEL.SWIN_MLP = CN(dict(swin_common))
    C.MODEL.MULTI = CN(dict(RESUME=""))
    C.TRAIN = CN(dict(START_EPOCH=0, EPOCHS=500, WARMUP_EPOCHS=20, WEIGHT_DECAY=0.005, BASE_LR=5e-5,
                      WARMUP_LR=5e-7, MIN_LR=5e-6, CLIP_GRAD=5.0, AUTO_RESUME=False, BEST_RESUME=True,
                      ACCUMULATION_STEPS=1, USE_CHECKPOINT=False,
                      DATA_PATH="datasets/total/train_balanced.txt"))
    C.TRAIN.LR_SCHEDULER = CN(dict(NAME="cosine", DECAY_EPOCHS=30, DECAY_RATE=0.1))
    C.TRAIN.OPTIMIZER = CN(dict(NAME="adamw", EPS=1e-8, BETAS=(0.9, 0.999), MOMENTUM=0.9))
    C.TRAIN.MOE = CN(dict(SAVE_MASTER=False))
    C.AUG = CN(dict(COLOR_JITTER=0.4, AUTO_AUGMENT="rand-m9-mstd0.5-inc1", REPROB=0.25, REMODE="pixel",
                    RECOUNT=1, MIXUP=0.8, CUTMIX=1.0, CUTMIX_MINMAX=None, MIXUP_PROB=1.0,
                    MIXUP_SWITCH_PROB=0.5, MIXUP_MODE="batch"))
    C.TEST = CN(dict(CROP=False, SEQUENTIAL=False, SHUFFLE=False, DATA_PATH="datasets/total/test.txt"))
    C.VAL = CN(dict(DATA_PATH="datasets/total/valid.txt"))
    C.AMP_ENABLE = True
    C.AMP_OPT_LEVEL = ""
    C.OUTPUT = "output"
    C.MULTI_OUTPUT = "myoutput/Multi_DefectModel_new_GCN/3"
    C.TAG = "default"
    C.SAVE_FREQ = 1
    C.PRINT_FREQ = 50
    C.SEED = 0
    C.EVAL_MODE = False
    C.THROUGHPUT_MODE = False
    C.LOCAL_RANK = 0
    return C


_C = _defaults()


def _update_config_from_file(config: CfgNode, cfg_file: str):
    """mvuld/config.py:324-336 (recursive BASE merge)."""
    config.defrost()
    with open(cfg_file, "r") as f:
        yaml_cfg = yaml.safe_load(f) or {}
    for base in yaml_cfg.setdefault("BASE", [""]):
        if base:
            _update_config_from_file(config, os.path.join(os.path.dirname(cfg_file), base))
    config._merge(yaml_cfg)
    config.freeze()


def update_config(config: CfgNode, args):
    """mvuld/config.py:339-390."""
    if getattr(args, "cfg", None):
        _update_config_from_file(config, args.cfg)
    config.defrost()
    ga = lambda n: getattr(args, n, None)
    if ga("opts"):
        config.merge_from_list(args.opts)
    if ga("batch_size"):
        config.DATA.BATCH_SIZE = args.batch_size
    if ga("data_path"):
        config.DATA.DATA_PATH = args.data_path
    if ga("test_data_path"):
        config.TEST.DATA_PATH = args.test_data_path
    if ga("zip"):
        config.DATA.ZIP_MODE = True
    if ga("cache_mode"):
        config.DATA.CACHE_MODE = args.cache_mode
    if ga("pretrained"):
        config.MODEL.PRETRAINED = args.pretrained
    if ga("resume"):
        config.MODEL.RESUME = args.resume
    if ga("myresume"):
        config.MODEL.MULTI.RESUME = args.myresume
    if ga("accumulation_steps"):
        config.TRAIN.ACCUMULATION_STEPS = args.accumulation_steps
    if ga("use_checkpoint"):
        config.TRAIN.USE_CHECKPOINT = True
    if ga("amp_opt_level") == "O0":
        config.AMP_ENABLE = False
    if ga("disable_amp"):
        config.AMP_ENABLE = False
    if ga("output"):
        config.OUTPUT = args.output
    if ga("tag"):
        config.TAG = args.tag
    if ga("eval"):
        config.EVAL_MODE = True
    if ga("throughput"):
        config.THROUGHPUT_MODE = True
    config.LOCAL_RANK = ga("local_rank") or 0
    config.OUTPUT = os.path.join(config.OUTPUT, config.MODEL.NAME, config.TAG)
    config.MULTI_OUTPUT = os.path.join(config.MULTI_OUTPUT, config.MODEL.NAME, config.TAG)
    config.freeze()


def get_config(args) -> CfgNode:
    """mvuld/config.py:393-400."""
    config = _C.clone()
    update_config(config, args)
    return config


_HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_YAML = os.path.join(_HERE, "configs", "swinv2_base_patch4_window24to28_384to448_1ktoMYDATA_ft.yaml")


def default_config(**overrides) -> CfgNode:
    """The BASELINE configuration (448 px, window 28) without an argparse namespace."""
    class _A:
        cfg = DEFAULT_YAML
        opts = [x for kv in overrides.items() for x in kv] or None
    return get_config(_A())
