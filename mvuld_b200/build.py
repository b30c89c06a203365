"""``build_model(config)`` -- mirror of /root/reference/mvuld/models/build.py:14-102.

Only the ``swinv2`` branch (:33-50) is on the MVulD hot path: every YAML under the reference's
``mvuld/configs/mySwin`` sets ``MODEL.TYPE: swinv2``.  ``swin`` / ``swin_moe`` / ``swin_mlp`` are recognised model
types of the reference that this framework does not build (out of scope); any other string
raises ``NotImplementedError`` with the reference's message.
"""
from .swin_transformer_v2 import SwinTransformerV2


def build_model(config):
    model_type = config.MODEL.TYPE
    if model_type == 'swinv2':
        return SwinTransformerV2(img_size=config.DATA.IMG_SIZE,
                                 patch_size=config.MODEL.SWINV2.PATCH_SIZE,
                                 in_chans=config.MODEL.SWINV2.IN_CHANS,
                                 num_classes=config.MODEL.NUM_CLASSES,
                                 embed_dim=config.MODEL.SWINV2.EMBED_DIM,
                                 depths=config.MODEL.SWINV2.DEPTHS,
                                 num_heads=config.MODEL.SWINV2.NUM_HEADS,
                                 window_size=config.MODEL.SWINV2.WINDOW_SIZE,
                                 mlp_ratio=config.MODEL.SWINV2.MLP_RATIO,
                                 qkv_bias=config.MODEL.SWINV2.QKV_BIAS,
                                 drop_rate=config.MODEL.DROP_RATE,
                                 drop_path_rate=config.MODEL.DROP_PATH_RATE,
                                 ape=config.MODEL.SWINV2.APE,
                                 patch_norm=config.MODEL.SWINV2.PATCH_NORM,
                                 use_checkpoint=config.TRAIN.USE_CHECKPOINT,
                                 pretrained_window_sizes=config.MODEL.SWINV2.PRETRAINED_WINDOW_SIZES)
    if model_type in ('swin', 'swin_moe', 'swin_mlp'):
        raise NotImplementedError(f"model type {model_type!r} is not on the MVulD hot path (every reference YAML "
                                  "selects 'swinv2'); mvuld_b200 builds 'swinv2' only")
    raise NotImplementedError(f"Unkown model: {model_type}")
