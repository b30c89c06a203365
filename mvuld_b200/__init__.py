"""mvuld_b200 -- B200-native (sm_100a) implementation of the MVulD multimodal forward hot path.

Host side mirrors the reference's model-layer interface (``build_model`` / ``get_config`` / model classes with the
reference's state-dict keys); compute is hand-written CUDA behind the C ABI in ``include/mvuld_b200.h``.
"""
from .config import get_config, default_config, CfgNode            # noqa: F401
from .build import build_model                                      # noqa: F401
from .swin_transformer_v2 import SwinTransformerV2                  # noqa: F401
from .unixcoder import MyUniXcoder, RobertaEncoder, build_MyUniXcoder, roberta_base_config   # noqa: F401
from .graph_model import (Multi_DefectModel_new_GCN, Multi_DefectModel, Rs_GCN, GATConv, GatedGraphConv,  # noqa: F401
                          GGNNSum)
from .fusion_variants import (Multi_DefectModel_noGraph, Multi_DefectModel_000, Multi_DefectModel_001,   # noqa: F401
                              Multi_DefectModel_100, Multi_DefectModel_NOGAT2, Multi_DefectModel_noFunc,
                              Multi_DefectModel_noGlobalImage, ABLATIONS, Multi_DefectModel_110,
                              Multi_DefectModel_GATPOS, Multi_DefectModel_011, Multi_DefectModel_NOGAT,
                              Multi_DefectModel_NOGAT3, Multi_DefectModel_NOGAT4, GRID_VARIANTS)
from . import my_models                                             # noqa: F401
from .mvuld import MVulD                                            # noqa: F401
from . import graph, checkpoint                                     # noqa: F401

__all__ = ["get_config", "default_config", "CfgNode", "build_model", "SwinTransformerV2", "MyUniXcoder",
           "RobertaEncoder", "build_MyUniXcoder", "roberta_base_config", "Multi_DefectModel_new_GCN", "Rs_GCN",
           "GATConv", "GatedGraphConv", "GGNNSum", "MVulD", "graph", "Multi_DefectModel",
           "Multi_DefectModel_noGraph", "Multi_DefectModel_000", "Multi_DefectModel_001", "Multi_DefectModel_100",
           "Multi_DefectModel_NOGAT2", "Multi_DefectModel_noFunc", "Multi_DefectModel_noGlobalImage", "ABLATIONS",
           "Multi_DefectModel_110", "Multi_DefectModel_GATPOS", "Multi_DefectModel_011", "Multi_DefectModel_NOGAT",
           "Multi_DefectModel_NOGAT3", "Multi_DefectModel_NOGAT4", "GRID_VARIANTS", "my_models"]
