"""Composed end-to-end MVulD forward.

The reference never runs the three modalities in one pass: SwinV2 and UniXcoder vectors are cached to disk offline
(/root/reference/mvuld/data/data_list.py:179-211,292-313) and only the fusion model runs in the training loop
(/root/reference/mvuld/main_bigvul.py:308-329).  This module wires the same three pieces back to back on one stream:

    img_embedding       = swin.forward_features(image)          # swin_transformer_v2.py:623-635
    func_text_embedding = unix.get_repr(token_ids)[0]           # unixcoder.py:91-95
    logits              = fusion(g, img_embedding, func_text_embedding)   # GraphModel.py:150-211
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .build import build_model
from .graph_model import Multi_DefectModel_new_GCN
from .unixcoder import build_MyUniXcoder


class MVulD(nn.Module):
    def __init__(self, config, roberta_config=None):
        super().__init__()
        self.config = config
        self.swin = build_model(config)
        self.unix = build_MyUniXcoder(roberta_config)
        self.fusion = Multi_DefectModel_new_GCN(config)
        # the graph half of the fusion model needs nothing from the encoders: with overlap_graph_branch it runs on a
        # second CUDA stream while the image / text branches compute (its kernels are small: GAT and node-MLP GEMMs over
        # ~12 k rows, 64-graph Rs_GCN blocks).  Measured neutral on B200 (2 133-2 177 vs 2 148-2 156 functions/s: the
        # step runs into the power cap, so filling idle SMs lowers the clock instead), hence off by default.
        self.overlap_graph_branch = False
        self._side = None

    @torch.no_grad()
    def forward(self, image: torch.Tensor, token_ids, g) -> torch.Tensor:
        """``token_ids``: ``[B, 512]`` ids as the reference's tokenizer pads them, or the same batch packed at
        data-loading time (``self.unix.encoder.pack_host(ids)``), which skips the pad tokens' share of the encoder."""
        if not self.overlap_graph_branch:
            img_embedding = self.swin.forward_features(image)
            func_text_embedding, _ = self.unix.get_repr(token_ids)
            return self.fusion(g, img_embedding, func_text_embedding)
        from .graph import Graph, from_dgl
        if not isinstance(g, Graph):
            g = from_dgl(g)
        main = torch.cuda.current_stream()
        if self._side is None:
            self._side = torch.cuda.Stream(device=main.device)
        self._side.wait_stream(main)                       # the inputs were produced on (or before) the main stream
        with torch.cuda.stream(self._side):
            z32 = self.fusion.graph_features(g)
        img_embedding = self.swin.forward_features(image)
        func_text_embedding, _ = self.unix.get_repr(token_ids)
        main.wait_stream(self._side)
        z32.record_stream(main)                            # allocated on the side stream, consumed on the main one
        return self.fusion.head(z32, img_embedding, func_text_embedding)
