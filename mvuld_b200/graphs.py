"""CUDA graphs over the fixed-shape launch sequences of the forward path.

At the reference's documented batch size (README.md:63-66: 4 functions per step) a forward is ~300 kernel launches of a
few microseconds each: the GPU finishes a launch faster than the Python / ctypes boundary can issue the next one, and
the step is launch bound.  The SwinV2 branch (fixed [B, 3, 448, 448] input, preallocated workspace) and the UniXcoder
branch on the tokenizer's padded [B, 512] rows are fixed launch sequences for a given batch size, so each is captured
once per batch size into a CUDA graph and replayed with one driver call.  The captured kernels are the same launches
(same tensor maps, same workspaces) -- replay results are bit-identical to the eager path.  The graph / fusion branch
has data-dependent sizes (nodes, edges, CSR build) and stays eager.

At small batches the kernels themselves are latency bound (a 28 x 28 window attention launch at one image is 16 CTAs on
148 SMs), so the three independent branches -- image graph, text graph, eager graph branch -- also run CONCURRENTLY on
three streams and meet at the head (``concurrent_below``: per-GPU batches up to that size; at 64 functions per step the
branches already fill the GPU and overlapping them was measured neutral, see MVulD.overlap_graph_branch).

    model = MVulD(cfg).eval().cuda()
    fast = GraphedMVulD(model)           # drop-in: fast(image, token_ids, g) -> logits
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch

from .unixcoder import PackedLines


class _Captured:
    def __init__(self, graph, static_in, static_out, plan, keep):
        self.graph, self.static_in, self.static_out = graph, static_in, static_out
        # the graph holds raw pointers into the packed weights (`plan`) and the workspace (`keep`): both stay referenced
        # here for the graph's lifetime, and a re-packed plan (weights changed, model moved) invalidates the capture
        self.plan, self.keep = plan, keep


class GraphedMVulD:
    """``MVulD.forward`` with the image branch and the padded text branch replayed from CUDA graphs (one per batch
    size).  Packed token ids (``PackedLines``: row count varies from batch to batch) run eagerly."""

    def __init__(self, model, warmup: int = 2, concurrent_below: int = 16):
        if model.training:
            raise RuntimeError("GraphedMVulD wraps the eval-mode forward: call model.eval()")
        self.model, self.warmup, self.concurrent_below = model, int(warmup), int(concurrent_below)
        self._streams = None
        self._swin: Dict[int, _Captured] = {}
        self._text: Dict[Tuple[int, int], _Captured] = {}

    # ------------------------------------------------------------------------------------------------
    def _capture(self, fn, example: torch.Tensor, plan_of, keep_of) -> _Captured:
        static_in = example.clone()
        side = torch.cuda.Stream(device=example.device)
        side.wait_stream(torch.cuda.current_stream(example.device))
        with torch.cuda.stream(side):                   # warm-up off the capture: lazy plans, workspaces, attributes
            for _ in range(self.warmup):
                fn(static_in)
        torch.cuda.current_stream(example.device).wait_stream(side)
        torch.cuda.synchronize(example.device)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            static_out = fn(static_in)
        return _Captured(graph, static_in, static_out, plan_of(), keep_of())

    def image_features(self, image: torch.Tensor) -> torch.Tensor:
        swin = self.model.swin
        swin._check_input(image)
        if swin._plan is None:
            swin.prepare()
        B = image.shape[0]
        cap = self._swin.get(B)
        if cap is None or cap.plan is not swin._plan:
            x = image.to(torch.float32).contiguous()
            ws = swin._workspace(B)
            cap = self._capture(lambda t: swin._forward_features_ws(t, ws), x, lambda: swin._plan, lambda: ws)
            self._swin[B] = cap
        cap.static_in.copy_(image, non_blocking=True)
        cap.graph.replay()
        return cap.static_out.clone()

    def text_features(self, token_ids) -> torch.Tensor:
        unix = self.model.unix
        if isinstance(token_ids, PackedLines) or not token_ids.is_cuda:
            return unix.get_repr(token_ids)[0]
        ids = token_ids.view(-1, unix.max_source_length).to(torch.int64).contiguous()
        key = tuple(ids.shape)
        cap = self._text.get(key)
        enc = unix.encoder
        if cap is None or cap.plan is not enc._plan:
            cap = self._capture(lambda t: unix.get_repr(t)[0], ids, lambda: enc._plan, lambda: enc._plan["ws"].get(key))
            self._text[key] = cap
        cap.static_in.copy_(ids, non_blocking=True)
        cap.graph.replay()
        return cap.static_out.clone()

    @torch.no_grad()
    def __call__(self, image: torch.Tensor, token_ids, g) -> torch.Tensor:
        if image.shape[0] > self.concurrent_below:
            img_embedding = self.image_features(image)
            func_text_embedding = self.text_features(token_ids)
            return self.model.fusion(g, img_embedding, func_text_embedding)
        from .graph import Graph, from_dgl
        if not isinstance(g, Graph):
            g = from_dgl(g)
        main = torch.cuda.current_stream(image.device)
        if self._streams is None:
            self._streams = (torch.cuda.Stream(device=image.device), torch.cuda.Stream(device=image.device))
        s_img, s_txt = self._streams
        s_img.wait_stream(main)                            # the inputs were produced on (or before) the caller's stream
        s_txt.wait_stream(main)
        with torch.cuda.stream(s_img):
            img_embedding = self.image_features(image)
        with torch.cuda.stream(s_txt):
            func_text_embedding = self.text_features(token_ids)
        z32 = self.model.fusion.graph_features(g)          # eager, on the caller's stream, under the two replays
        main.wait_stream(s_img)
        main.wait_stream(s_txt)
        img_embedding.record_stream(main)                  # allocated on the side streams, consumed on this one
        func_text_embedding.record_stream(main)
        return self.model.fusion.head(z32, img_embedding, func_text_embedding)

    forward = __call__
