"""Host -> device input pipeline for the MVulD hot path: double-buffered staging on a copy stream.

The reference feeds its model from a ``DataLoader(pin_memory=True)`` and ``.cuda(non_blocking=True)`` on the compute
stream (/root/reference/mvuld/main_bigvul.py:308-324, mvuld/data/bigvul_dataset.py:177-205), so every step waits for
its own 190 MB of inputs.  Here batch i+1 is copied on a second CUDA stream while batch i computes, and each step's
result goes back through a pinned buffer without a host synchronisation, so the timed region still contains every copy
but the copies (and the launch latency of ~330 kernels per step) overlap the previous step's compute.
"""
from __future__ import annotations

from typing import Callable, Iterable, Iterator, List

import torch

from .graph import Graph
from .unixcoder import PackedLines


def _tensors(obj):
    if isinstance(obj, torch.Tensor):
        yield obj
    elif isinstance(obj, Graph):
        yield obj._src
        yield obj._dst
        yield from obj.ndata.values()
        yield from obj.edata.values()
    elif isinstance(obj, PackedLines):
        yield from obj.tensors()
    elif isinstance(obj, dict):
        for v in obj.values():
            yield from _tensors(v)
    elif isinstance(obj, (list, tuple)):
        for v in obj:
            yield from _tensors(v)


def to_device(batch, device, non_blocking: bool = True):
    """Move a batch (tensor / Graph / PackedLines / dict / list of those) to ``device``."""
    if isinstance(batch, (torch.Tensor, Graph, PackedLines)):
        return batch.to(device, non_blocking=non_blocking)
    if isinstance(batch, dict):
        return {k: to_device(v, device, non_blocking) for k, v in batch.items()}
    if isinstance(batch, (list, tuple)):
        return type(batch)(to_device(v, device, non_blocking) for v in batch)
    return batch


class DevicePrefetcher:
    """Iterate device-resident batches; the copy of the NEXT batch runs on a side stream while the caller computes.

    ``host_batches`` yields pinned host batches (tensor / ``Graph`` / dict / list).  Device buffers come from the
    COMPUTE stream's allocator pool (no cross-stream pools, no ``record_stream``, hence no allocator stalls): the copy
    stream first waits for everything already enqueued on the compute stream -- at that moment the compute of the
    previous batch -- and the batch's copy then overlaps the compute of the batch handed out just before it."""

    def __init__(self, host_batches: Iterable, device, stage: Callable = to_device):
        self.it: Iterator = iter(host_batches)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("DevicePrefetcher stages onto a CUDA device (no CPU fallback)")
        self.stage = stage
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self._next = None
        self._issue()

    def _issue(self):
        try:
            host = next(self.it)
        except StopIteration:
            self._next = None
            return
        cur = torch.cuda.current_stream(self.device)
        # allocate on the compute stream's pool, fill on the copy stream
        shells = _allocate_like(host, self.device)
        self.copy_stream.wait_stream(cur)
        with torch.cuda.stream(self.copy_stream):
            _fill(shells, host)
            ev = torch.cuda.Event()
            ev.record(self.copy_stream)
        self._next = (shells, ev)

    def __iter__(self):
        return self

    def __next__(self):
        if self._next is None:
            raise StopIteration
        dev, ev = self._next
        torch.cuda.current_stream(self.device).wait_event(ev)
        self._issue()                       # overlaps with whatever the caller launches next
        return dev


def _allocate_like(host, device):
    if isinstance(host, torch.Tensor):
        return torch.empty(host.shape, dtype=host.dtype, device=device)
    if isinstance(host, Graph):
        g = Graph.__new__(Graph)
        g._src, g._dst = _allocate_like(host._src, device), _allocate_like(host._dst, device)
        g._num_nodes, g._bnn, g._bne = host._num_nodes, host._bnn, host._bne
        g.ndata = {k: _allocate_like(v, device) for k, v in host.ndata.items()}
        g.edata = {k: _allocate_like(v, device) for k, v in host.edata.items()}
        g._csr = g._ocsr = g._offsets = None
        return g
    if isinstance(host, PackedLines):
        return PackedLines(host.n_lines, host.n_rows, host.n_tokens,
                           [{k: _allocate_like(v, device) for k, v in ps.items()} for ps in host.passes])
    if isinstance(host, dict):
        return {k: _allocate_like(v, device) for k, v in host.items()}
    if isinstance(host, (list, tuple)):
        return type(host)(_allocate_like(v, device) for v in host)
    return host


def _fill(dev, host):
    if isinstance(host, torch.Tensor):
        dev.copy_(host, non_blocking=True)
    elif isinstance(host, Graph):
        _fill(dev._src, host._src)
        _fill(dev._dst, host._dst)
        _fill(dev.ndata, host.ndata)
        _fill(dev.edata, host.edata)
    elif isinstance(host, PackedLines):
        for d, h in zip(dev.passes, host.passes):
            _fill(d, h)
    elif isinstance(host, dict):
        for k in host:
            _fill(dev[k], host[k])
    elif isinstance(host, (list, tuple)):
        for d, h in zip(dev, host):
            _fill(d, h)


class ResultSink:
    """Per-step device -> host read of a small result through ONE pinned ring buffer, without blocking the launching
    thread (``results()`` synchronises once at the end)."""

    def __init__(self, capacity: int):
        self.capacity, self.buf, self.n = int(capacity), None, 0

    def push(self, t: torch.Tensor):
        if self.buf is None:
            self.buf = torch.empty((self.capacity,) + tuple(t.shape), dtype=t.dtype, pin_memory=True)
        if self.n >= self.capacity:
            raise IndexError("ResultSink is full")
        self.buf[self.n].copy_(t, non_blocking=True)
        self.n += 1

    def results(self) -> List[torch.Tensor]:
        torch.cuda.current_stream().synchronize()
        return [self.buf[i] for i in range(self.n)]
