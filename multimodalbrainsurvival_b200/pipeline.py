"""Host <-> device staging helpers for the extract / savescore loops.

The reference moves every batch with a blocking ``.to(device)`` and reads features back with
``.cpu()`` (/root/reference/1_HistoPathology/4_HistoPath_extractfeatures.py:60-71).
``prefetch_to_device`` overlaps the H2D copy of batch i+1 with the kernels of batch i on a
side stream, through a fixed ring of device staging buffers (no allocation, and therefore no
implicit synchronisation, inside the loop)."""
from __future__ import annotations

import torch


def prefetch_to_device(batches, device, depth: int = 2):
    """Iterate over host tensors, yielding device tensors whose copy was issued ``depth`` steps
    ahead.  A yielded tensor is only valid until the next one is requested (its buffer is reused)."""
    dev = torch.device(device)
    copy_stream = torch.cuda.Stream(device=dev)
    slots = depth + 1
    ring = [None] * slots          # device staging buffers
    ready = [None] * slots         # copy finished (recorded on the copy stream)
    released = [None] * slots      # consumer finished with the buffer (recorded on its stream)
    it = iter(batches)
    head = 0                       # next slot to fill
    pending = []

    def issue():
        nonlocal head
        try:
            host = next(it)
        except StopIteration:
            return False
        if not host.is_pinned():
            host = host.pin_memory()
        k = head
        head = (head + 1) % slots
        if ring[k] is None or ring[k].shape != host.shape or ring[k].dtype != host.dtype:
            ring[k] = torch.empty(host.shape, dtype=host.dtype, device=dev)
        with torch.cuda.stream(copy_stream):
            if released[k] is not None:
                copy_stream.wait_event(released[k])
            ring[k].copy_(host, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        ready[k] = ev
        pending.append((k, host))   # keep the pinned source alive until its copy is consumed
        return True

    for _ in range(depth):
        if not issue():
            break
    while pending:
        k, _host = pending.pop(0)
        cur = torch.cuda.current_stream(dev)
        cur.wait_event(ready[k])
        issue()
        yield ring[k]
        rel = torch.cuda.Event()
        rel.record(torch.cuda.current_stream(dev))
        released[k] = rel


class HostWriter:
    """Device -> pinned-host copies on a side stream: the D2H read-back of step i overlaps the kernels of step i+1
    instead of sitting in the compute stream (a 4 MB feature block at PCIe speed stalls it for ~0.15 ms).
    ``write`` orders the copy after everything enqueued so far on the current stream; ``wait`` makes the host
    wait for all pending copies (call it before reading the host buffers)."""

    def __init__(self, device):
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(device=self.device)

    def write(self, src: torch.Tensor, dst_host: torch.Tensor) -> None:
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream(self.device))
        self.stream.wait_event(ready)
        with torch.cuda.stream(self.stream):
            dst_host.copy_(src, non_blocking=True)
        src.record_stream(self.stream)     # the caching allocator must not recycle src before the copy ran

    def wait(self) -> None:
        self.stream.synchronize()
