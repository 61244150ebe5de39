"""Host <-> device staging helpers for the extract / savescore loops.

The reference moves every batch with a blocking ``.to(device)`` and reads features back with
``.cpu()`` (/root/reference/1_HistoPathology/4_HistoPath_extractfeatures.py:60-71).
``prefetch_to_device`` overlaps the H2D copy of batch i+1 with the kernels of batch i on a
side stream (pinned staging buffers), which is what keeps the B200 busy at PCIe rates."""
from __future__ import annotations

import torch


def prefetch_to_device(batches, device, depth: int = 2):
    """Iterate over host tensors, yielding device tensors whose copy was issued one step ahead."""
    dev = torch.device(device)
    copy_stream = torch.cuda.Stream(device=dev)
    queue = []
    it = iter(batches)

    def issue():
        try:
            host = next(it)
        except StopIteration:
            return False
        if not host.is_pinned():
            host = host.pin_memory()
        with torch.cuda.stream(copy_stream):
            d = host.to(dev, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        queue.append((d, ev, host))
        return True

    for _ in range(depth):
        if not issue():
            break
    while queue:
        d, ev, _host = queue.pop(0)
        torch.cuda.current_stream(dev).wait_event(ev)
        d.record_stream(torch.cuda.current_stream(dev))
        issue()
        yield d
