"""GPU test of the staging helper: order and contents of the prefetched batches."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_prefetch_to_device_preserves_order_and_contents():
    from multimodalbrainsurvival_b200 import pipeline
    host = [torch.full((64, 1024), float(i)).pin_memory() for i in range(7)]
    sums = []
    for i, d in enumerate(pipeline.prefetch_to_device(iter(host), "cuda:0", depth=2)):
        assert d.is_cuda and d.shape == (64, 1024)
        y = d * 2 + 1                      # consumer work on the current stream
        sums.append(float(y.sum()))
    assert sums == [float((2 * i + 1) * 64 * 1024) for i in range(7)]
    u8 = [torch.randint(0, 256, (5, 3, 8, 8), dtype=torch.uint8) for _ in range(3)]     # pageable, uint8
    got = [d.clone().cpu() for d in pipeline.prefetch_to_device(iter(u8), "cuda:0", depth=1)]
    assert all(torch.equal(a, b) for a, b in zip(got, u8))
