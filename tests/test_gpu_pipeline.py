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


def test_host_writer_copies_on_a_side_stream():
    from multimodalbrainsurvival_b200 import pipeline
    w = pipeline.HostWriter("cuda:0")
    outs = [torch.empty(256, 512).pin_memory() for _ in range(4)]
    for i, o in enumerate(outs):
        x = torch.full((256, 512), float(i), device="cuda:0") * 3 + 1     # produced on the current stream
        w.write(x, o)
        del x                                                              # the writer keeps the storage alive
    w.wait()
    assert [float(o[0, 0]) for o in outs] == [1.0, 4.0, 7.0, 10.0]
    assert all(bool((o == o[0, 0]).all()) for o in outs)


def test_prefetch_to_device_takes_tuples():
    """(patch_bag, rna) pairs of the joint-fusion loader travel in one slot, in order, on the copy stream."""
    from multimodalbrainsurvival_b200 import pipeline
    g = torch.Generator().manual_seed(0)
    host = [(torch.randn(4, 3, 8, 8, generator=g), torch.randn(4, 17, generator=g)) for _ in range(5)]
    for i, (x, r) in enumerate(pipeline.prefetch_to_device(iter(host), "cuda:0", depth=2)):
        assert x.is_cuda and r.is_cuda
        assert torch.equal(x.cpu(), host[i][0]) and torch.equal(r.cpu(), host[i][1])
