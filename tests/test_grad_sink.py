"""Host logic of the model's gradient sink (model._GradSink): how the flat gradient arena reaches ``param.grad``
with torch's accumulate-unless-zeroed semantics.  Pure tensor bookkeeping: runs on CPU."""
import torch

from street_gaussians_ns_b200 import raster
from street_gaussians_ns_b200.model import _GradSink


def _static(shapes):
    sizes = [[(int(torch.Size(s).numel()) + 3) // 4 * 4 for s in seg] for seg in shapes]
    return dict(sizes=sizes, shapes=shapes)


def test_arena_layout_and_views_are_16_byte_slices():
    st = _static([[(5, 3), (5, 1), (5, 2, 3)], [(2, 3), (2, 1), (2, 2, 3)]])
    sizes, shapes, numels = raster.arena_layout(st)
    assert sizes == [16, 8, 32, 8, 4, 12] and numels == [15, 5, 30, 6, 2, 12]
    arena = torch.arange(sum(sizes), dtype=torch.float32)
    views = raster.arena_views(arena, st)
    assert [tuple(v.shape) for v in views] == [tuple(s) for seg in st["shapes"] for s in seg]
    off = 0
    for v, sz in zip(views, sizes):
        assert v.data_ptr() == arena.data_ptr() + 4 * off and off % 4 == 0   # every slice starts 16-byte aligned
        assert torch.equal(v.reshape(-1), arena[off: off + v.numel()])
        off += sz


def test_sink_overwrites_when_zeroed_and_accumulates_otherwise():
    shapes = [[(4, 3), (4, 1)], [(3, 3), (3, 1)]]
    st = _static(shapes)
    params = [torch.zeros(s, requires_grad=True) for seg in shapes for s in seg]
    sink = _GradSink()
    sink.bind(params)
    total = sum(raster.arena_layout(st)[0])

    def backward(fill):
        out = sink.target(st, torch.device("cpu"))       # None => some .grad is still set: render into a temporary
        arena = out if out is not None else torch.empty(total)
        arena.fill_(fill)
        sink.publish(arena, st)
        return out is not None

    assert backward(1.0) is True                          # fresh: written in place into the persistent arena
    assert all(float(p.grad.min()) == 1.0 for p in params)
    assert all(p.grad.data_ptr() >= sink.arena.data_ptr() for p in params)
    assert backward(2.0) is False                         # grads still set: torch semantics = accumulate
    assert all(float(p.grad.min()) == 3.0 and float(p.grad.max()) == 3.0 for p in params)
    params[0].grad = None                                 # partially zeroed
    assert backward(5.0) is False
    assert float(params[0].grad.max()) == 5.0 and float(params[1].grad.max()) == 8.0
    for p in params:
        p.grad = None                                     # zero_grad(set_to_none=True)
    assert backward(7.0) is True
    assert all(float(p.grad.max()) == 7.0 for p in params)
    # new parameter objects (densification re-creates them): the sink rebinds and reallocates
    params2 = [torch.zeros(s, requires_grad=True) for seg in shapes for s in seg]
    old = sink.arena
    sink.bind(params2)
    assert sink.arena is None and backward(1.0) is True and sink.arena is not old
    assert all(p.grad is not None for p in params2)
