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


def test_full_arena_sink_places_a_frame_inside_the_model_layout():
    """Data-parallel sink: the arena has the layout of ALL sub-models; a frame that sees sub-models [0, 2] writes their
    slices through per-tensor offsets, the absent sub-model's slice is zeroed (it may hold the other replicas' sum)."""
    import numpy as np
    from street_gaussians_ns_b200.model import _FullArenaSink
    shapes6 = lambda n, F: [(n, 3), (n, 3), (n, 4), (n, F, 3), (n, 15, 3), (n, 1)]  # noqa: E731
    model_params = [[torch.zeros(s, requires_grad=True) for s in shapes6(n, F)] for n, F in ((5, 1), (3, 5), (2, 5))]
    sink = _FullArenaSink()
    sink.bind_model(model_params, [0, 2])
    sizes = np.array([(p.numel() + 3) // 4 * 4 for ps in model_params for p in ps])
    offs = np.concatenate([[0], np.cumsum(sizes)[:-1]])
    assert list(sink.grad_offsets(None)) == list(offs[[0, 1, 2, 3, 4, 5, 12, 13, 14, 15, 16, 17]])
    arena = sink.target(None, torch.device("cpu"))
    assert arena.numel() == sizes.sum() and float(arena.abs().sum()) == 0.0
    arena.fill_(3.0)                                   # "all-reduce result of the previous step"
    again = sink.target(None, torch.device("cpu"))     # next step: same arena, the absent sub-model's slice zeroed
    assert again is arena
    lo, hi = int(offs[6]), int(offs[12])
    assert float(arena[lo:hi].abs().sum()) == 0.0 and float(arena[:lo].min()) == 3.0 and float(arena[hi:].min()) == 3.0
    sink.publish(arena, None)
    for i in (0, 2):
        for k, p in enumerate(model_params[i]):
            assert p.grad is not None and p.grad.shape == p.shape
            assert p.grad.data_ptr() == arena.data_ptr() + 4 * int(offs[6 * i + k])
    assert all(p.grad is None for p in model_params[1])
    # grads still set -> accumulate through a temporary arena in the FRAME's layout, like the plain sink
    assert sink.target(None, torch.device("cpu")) is None
    # a replica that rendered nothing contributes zeros in the common layout
    for ps in model_params:
        for p in ps:
            p.grad = None
    sink.bind_model(model_params, [])
    z = sink.target(None, torch.device("cpu"))
    assert z is arena and float(z.abs().sum()) == 0.0
    # different order of the visible sub-models (annotation order): offsets follow the frame's order
    sink.bind_model(model_params, [0, 2, 1])
    assert list(sink.grad_offsets(None)[6:12]) == list(offs[12:18])
