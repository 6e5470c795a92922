"""CPU: host-side logic added around the kernels -- the gradient-arena layout and the FanIn accumulation protocol
(autograd_conv.FanIn) with the dgrad launch replaced by a torch stand-in.  No CUDA needed."""
import torch

from efficientteacher_b200.autograd_conv import FanIn, _empty_cl, _inplace_nhwc
from efficientteacher_b200.parallel import GradArena


def test_grad_arena_views_are_16_byte_aligned_and_disjoint():
    ps = [torch.nn.Parameter(torch.randn(s)) for s in ((255,), (64, 32, 1, 1), (7,), (32,), (3, 85, 64))]
    arena = GradArena(ps)
    assert arena.check_views()
    seen = set()
    for p, o in zip(arena.params, arena.offsets):
        assert o % 4 == 0 and p.grad.data_ptr() == arena.flat.data_ptr() + 4 * o and p.grad.shape == p.shape
        span = set(range(o, o + p.numel()))
        assert not (span & seen)
        seen |= span
    for p in ps:
        p.grad.fill_(1.0)
    assert float(arena.flat.sum()) == sum(p.numel() for p in ps)       # the pads stay zero
    arena.zero()
    assert float(arena.flat.abs().sum()) == 0.0 and arena.check_views()


def _fake_dgrad(value):
    """stand-in for etb_conv_dgrad: writes (or adds) a constant into the NHWC view it is given"""
    def run(out_nhwc, out_cstride, accumulate):
        if accumulate:
            out_nhwc += value
        else:
            out_nhwc.fill_(value)
    return run


def test_fanin_two_convolutions():
    f = FanIn()
    f.n = 2
    assert f.add_dgrad(_fake_dgrad(1.5), 2, 8, 4, 4, "cpu") is None           # first contribution: becomes the buffer
    out = f.add_dgrad(_fake_dgrad(2.0), 2, 8, 4, 4, "cpu")                     # second: accumulated in place, handed out
    assert out is not None and out.shape == (2, 8, 4, 4) and torch.all(out.float() == 3.5)
    assert f.buf is None


def test_fanin_shortcut_then_convolution_accumulates_into_the_incoming_gradient():
    f = FanIn()
    f.n = 2
    whole = _empty_cl(2, 16, 4, 4, "cpu")
    whole.fill_(1.0)
    da = whole[:, 8:16]                                  # a channel slice of a wider gradient (JoinFn backward hands out such views)
    assert _inplace_nhwc(da, 8) is not None
    assert f.put(da) is None
    out = f.add_dgrad(_fake_dgrad(0.25), 2, 8, 4, 4, "cpu")
    assert out.data_ptr() == da.data_ptr() and torch.all(out.float() == 1.25)
    assert torch.all(whole[:, :8].float() == 1.0)        # the neighbouring slice is untouched


def test_fanin_mixed_with_a_non_participating_consumer_and_odd_layouts():
    # only one participant registered: its gradient goes straight through (autograd sums the rest)
    f = FanIn()
    f.n = 1
    out = f.add_dgrad(_fake_dgrad(4.0), 1, 8, 2, 2, "cpu")
    assert out is not None and torch.all(out.float() == 4.0)
    # a pass-through gradient in a layout the kernel cannot address in place (NCHW-contiguous fp32): out-of-place fallback
    f = FanIn()
    f.n = 2
    g = torch.ones(1, 8, 2, 2)
    assert _inplace_nhwc(g, 8) is None
    assert f.put(g) is None
    out = f.add_dgrad(_fake_dgrad(2.0), 1, 8, 2, 2, "cpu")
    assert torch.allclose(out.float(), torch.full((1, 8, 2, 2), 3.0)) and torch.all(g == 1.0)
    # two pass-through contributions
    f = FanIn()
    f.n = 2
    assert f.put(torch.ones(1, 8, 2, 2)) is None
    assert torch.all(f.put(torch.ones(1, 8, 2, 2)) == 2.0)
