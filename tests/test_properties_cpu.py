"""Property tests (hypothesis) of the host-side index logic around the hot path: clip windows, frame shards, caption buckets,
the packed feature store and the chunk planner of the host pipeline."""
import numpy as np
import torch
from hypothesis import given, settings
from hypothesis import strategies as st

from tennis_b200.dataset import window_frames
from tennis_b200.parallel import plan_chunks, shard_range
from tennis_b200.utils.captioning import FixedBucketSampler, pad_stack


@settings(max_examples=200, deadline=None)
@given(center=st.integers(0, 5000), window=st.integers(1, 64), stride=st.integers(1, 8), every=st.integers(1, 16),
       length=st.integers(40, 6000))
def test_window_frames_properties(center, window, stride, every, length):
    """dataset.py:190-201: W offsets, clamped into [0, last 'every' frame]; non-decreasing; unclamped entries are center+off*stride."""
    center = min(center, length - 1)
    fr = window_frames(center, window, stride, every, length)
    assert len(fr) == window
    max_frame = length - every
    max_frame -= max_frame % every
    assert all(0 <= f <= max(0, max_frame) for f in fr)
    assert all(a <= b for a, b in zip(fr, fr[1:]))
    offs = list(range(int(-window / 2), int(np.ceil(window / 2))))
    for f, o in zip(fr, offs):
        raw = center + o * stride
        if 0 < raw < max_frame:
            assert f == raw


@settings(max_examples=200, deadline=None)
@given(n=st.integers(0, 100000), world=st.integers(1, 16))
def test_shard_range_partitions_the_frames(n, world):
    """split_and_load(even_split=False): contiguous shards, equal size except the last, covering [0, n) exactly once."""
    spans = [shard_range(n, r, world) for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][1] == n
    assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    sizes = [hi - lo for lo, hi in spans]
    assert all(s == n // world for s in sizes[:-1]) and sizes[-1] == n - (world - 1) * (n // world)


@settings(max_examples=100, deadline=None)
@given(lengths=st.lists(st.tuples(st.integers(1, 300), st.integers(3, 52)), min_size=1, max_size=200),
       batch=st.integers(1, 64), buckets=st.integers(1, 8), shuffle=st.booleans())
def test_bucket_sampler_covers_every_sample_once(lengths, batch, buckets, shuffle):
    s = FixedBucketSampler(lengths, batch, buckets, shuffle=shuffle)
    seen = [i for b in s for i in b]
    assert sorted(seen) == list(range(len(lengths))) and all(1 <= len(b) <= batch for b in s) and len(s) == len(list(s))


@settings(max_examples=100, deadline=None)
@given(lens=st.lists(st.integers(0, 40), min_size=1, max_size=12), width=st.integers(1, 5))
def test_pad_stack_pads_with_zero_to_the_longest(lens, width):
    seqs = [torch.full((n, width), float(i + 1)) for i, n in enumerate(lens)]
    out = pad_stack(seqs)
    assert tuple(out.shape) == (len(lens), max(lens), width)
    for i, n in enumerate(lens):
        assert (out[i, :n] == i + 1).all() and (out[i, n:] == 0).all()


@settings(max_examples=100, deadline=None)
@given(B=st.integers(1, 256), copy_ms=st.floats(0.01, 2.0), fixed_ms=st.floats(0.0, 2.0), compute_ms=st.floats(0.01, 2.0),
       max_chunks=st.integers(1, 8))
def test_host_pipeline_chunk_plan_is_a_partition(B, copy_ms, fixed_ms, compute_ms, max_chunks):
    """HostPipeline's chunk plan: increasing cut points from 0 to B, at most max_chunks chunks, never an empty chunk."""
    cuts = plan_chunks(B, copy_ms, fixed_ms, compute_ms, max_chunks=max_chunks)
    assert cuts[0] == 0 and cuts[-1] == B and len(cuts) - 1 <= max_chunks
    assert all(a < b for a, b in zip(cuts, cuts[1:]))


def test_fixed_bucket_sampler_reshuffles_inside_buckets_every_epoch():
    """gluonnlp's FixedBucketSampler(shuffle=True) draws new batches every epoch: the batch MEMBERSHIP must change
    between epochs, every sample must appear exactly once per epoch, and a batch never mixes buckets."""
    from tennis_b200.utils.captioning import FixedBucketSampler
    lengths = [(5 + (i * 7) % 40, 3 + (i * 3) % 20) for i in range(200)]
    s = FixedBucketSampler(lengths, batch_size=16, num_buckets=4, shuffle=True, seed=1)
    e1, e2 = [list(b) for b in s], [list(b) for b in s]
    assert sorted(i for b in e1 for i in b) == list(range(200)) == sorted(i for b in e2 for i in b)
    assert len(e1) == len(s) == len(e2)
    assert set(map(frozenset, e1)) != set(map(frozenset, e2))
    keys = [max(l) for l in lengths]
    lo, hi = min(keys), max(keys)
    width = -(-(hi - lo + 1) // 4)
    for b in e1:
        assert len({min((keys[i] - lo) // width, 3) for i in b}) == 1
    s0 = FixedBucketSampler(lengths, batch_size=16, num_buckets=4, shuffle=False)
    assert [list(b) for b in s0] == [list(b) for b in s0]


@settings(max_examples=40, deadline=None)
@given(n=st.integers(1, 3), h=st.integers(1, 9), w=st.integers(1, 14), cin=st.sampled_from([8, 16]), cout=st.sampled_from([8, 24]),
       seed=st.integers(0, 10 ** 6))
def test_conv3x3_tap_lists_for_any_geometry(n, h, w, cin, cout, seed):
    """tcgemm.taps_conv3x3_* (forward, data gradient, weight gradient with the 8-pixel pitch and dx-shifted copies) reproduce
    torch's conv2d and its autograd for arbitrary frame counts and map sizes, including 1-pixel maps and widths whose padded pitch
    is already a multiple of 8 (fp64 emulation of the documented tn_split_bf16 / tn_gemm_tc index semantics)."""
    import torch.nn.functional as F
    from test_host_cpu import _emu_gemm, _emu_planes, _unpad
    from tennis_b200 import tcgemm
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, cin, h, w, generator=g, dtype=torch.float64, requires_grad=True)
    wt = torch.randn(cout, cin, 3, 3, generator=g, dtype=torch.float64, requires_grad=True)
    dy = torch.randn(n, cout, h, w, generator=g, dtype=torch.float64)
    y = F.conv2d(x, wt, padding=1)
    (y * dy).sum().backward()
    x2 = x.detach().permute(0, 2, 3, 1).reshape(-1, cin)
    dy2 = dy.permute(0, 2, 3, 1).reshape(-1, cout)
    Mp = tcgemm.padded_rows(n, h, w)
    Y = _emu_gemm(_emu_planes(x2, pad_hw=(h, w)), wt.detach().permute(0, 2, 3, 1).reshape(cout, 9 * cin), Mp, cout, cin,
                  tcgemm.taps_conv3x3_forward(w, cin))
    assert torch.allclose(_unpad(Y, n, h, w), y.detach().permute(0, 2, 3, 1).reshape(-1, cout), atol=1e-9)
    DX = _emu_gemm(_emu_planes(dy2, pad_hw=(h, w)), wt.detach().permute(1, 2, 3, 0).reshape(cin, 9 * cout), Mp, cin, cout,
                   tcgemm.taps_conv3x3_dgrad(w, cout))
    assert torch.allclose(_unpad(DX, n, h, w), x.grad.permute(0, 2, 3, 1).reshape(-1, cin), atol=1e-9)
    P8 = tcgemm.pitch8(w)
    xT = _emu_planes(x2, transpose=True, pad_hw=(h, w), pitch=P8)
    dyT = torch.cat([_emu_planes(dy2, transpose=True, pad_hw=(h, w), pitch=P8, shift=dx) for dx in (-1, 0, 1)], 0)
    Dt = _emu_gemm(xT, dyT, cin, cout, tcgemm.padded_rows(n, h, w, P8), tcgemm.taps_conv3x3_wgrad(w, cout), tile_taps=True)
    dwk = torch.stack([d.t() for d in Dt], 1).reshape(cout, 3, 3, cin).permute(0, 3, 1, 2)
    assert torch.allclose(dwk, wt.grad, atol=1e-8)


@settings(max_examples=50, deadline=None)
@given(seed=st.integers(0, 10 ** 6), scale=st.sampled_from([1e-6, 1e-2, 1.0, 37.0, 1e4]))
def test_split_bf16_representation_and_three_product_error(seed, scale):
    """The arithmetic of the split-bf16 modes (inference `precision='split_bf16'`, training `TN_TRAIN_GEMM=x3`): x = hi + lo with
    hi = bf16(x), lo = bf16(x - hi) represents x to 2^-16 relative (DESIGN.md states 2^-18 typical), and the three products
    hi*hi' + hi*lo' + lo*hi' accumulated in fp32 reproduce a K = 256 dot product to 1e-4 of sum|x||y| -- against 4e-3 for plain
    bf16 operands."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(64, 256, generator=g) * scale
    y = torch.randn(256, generator=g)
    def split(t):
        hi = t.bfloat16().float()
        return hi, (t - hi).bfloat16().float()
    xh, xl = split(x)
    yh, yl = split(y)
    assert ((xh + xl) - x).abs().max().item() <= 2.0 ** -16 * x.abs().max().item()
    ref = x.double() @ y.double()
    bound = (x.abs().double() @ y.abs().double())
    x3 = (xh @ yh + xh @ yl + xl @ yh).double()
    x1 = (xh @ yh).double()
    assert ((x3 - ref).abs() <= 1e-4 * bound).all()
    assert (x3 - ref).abs().max() <= (x1 - ref).abs().max() + 1e-12
