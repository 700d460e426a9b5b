"""K5: DnCNN / FDnCNN forward on the tensor cores (csrc/dncnn_tc.cuh) against PyTorch.

CPU part: the weight packing (the shared-memory image of the B operand) and the tap convention, by evaluating
the implicit GEMM exactly as the kernel indexes it.  GPU part: one 64->64 layer and whole networks through the
C ABI against float32 PyTorch convolutions of the same bf16-rounded weights.
Tolerances: a layer's output is rounded to bf16 (relative 2^-9 per value); a 17-layer network accumulates those
roundings exactly like a bf16 PyTorch module does, so the gate is the error of PyTorch's own bf16 forward.
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from pnp_admm_cnc_mri_b200 import denoisers
from pnp_admm_cnc_mri_b200 import dncnn_fused as df


def _emulate_packed_conv(x_nhwc, wp, bias, relu):
    """out[y, x, co] = sum_{kx, kc, ky, c8} in[y+ky-1, x+kx-1, 8 kc + c8] * wp[kx][kc][ky][co][c8]  (zero padding)."""
    B, H, W, _ = x_nhwc.shape
    n_out = wp.shape[3]
    xp = F.pad(x_nhwc.float(), (0, 0, 1, 1, 1, 1))
    out = torch.zeros(B, H, W, n_out)
    for ky in range(3):
        for kx in range(3):
            a = xp[:, ky:ky + H, kx:kx + W, :].reshape(B, H, W, 8, 8)              # ..., kc, c8
            out += torch.einsum('bhwkc,knc->bhwn', a, wp[kx, :, ky].float())
    out += bias.float()
    return out.clamp_min(0) if relu else out


def test_pack_conv64_matches_conv2d():
    g = torch.Generator().manual_seed(0)
    w = torch.randn(64, 64, 3, 3, generator=g).to(torch.bfloat16).float()
    b = torch.randn(64, generator=g)
    x = torch.randn(2, 5, 7, 64, generator=g).to(torch.bfloat16)
    wp = df.pack_conv64(w)
    assert wp.shape == (3, 8, 3, 64, 8) and wp.dtype == torch.bfloat16
    got = _emulate_packed_conv(x, wp, b, relu=True)
    want = F.relu(F.conv2d(x.float().permute(0, 3, 1, 2), w, b, padding=1)).permute(0, 2, 3, 1)
    assert torch.allclose(got, want, rtol=1e-5, atol=1e-4)
    # last layer: one real output row, padded to N = 16
    wt = torch.randn(1, 64, 3, 3, generator=g)
    wpt = df.pack_conv64(wt, 16)
    assert wpt.shape == (3, 8, 3, 16, 8) and float(wpt[:, :, :, 1:].abs().max()) == 0.0
    got = _emulate_packed_conv(x, wpt, torch.zeros(16), relu=False)[..., 0]
    want = F.conv2d(x.float().permute(0, 3, 1, 2), wt.to(torch.bfloat16).float(), None, padding=1)[:, 0]
    assert torch.allclose(got, want, rtol=1e-5, atol=1e-4)


def test_pack_dncnn_layer_split():
    net = denoisers.build_model('dncnn_25')
    packed, cin, n_mid = df.pack_dncnn(net, 'cpu')
    assert (cin, n_mid) == (1, 15)
    assert packed['w_mid'].shape == (15, 3, 8, 3, 64, 8) and packed['b_mid'].shape == (15, 64)
    assert packed['w_head'].shape == (64, 1, 3, 3) and packed['w_tail'].shape == (3, 8, 3, 16, 8)
    net2 = denoisers.build_model('fdncnn_gray')
    _, cin2, n_mid2 = df.pack_dncnn(net2, 'cpu')
    assert (cin2, n_mid2) == (2, 18)
    p3, cin3, n_mid3 = df.pack_dncnn(denoisers.build_model('ircnn_gray'), 'cpu')      # dilated middle layers (IRCNN)
    assert (cin3, n_mid3) == (1, 5) and p3['dilations'] == [2, 3, 4, 3, 2]
    bad = denoisers.build_model('ircnn_gray')
    bad.model[0].dilation = (2, 2); bad.model[0].padding = (2, 2)
    with pytest.raises(ValueError):
        df.conv_layers(bad)                                      # the first layer must be undilated


def _simulate_input_stationary_schedule(strip_rows, blocks=8):
    """Model of the MMA warp's bookkeeping in conv64_tc_kernel (csrc/dncnn_tc.cuh): input row i of a strip of R output
    rows feeds the output rows j = i - dy; output row t (running count) accumulates in TMEM block (-t) mod 8; a window of
    consecutive dy taps is one instruction unless it wraps around the block ring.  Returns (per output row: the set of
    (input row, dy) contributions it received, instructions issued per input row)."""
    blk = lambda t: (blocks - (t % blocks)) % blocks
    acc = {b: None for b in range(blocks)}          # block -> (owner t, set of contributions)
    done, n_instr = {}, []
    t_base = 0
    for R in strip_rows:
        for i in range(R + 2):
            if i < R:                                # first contribution to output row t_base + i: its block must be free
                b = blk(t_base + i)
                assert acc[b] is None, 'block reused before it was drained'
                acc[b] = (t_base + i, set())
            dy_hi = min(i, 2)
            dy = max(i - (R - 1), 0)
            count = 0
            while dy <= dy_hi:
                b = blk(t_base + i - dy)
                n = min(dy_hi - dy + 1, blocks - b)
                for k in range(n):                   # one instruction of N = 64 n: blocks b .. b + n - 1 <-> taps dy .. dy + n - 1
                    owner, contrib = acc[b + k]
                    assert owner == t_base + i - (dy + k), 'window block does not belong to the intended output row'
                    contrib.add((i, dy + k))
                count += 1
                dy += n
            n_instr.append(count)
            if i >= 2:                               # output row i - 2 is complete: the epilogue drains and frees its block
                b = blk(t_base + i - 2)
                owner, contrib = acc[b]
                done[owner] = (i - 2, contrib)
                acc[b] = None
        t_base += R
    assert all(v is None for v in acc.values())
    return done, n_instr


@pytest.mark.parametrize('strips', [[64, 64, 64, 64], [1], [2, 1, 3], [5, 7, 64, 2], [64] * 9])
def test_input_stationary_schedule_bookkeeping(strips):
    done, n_instr = _simulate_input_stationary_schedule(strips)
    assert len(done) == sum(strips)
    for t, (j, contrib) in done.items():
        # output row j of its strip = taps dy = 0, 1, 2 applied to the input rows j, j + 1, j + 2 (strip-local, row 0 = y0 - 1)
        assert contrib == {(j + dy, dy) for dy in range(3)}, (t, j, contrib)
    assert max(n_instr) <= 2 and min(n_instr) >= 1          # a window is split only where it wraps around the ring
    if strips == [64] * 9:
        assert sum(c == 2 for c in n_instr) / len(n_instr) < 0.3


# ------------------------------------------------------------------------------------------------ GPU
def _rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())


@pytest.mark.gpu
@pytest.mark.parametrize('shape', [(1, 8, 128), (2, 64, 256), (3, 70, 200), (1, 3, 40), (2, 130, 129), (1, 256, 256)])
@pytest.mark.parametrize('relu', [True, False])
def test_conv64_layer_gpu(shape, relu):
    B, H, W = shape
    g = torch.Generator().manual_seed(B * 1000 + H + W)
    x = torch.randn(B, H, W, 64, generator=g).to(torch.bfloat16).cuda()
    w = (torch.randn(64, 64, 3, 3, generator=g) / 24.0).to(torch.bfloat16).float().cuda()
    b = torch.randn(64, generator=g).cuda()
    got = df.conv64(x, w, b, relu=relu).float()
    with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
        want = F.conv2d(x.float().permute(0, 3, 1, 2), w, b, padding=1).permute(0, 2, 3, 1)
    if relu:
        want = want.clamp_min(0)
    # fp32 accumulation of exact bf16 products, output rounded once to bf16: |err| <= 2^-9 |want| + accumulation-order noise
    err = (got - want).abs()
    assert float((err - (2.0 ** -8) * want.abs()).max()) < 2e-3, float(err.max())
    assert _rel(got, want) < 3e-3


@pytest.mark.gpu
@pytest.mark.parametrize('dil', [2, 3, 4])
@pytest.mark.parametrize('shape', [(2, 64, 256), (3, 70, 200), (1, 7, 40), (2, 130, 129), (1, 256, 256)])
def test_conv64_dilated_layer_gpu(shape, dil):
    """IRCNN's dilated 64 -> 64 layers (dilation = padding = 2, 3, 4) on the same kernel: strips of row sub-images, tap shift
    of `dil` slots along x.  Same gate as the undilated layer."""
    B, H, W = shape
    g = torch.Generator().manual_seed(B * 1000 + H + W + dil)
    x = torch.randn(B, H, W, 64, generator=g).to(torch.bfloat16).cuda()
    w = (torch.randn(64, 64, 3, 3, generator=g) / 24.0).to(torch.bfloat16).float().cuda()
    b = torch.randn(64, generator=g).cuda()
    got = df.conv64(x, w, b, relu=True, dilation=dil).float()
    with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
        want = F.conv2d(x.float().permute(0, 3, 1, 2), w, b, padding=dil, dilation=dil).permute(0, 2, 3, 1).clamp_min(0)
    err = (got - want).abs()
    assert float((err - (2.0 ** -8) * want.abs()).max()) < 2e-3, float(err.max())
    assert _rel(got, want) < 3e-3


@pytest.mark.gpu
@pytest.mark.parametrize('name,shape', [('dncnn_25', (4, 256, 256)), ('dncnn_25', (2, 96, 160)), ('dncnn3', (1, 64, 64)),
                                        ('fdncnn_gray', (2, 128, 128)), ('ircnn_gray', (2, 256, 256)), ('ircnn_gray', (3, 70, 200))])
def test_dncnn_forward_gpu(name, shape):
    B, H, W = shape
    net = denoisers.build_model(name, seed=3)
    # random-init weights scaled up so that n(x) is not vanishingly small after 17-20 layers
    with torch.no_grad():
        for c in df.conv_layers(net):
            c.weight.mul_(1.6)
            c.weight.copy_(c.weight.to(torch.bfloat16).float())
            c.bias.copy_(c.bias.to(torch.bfloat16).float())
    net = net.cuda()
    g = torch.Generator().manual_seed(11)
    cin = df.conv_layers(net)[0].in_channels
    x = torch.rand(B, cin, H, W, generator=g).cuda()
    x = x.to(torch.bfloat16).float()             # the kernels round the network input to bf16, like net.to(bf16)(x.to(bf16))
    fused = df.FusedDnCNN(net, residual=(cin == 1))
    got = fused(x)
    with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
        want = net(x)                                                      # fp32 PyTorch, same (bf16-representable) weights
        ref16 = net.to(torch.bfloat16)(x.to(torch.bfloat16)).float()       # PyTorch's own bf16 forward
    assert got.shape == want.shape == (B, 1, H, W)
    n_got = (x[:, :1] - got) if cin == 1 else got                          # the network part n(x)
    n_want = (x[:, :1] - want) if cin == 1 else want
    n_ref16 = (x[:, :1] - ref16) if cin == 1 else ref16
    e_ours, e_torch = _rel(n_got, n_want), _rel(n_ref16, n_want)
    assert e_ours < max(2.0 * e_torch, 2e-2), (e_ours, e_torch)
    assert _rel(got, want) < 1e-2


def _forward_with_bf16_rounding_points(net, x):
    """fp32 PyTorch convolutions with the roundings K5 applies: network input, weights and biases in bf16, every activation
    rounded to bf16 after the ReLU, fp32 accumulation, fp32 last layer and residual."""
    convs = df.conv_layers(net)
    h = x.to(torch.bfloat16).float()
    with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
        for k, c in enumerate(convs):
            h = F.conv2d(h, c.weight.to(torch.bfloat16).float(), c.bias.to(torch.bfloat16).float(), padding=c.padding, dilation=c.dilation)
            if k < len(convs) - 1:
                h = F.relu(h).to(torch.bfloat16).float()
    return h


@pytest.mark.gpu
@pytest.mark.parametrize('name,shape', [('dncnn_25', (2, 256, 256)), ('fdncnn_gray', (1, 96, 160)), ('ircnn_gray', (2, 131, 256))])
def test_dncnn_forward_matches_same_rounding_points_gpu(name, shape):
    """Against an fp32 evaluation that rounds where K5 rounds, only the fp32 summation order is left, and the bf16 rounding
    flips it causes (a pre-rounding difference of 1e-6 relative moves ~1e-4 of the activations across a bf16 boundary per
    layer, each by 2^-8 relative).  Measured on B200: rel-L2 of n(x) 5.0e-4 (DnCNN-17), 1.4e-3 (FDnCNN-20); gate 5e-3.
    (PyTorch's own bf16 forward is 3.6e-2 from fp32 on the same network.)"""
    B, H, W = shape
    net = denoisers.build_model(name, seed=5).cuda()       # default init (a contraction: rounding differences do not amplify)
    cin = df.conv_layers(net)[0].in_channels
    x = torch.rand(B, cin, H, W, generator=torch.Generator().manual_seed(3)).cuda()
    got = df.FusedDnCNN(net, residual=(cin == 1))(x)
    with torch.no_grad():
        n_want = _forward_with_bf16_rounding_points(net, x)
    n_got = (x[:, :1] - got) if cin == 1 else got
    err = _rel(n_got, n_want)
    print('rel-L2 of n(x) against the same-rounding-points evaluation:', err)
    assert err < 5e-3, err


@pytest.mark.gpu
def test_denoiser_dispatch_uses_fused_gpu():
    d = denoisers.build_denoiser('dncnn_25', seed=1)
    assert d.fused is not None
    x = torch.rand(2, 1, 64, 64, device='cuda')
    y = d(x, 0)
    d_ref = denoisers.build_denoiser('dncnn_25', seed=1, fused=False)
    y_ref = d_ref(x, 0)
    assert _rel(y, y_ref) < 1e-2


def test_conv_plan_invariants_cpu():
    """Host logic of the K5 work decomposition (conv_geometry through pnpadmm_debug_conv_plan, no device): every strip of every
    row sub-image is non-empty, the items cover the image exactly once, and the chosen strips never cost more waves x rows than
    the fixed 64-row strips of round 1."""
    import ctypes
    from pnp_admm_cnc_mri_b200 import _abi
    lib = _abi.load()
    f = lib.pnpadmm_debug_conv_plan
    f.restype = ctypes.c_int
    f.argtypes = [ctypes.c_int] * 5 + [ctypes.POINTER(ctypes.c_int)] * 3
    sms = 148
    for B, H, W, d in [(1, 256, 256, 1), (15, 256, 256, 1), (256, 256, 256, 1), (256, 128, 128, 1), (3, 70, 200, 1), (2, 45, 150, 3),
                       (1, 7, 40, 4), (64, 288, 288, 2), (1, 3, 40, 1), (5, 1024, 1024, 1), (2, 130, 129, 4)]:
        xt, ys, it = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        assert f(B, H, W, d, sms, xt, ys, it) == 0
        xt, ys, it = xt.value, ys.value, it.value
        assert xt == (W + 127) // 128 and ys >= 1 and it == B * xt * ys * d
        covered = 0
        for py in range(d):
            sub = (H - py + d - 1) // d
            for k in range(ys):                                    # tc::decode_item's partition of a sub-image
                j0, j1 = k * sub // ys, (k + 1) * sub // ys
                assert j1 > j0, (B, H, W, d, py, k)
                covered += j1 - j0
        assert covered == H
        waves = lambda n: -(-n // sms)                             # noqa: E731
        sub_max = -(-H // d)
        cost = waves(it) * (-(-sub_max // ys) + 2)
        ys64 = max(1, -(-sub_max // 64))
        if ys64 <= H // d:
            assert cost <= waves(B * xt * ys64 * d) * (-(-sub_max // ys64) + 2), (B, H, W, d, ys, ys64)
    bad = ctypes.c_int()
    assert f(1, 3, 40, 4, sms, bad, bad, bad) != 0                # image shorter than the dilation: refused


# ---------------------------------------------------------------------------------------- FFDNet on the same kernels
def test_pack_ffdnet_layer_split():
    """CPU: the first layer of an FFDNet becomes a 64 -> 64 image with input channels 5..63 zero; the tail keeps 4 outputs."""
    net = denoisers.build_model('ffdnet_gray', seed=2)
    packed, n_mid = df.pack_ffdnet(net, 'cpu')
    assert n_mid == 13
    convs = [m for m in net.model if isinstance(m, torch.nn.Conv2d)]
    wh = packed['w_head'].float()                      # [kx][c_in // 8][ky][c_out][c_in % 8]
    assert wh.shape == (3, 8, 3, 64, 8)
    assert float(wh[:, 1:].abs().max()) == 0.0 and float(wh[:, 0, :, :, 5:].abs().max()) == 0.0
    want = convs[0].weight.detach().to(torch.bfloat16).float()          # [64][5][3][3]
    assert torch.equal(wh[:, 0, :, :, :5].permute(2, 3, 1, 0), want)   # -> [c_out][c_in][ky][kx]
    wt = packed['w_tail'].float()
    assert wt.shape == (3, 8, 3, 16, 8) and float(wt[:, :, :, 4:].abs().max()) == 0.0
    assert packed['b_tail'].shape == (4,)


def _ffdnet_with_bf16_rounding_points(net, x, sigma):
    """fp32 convolutions with the roundings the kernels apply (input, map, weights, biases, activations after each ReLU)."""
    h, w = x.shape[-2:]
    xp = F.pad(x, (0, (-w) % 2, 0, (-h) % 2), mode='replicate')
    t = F.pixel_unshuffle(xp.to(torch.bfloat16).float(), 2)
    m = torch.full_like(t[:, :1], float(torch.tensor(sigma).to(torch.bfloat16)))
    t = torch.cat((t, m), 1)
    convs = [c for c in net.model if isinstance(c, torch.nn.Conv2d)]
    with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
        for k, c in enumerate(convs):
            t = F.conv2d(t, c.weight.to(torch.bfloat16).float(), c.bias.to(torch.bfloat16).float(), padding=1)
            if k < len(convs) - 1:
                t = F.relu(t).to(torch.bfloat16).float()
    return F.pixel_shuffle(t, 2)[..., :h, :w]


@pytest.mark.gpu
@pytest.mark.parametrize('shape', [(2, 256, 256), (3, 70, 200), (1, 65, 131), (1, 16, 16)])
def test_ffdnet_forward_gpu(shape):
    """FFDNet through pnpadmm_ffdnet_forward_bf16 (pack -> thin first layer -> 13 x conv64 -> shuffled four-channel tail)
    against the fp32 module with the same bf16-representable weights, PyTorch's own bf16 forward as the yardstick, and the
    same-rounding-points evaluation (odd sizes exercise the replicate padding and the crop)."""
    B, H, W = shape
    net = denoisers.build_model('ffdnet_gray', seed=4)
    with torch.no_grad():
        for c in net.model:
            if isinstance(c, torch.nn.Conv2d):
                c.weight.mul_(1.5)
                c.weight.copy_(c.weight.to(torch.bfloat16).float())
                c.bias.copy_(c.bias.to(torch.bfloat16).float())
    net = net.cuda()
    x = torch.rand(B, 1, H, W, generator=torch.Generator().manual_seed(H + W)).cuda().to(torch.bfloat16).float()
    sigma = 15 / 255.
    got = df.FusedFFDNet(net)(x, sigma)
    sig = torch.full((1, 1, 1, 1), sigma, device='cuda')
    with torch.no_grad(), torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
        want = net(x, sig.to(torch.bfloat16).float())
        ref16 = net.to(torch.bfloat16)(x.to(torch.bfloat16), sig.to(torch.bfloat16)).float()
        net = net.float()
        same = _ffdnet_with_bf16_rounding_points(net, x, sigma)
    assert got.shape == want.shape == (B, 1, H, W)
    e_ours, e_torch, e_same = _rel(got, want), _rel(ref16, want), _rel(got, same)
    print(f'FFDNet {shape}: ours vs fp32 {e_ours:.2e}, torch bf16 vs fp32 {e_torch:.2e}, ours vs same rounding points {e_same:.2e}')
    assert e_ours < max(2.0 * e_torch, 2e-2), (e_ours, e_torch)
    assert e_same < 5e-3, e_same


@pytest.mark.gpu
def test_denoiser_dispatch_ffdnet_fused_gpu():
    d = denoisers.build_denoiser('ffdnet_gray', seed=1)
    assert d.fused is not None
    x = torch.rand(2, 1, 96, 64, device='cuda')
    y = d(x, 0)
    y_ref = denoisers.build_denoiser('ffdnet_gray', seed=1, fused=False)(x, 0)
    assert y.shape == y_ref.shape and _rel(y, y_ref) < 2e-2


@pytest.mark.gpu
def test_denoiser_dispatch_ircnn_fused_weight_switch_gpu():
    """IRCNN through the Denoiser with the 25-set sigma-indexed weight switch (S3:280-288): the tensor-core path re-packs its
    weights when the index changes and follows the stock bf16 module (same sets) at every iteration probed."""
    sets = {str(k): denoisers.build_model('ircnn_gray', seed=100 + k).state_dict() for k in range(25)}
    d = denoisers.build_denoiser('ircnn_gray', seed=1, iter_num=50, ircnn_weights=sets)
    d_ref = denoisers.build_denoiser('ircnn_gray', seed=1, iter_num=50, ircnn_weights=sets, fused=False)
    d32 = denoisers.build_denoiser('ircnn_gray', seed=1, iter_num=50, ircnn_weights=sets, dtype=torch.float32)
    assert d.fused is not None and d_ref.fused is None
    x = torch.rand(2, 1, 96, 160, device='cuda')
    seen = set()
    for i in (0, 1, 10, 25, 49):
        y, y16, y32 = d(x, i), d_ref(x, i), d32(x, i)
        seen.add(d._ircnn_idx)
        assert d._ircnn_idx == d_ref._ircnn_idx == d32._ircnn_idx
        n, n16, n32 = x - y, x - y16, x - y32
        assert _rel(n, n32) < max(2.0 * _rel(n16, n32), 2e-2), (i, _rel(n, n32), _rel(n16, n32))
    assert len(seen) >= 4
