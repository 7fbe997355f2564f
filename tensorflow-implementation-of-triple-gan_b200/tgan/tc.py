"""bf16 tensor-core (tcgen05 / TMEM / TMA) routing of the dense contractions.

`ops.conv2d` / `ops.conv2d_transpose` call in here when ctx.math == 'bf16'.  Every contraction is expressed
as a list of taps (dy, dx) over shifted, optionally strided windows of an NHWC bf16 tensor:

  conv fprop        x window (oy*s + r - pt, ox*s + c - pl)                         -> tgan_igemm_bf16
  conv dgrad s=1    dz window (y + pt - r, x + pl - c), weights with roles swapped  -> tgan_igemm_bf16
  conv dgrad s=2    4 output-parity classes of the transposed conv                  -> 4 x tgan_igemm_bf16
  conv wgrad        pixel axis = GEMM K, MN-major operands                          -> tgan_wgrad_bf16
  deconv fprop      4 output-parity classes (5x5/s2: 4/6/6/9 taps)                  -> 4 x tgan_igemm_bf16
  deconv dgrad      strided forward conv of dy                                      -> tgan_igemm_bf16
  deconv wgrad      wgrad with operand roles exchanged                              -> tgan_wgrad_bf16
  dense / NiN / ZCA 1 tap, H = 1                                                    -> same kernels

Skinny layers (Cout < 16: RGB output, logits, D head) stay on the fp32 SIMT path.
"""
import contextlib
import ctypes

import torch

from . import _lib
from .core import BF16, ctx, dt_code


def _st():
    return torch.cuda.current_stream().cuda_stream


def _new(shape, dtype):
    return torch.empty(tuple(int(s) for s in shape), dtype=dtype, device=ctx.device)


def _bf16_padded(t, rows, C, ld):
    """-> (bf16 tensor [rows, ld8], ld8) with ld8 % 8 == 0 and zero padding (TMA needs 16-byte pixel strides)."""
    if t.dtype == torch.bfloat16 and ld % 8 == 0:
        return t, ld
    ld8 = (C + 7) // 8 * 8
    o = _new((rows, ld8), torch.bfloat16)
    _lib.call('tgan_concat_label', t.data_ptr(), dt_code(t), rows, C, ld, t.data_ptr(), 0, 1, o.data_ptr(), BF16, ld8,
              _st())
    return o, ld8


def _wcache(w):
    c = w.key.cache.setdefault('tc', {})
    ver = ctx.store.group_version(w.key.group)
    if c.get('version') != ver:
        c.clear()
        c['version'] = ver
    return c


_taps_dev = {}


def _taps_tensor(taps):
    key = tuple(taps)
    if key not in _taps_dev:
        from .prep import to_device_table
        _taps_dev[key] = to_device_table(torch.tensor(list(taps), dtype=torch.int32))
    return _taps_dev[key]


def _pack(w, key, T, Nrows, K, st, sn, sk, taps=None, scale_mod=0):
    """Packed bf16 K-major weights [T][Nrows][Kpad] for the current optimiser version.  All operands of a network are
    (re)packed by one multi-tensor launch when its version changes (prep.PackGroup)."""
    from .prep import PackGroup
    return PackGroup.of(w.key.group).get(w, key, T, Nrows, K, st, sn, sk, None if taps is None else _taps_tensor(taps),
                                         scale_mod)


def _igemm(x, N, H, W, C, ldx, wp, Kpad, taps, Nout, gh, gw, out, OH, OW, ldo, s=1, os_=1, oo=(0, 0), vh=0, vw=0,
           colsum=None, segs=None, bias=None, act=0, classes=None, fused=None):
    a = _lib.TganIgemmArgs()
    a.x, a.N, a.H, a.W, a.C, a.ldx = x.data_ptr(), N, H, W, C, ldx
    a.wp, a.T, a.Nout, a.Kpad = wp.data_ptr(), len(taps), Nout, Kpad
    for i, (dy, dx) in enumerate(taps):
        a.dy[i], a.dx[i] = dy, dx
    a.gh, a.gw, a.sy, a.sx = gh, gw, s, s
    a.out, a.odt, a.OH, a.OW, a.ldo = out.data_ptr(), dt_code(out), OH, OW, ldo
    a.osy, a.osx, a.ooy, a.oox, a.vh, a.vw = os_, os_, oo[0], oo[1], vh, vw
    a.bias, a.colsum = (None if bias is None else bias.data_ptr()), (None if colsum is None else colsum.data_ptr())
    a.act, a.alpha = act, 1.0
    if classes:      # [(taps in class, output row offset, output col offset)]: parity classes in one launch
        a.ncls = len(classes)
        for i, (tn, oy, ox) in enumerate(classes):
            a.cls_T[i], a.cls_ooy[i], a.cls_oox[i] = tn, oy, ox
    if segs and len(segs) > 1:
        a.nseg, e = len(segs), 0
        for i, n in enumerate(segs[:-1]):
            e += n
            a.seg_end[i] = e
    if fused:      # mean-only BN fused into the epilogue (csrc/mobn_fused.cu): per-segment bias, class sums, lrelu masks
        a.bias_seg = 1 if fused.get('bias_seg') else 0
        if fused.get('clsum') is not None:
            a.clsum, a.cls_h, a.cls_w = fused['clsum'].data_ptr(), fused['cls_hw'][0], fused['cls_hw'][1]
        if fused.get('mask_out') is not None:
            a.mask_out = fused['mask_out'].data_ptr()
        if fused.get('mask_in') is not None:
            a.mask_in, a.mask_alpha = fused['mask_in'].data_ptr(), fused.get('mask_alpha', 0.2)
    _lib.call('tgan_igemm_bf16', ctypes.byref(a), _st())


def _wgrad(dz, N, gh, gw, Cout, lddz, x, H, W, Cin, ldx, taps, s, dw, cin_store=0):
    a = _lib.TganWgradArgs()
    a.dz, a.N, a.gh, a.gw, a.Cout, a.lddz = dz.data_ptr(), N, gh, gw, Cout, lddz
    a.x, a.H, a.W, a.Cin, a.ldx, a.sy, a.sx = x.data_ptr(), H, W, Cin, ldx, s, s
    a.T = len(taps)
    for i, (dy, dx) in enumerate(taps):
        a.dy[i], a.dx[i] = dy, dx
    a.dw, a.beta, a.cin_store = dw.data_ptr(), 1.0, cin_store
    ws = ctx.ws()
    a.ws, a.ws_bytes = ws.data_ptr(), ws.numel() * 4
    _lib.call('tgan_wgrad_bf16', ctypes.byref(a), _st())


# ------------------------------------------------------------------------------------------------
# convolution (tf.nn.conv2d) and dense (tf.matmul)
# ------------------------------------------------------------------------------------------------


def conv_eligible(g, x):
    # dense layers (tf.matmul) of any width >= 64 run on the tensor cores: the MNIST networks are 250 / 500 wide
    # (Good_GAN.py:93-124, :31-57); an output whose row pitch is not a multiple of 16 bytes is staged (conv_fwd)
    dense = g['kh'] == 1 and g['kw'] == 1 and g['s'] == 1 and g['Cout'] >= 64 and not os.environ.get('TGAN_NO_ODD_DENSE_TC')
    return g['Cout'] >= 16 and g['kh'] * g['kw'] <= 25 and g['s'] in (1, 2) and (g['Cout'] % 8 == 0 or dense) and \
        (g['s'] == 1 or (g['H'] % 2 == 0 and g['W'] % 2 == 0))


def _flat(g):
    """1x1/s1 convs and dense layers are plain GEMMs: view the pixels as one row of length rows."""
    if g['kh'] == 1 and g['kw'] == 1 and g['s'] == 1:
        rows = g['N'] * g['H'] * g['W']
        return dict(g, N=1, H=1, W=rows, Ho=1, Wo=rows)
    return g


import os
SMALL_CIN_MAX = int(os.environ.get('TGAN_SMALL_CIN_MAX', '16'))


def _small_cin(g):
    """few input channels (conv1_1: 3, D's first conv: 13): per-tap TMA boxes would move 6-26 useful bytes per 128-byte
    shared-memory row, so the taps are gathered once into a bf16 im2col matrix and the conv runs as one plain GEMM"""
    return g['kh'] * g['kw'] > 1 and g['C'] <= SMALL_CIN_MAX


def _im2col(x, g):
    K = g['kh'] * g['kw'] * g['C']
    Kc = (K + 7) // 8 * 8
    rows = g['N'] * g['Ho'] * g['Wo']
    col = _new((rows, Kc), torch.bfloat16)
    xd = x.data
    _lib.call('tgan_im2col_bf16', xd.data_ptr(), dt_code(xd), g['N'], g['H'], g['W'], g['C'], x.ld, g['kh'], g['kw'],
              g['s'], g['s'], g['pt'], g['pl'], g['Ho'], g['Wo'], col.data_ptr(), Kc, _st())
    return col, K, Kc


def fprop_operand(w, g):
    """-> (packed bf16 fprop weights, Kpad, tap stride, channel stride) of the operand conv_fwd reads"""
    C, Cout, kh, kw = g['C'], g['Cout'], g['kh'], g['kw']
    if _small_cin(g):
        wp, Kpad = _pack(w, 'fprop_col', 1, Cout, kh * kw * C, 0, 1, Cout)
        return wp, Kpad, C, Kpad
    wp, Kpad = _pack(w, 'fprop', kh * kw, Cout, C, C * Cout, 1, Cout)
    return wp, Kpad, Cout * Kpad, Kpad


def conv_fwd(x, w, g, colsum=None, segs=None, bias=None, act=0, ldo=None, fused=None):
    """-> bf16 [rows, ldo] with act(conv + bias) in channels [0, Cout) (ldo > Cout: the rest is left untouched)"""
    gf = _flat(g)
    C, Cout, kh, kw = g['C'], g['Cout'], g['kh'], g['kw']
    if _small_cin(g):
        col, K, Kc = _im2col(x, g)
        g['_col'] = (col, K, Kc)
        rows = col.shape[0]
        wp, Kpad = _pack(w, 'fprop_col', 1, Cout, K, 0, 1, Cout)      # HWIO flattened: k = (tap, ci) has stride Cout
        assert Kpad == Kc
        ldo = ldo or Cout
        z = _new((rows, ldo), torch.bfloat16)
        if segs:
            segs = [n * (rows // sum(segs)) for n in segs]
        _igemm(col, 1, 1, rows, K, Kc, wp, Kpad, [(0, 0)], Cout, 1, rows, z, 1, rows, ldo, colsum=colsum, segs=segs,
               bias=bias, act=act, fused=fused)
        return z
    xd, ld = _bf16_padded(x.data, x.rows, C, x.ld)
    g['_x'] = (xd, ld)
    wp, Kpad = _pack(w, 'fprop', kh * kw, Cout, C, C * Cout, 1, Cout)
    taps = [(r - g['pt'], c - g['pl']) for r in range(kh) for c in range(kw)]
    ldo = ldo or Cout
    rows = g['N'] * g['Ho'] * g['Wo']
    # TMA stores need a 16-byte row pitch: an odd width (250, 500: only ldo == Cout can be odd, concatenated outputs are
    # padded by their consumer) goes through an 8-channel-aligned staging tensor and one compaction
    lds = (ldo + 7) // 8 * 8
    z = _new((rows, lds), torch.bfloat16)
    if segs and gf is not g:          # plain GEMM: segments become ranges of rows
        rps = rows // sum(segs)
        segs = [n * rps for n in segs]
    _igemm(xd, gf['N'], gf['H'], gf['W'], C, ld, wp, Kpad, taps, Cout, gf['Ho'], gf['Wo'], z, gf['Ho'], gf['Wo'], lds,
           s=g['s'], colsum=colsum, segs=segs, bias=bias, act=act, fused=fused)
    if lds != ldo:
        zc = _new((rows, ldo), torch.bfloat16)
        _lib.call('tgan_copy_channels', z.data_ptr(), BF16, lds, zc.data_ptr(), BF16, ldo, rows, Cout, _st())
        return zc
    return z


def _ops():
    from . import ops
    return ops


def _small(rows, ch):
    """fewer 256-pixel x 128-channel tiles than SMs: the launch cannot fill the GPU on its own"""
    return not ctx.on_side and rows * ch < 148 * 256 * 128


def _grad_target(x, C):
    """-> (Var that receives the input gradient, number of leading channels it covers)"""
    src = x.aux.get('concat_src') if x.aux else None
    if src is not None:
        return src, x.aux['C0']
    return x, C


def conv_bwd(x, w, g, dz):
    from .core import add_grad
    gf = _flat(g)
    C, Cout, kh, kw, s, pt, pl = g['C'], g['Cout'], g['kh'], g['kw'], g['s'], g['pt'], g['pl']
    rows = g['N'] * g['Ho'] * g['Wo']
    dzb, lddz = _bf16_padded(dz, rows, Cout, Cout)      # (lddz > Cout: odd-width dense layers)
    tgt, Cg = _grad_target(x, C)
    # small layers (fewer tiles than SMs): the filter gradient runs on the side stream beside the input gradient
    par = w.requires_grad and tgt.requires_grad and _small(rows, max(C, Cout))
    if w.requires_grad:
        dw = w.grad_target()
        with (_ops().side_stream() if par else contextlib.nullcontext()):
            if _small_cin(g):
                col, K, Kc = g.get('_col') or _im2col(x, g)
                # dW[(tap, ci), co] = col^T dz: a plain GEMM whose [Kc][Cout] result starts with the K rows of the HWIO gradient
                _wgrad(dzb, 1, 1, rows, Cout, lddz, col, 1, rows, Kc, Kc, [(0, 0)], 1, dw, cin_store=K)
            else:
                xd, ld = g.get('_x') or _bf16_padded(x.data, x.rows, C, x.ld)
                taps = [(r - pt, c - pl) for r in range(kh) for c in range(kw)]
                # the kernel writes [t][ci][co] == HWIO
                _wgrad(dzb, gf['N'], gf['Ho'], gf['Wo'], Cout, lddz, xd, gf['H'], gf['W'], C, ld, taps, s, dw)
    if tgt.requires_grad:
        # a label-concatenated input only needs the gradient of its first Cg channels: it goes to the concat's source
        dxt = _new(tuple(x.shape[:-1]) + (Cg,), tgt.data.dtype if tgt.data.dtype == torch.bfloat16 else torch.float32)
        # narrow / fp32 gradients (the RGB image: 3 channels) go through an 8-channel bf16 staging tensor so that the
        # GEMM keeps its TMA-store epilogue, then one slice / convert
        staged = Cg % 8 != 0 or dxt.dtype != torch.bfloat16
        Cs = (Cg + 7) // 8 * 8
        dx = _new(tuple(x.shape[:-1]) + (Cs,), torch.bfloat16) if staged else dxt
        wp, Kpad = None, None
        if s == 1:
            wp, Kpad = _pack(w, ('dgrad', Cg), kh * kw, Cg, Cout, C * Cout, Cout, 1)
            taps = [(pt - r, pl - c) for r in range(kh) for c in range(kw)]
            # the producer of x was a fused mean-only-BN layer: its leaky-ReLU derivative (1-bit mask written by its forward
            # epilogue) and the per-segment sums of du = dy * lrelu'(y) are applied / accumulated in THIS epilogue
            mk = tgt.aux.get('mask') if (tgt.aux and tgt is x and gf is g and not staged and kh * kw == 9) else None
            if mk is not None and (tgt.grad is None or tgt.aux.get('du_ready')):
                sg = tgt.aux.get('segs') or [g['N']]
                if tgt.aux.get('du_q24') is None:
                    tgt.aux['du_q24'] = _ops().arena_take(2 * Cg * len(sg))
                tgt.aux['du_ready'] = True
                dx._tgan_du = True
                _igemm(dzb, g['N'], g['Ho'], g['Wo'], Cout, lddz, wp, Kpad, taps, Cg, g['H'], g['W'], dx, g['H'], g['W'], Cg,
                       colsum=tgt.aux['du_q24'], segs=sg, fused=dict(mask_in=mk, mask_alpha=tgt.aux.get('mask_alpha', 0.2)))
            else:
                _igemm(dzb, gf['N'], gf['Ho'], gf['Wo'], Cout, lddz, wp, Kpad, taps, Cg, gf['H'], gf['W'], dx, gf['H'],
                       gf['W'], Cs if staged else Cg)
        else:
            # input-gradient of a stride-2 conv = transposed conv: its output-parity classes, heaviest first, in ONE launch
            cl = []
            for py in range(2):
                for px in range(2):
                    rs = [r for r in range(kh) if (py + pt - r) % 2 == 0]
                    cs = [c for c in range(kw) if (px + pl - c) % 2 == 0]
                    if rs and cs:
                        cl.append((py, px, [r * kw + c for r in rs for c in cs],
                                   [((py + pt - r) // 2, (px + pl - c) // 2) for r in rs for c in cs]))
            assert len(cl) == 4, 'every output parity must receive a tap (else it would need an explicit zero fill)'
            cl.sort(key=lambda c: -len(c[2]))
            sel = [i for c in cl for i in c[2]]
            taps = [t for c in cl for t in c[3]]
            wp, Kpad = _pack(w, ('dgrad2', Cg), len(sel), Cg, Cout, C * Cout, Cout, 1, sel)
            _igemm(dzb, g['N'], g['Ho'], g['Wo'], Cout, lddz, wp, Kpad, taps, Cg, g['H'] // 2, g['W'] // 2, dx,
                   g['H'], g['W'], Cs if staged else Cg, os_=2, classes=[(len(c[2]), c[0], c[1]) for c in cl])
        if staged:
            _lib.call('tgan_copy_channels', dx.data_ptr(), BF16, Cs, dxt.data_ptr(), dt_code(dxt), Cg, x.rows, Cg, _st())
        add_grad(tgt, dxt if dxt.dtype == tgt.data.dtype else dxt.to(tgt.data.dtype))
    if par:      # the side stream still reads these; the join waits until the backward pass ends (ops.defer_join)
        _ops().defer_join([dzb, dz, g.get('_x'), g.get('_col'), x.data])
    g.pop('_x', None)
    g.pop('_col', None)


# ------------------------------------------------------------------------------------------------
# transposed convolution (tf.nn.conv2d_transpose 'SAME', stride 2)
# ------------------------------------------------------------------------------------------------


def deconv_eligible(g, x):
    return (_skinny(g) or (g['Cout'] >= 16 and g['Cout'] % 8 == 0)) and g['s'] == 2 and g['kh'] * g['kw'] <= 25


def _skinny(g):
    """the generator's last layer: 3 output channels (Good_GAN_cifar10.py:56).  Forward: the same parity-class launch
    into an 8-channel bf16 staging buffer, then a slice/convert to the fp32 image.  Backward: K per tap would be
    3 channels, so dy is gathered into a bf16 im2col matrix once and both gradients are plain GEMMs."""
    return g['Cout'] <= 8 and g['kh'] * g['kw'] * g['Cout'] <= 512


def _parity_classes(g):
    kh, kw, pt, pl = g['kh'], g['kw'], g['pt'], g['pl']
    for py in range(2):
        for px in range(2):
            rs = [r for r in range(kh) if (py + pt - r) % 2 == 0]
            cs = [c for c in range(kw) if (px + pl - c) % 2 == 0]
            if rs and cs:
                yield py, px, [r * kw + c for r in rs for c in cs], \
                    [((py + pt - r) // 2, (px + pl - c) // 2) for r in rs for c in cs]


def deconv_fwd(x, w, g, bias=None, act=0, ldo=None):
    Cin, Cout = g['Cin'], g['Cout']
    xd, ld = _bf16_padded(x.data, x.rows, Cin, x.ld)
    g['_x'] = (xd, ld)
    skinny = _skinny(g)
    if (skinny and g['h'] == 16 and g['w'] == 16 and g['kh'] == 5 and g['kw'] == 5 and g['s'] == 2 and Cin <= 144 and act in (0, 3)
            and getattr(w, 'scale', 1) is None and not os.environ.get('TGAN_NO_SKINNY_DECONV')):
        # the generator's RGB layer: pixels on the MMA rows, one CTA per image (csrc/deconv_skinny.cu)
        yf = _new((g['N'], g['Ho'], g['Wo'], Cout), torch.float32)
        _lib.call('tgan_deconv5s2_skinny', xd.data_ptr(), g['N'], ld, Cin, w.value().data_ptr(),
                  None if bias is None else bias.data_ptr(), Cout, yf.data_ptr(), act, _st())
        return yf
    ldo = 8 if skinny else (ldo or Cout)
    y = _new((g['N'], g['Ho'], g['Wo'], ldo), torch.bfloat16)
    cl = sorted(_parity_classes(g), key=lambda c: -len(c[2]))      # heaviest class first, all in ONE launch
    assert len(cl) == 4
    sel = [i for c in cl for i in c[2]]
    taps = [t for c in cl for t in c[3]]
    wp, Kpad = _pack(w, 'dfwd', len(sel), Cout, Cin, Cout * Cin, Cin, 1, sel)
    _igemm(xd, g['N'], g['h'], g['w'], Cin, ld, wp, Kpad, taps, Cout, g['h'], g['w'], y, g['Ho'], g['Wo'], ldo,
           os_=2, bias=bias, act=act, classes=[(len(c[2]), c[0], c[1]) for c in cl])
    if skinny:
        rows = g['N'] * g['Ho'] * g['Wo']
        yf = _new((g['N'], g['Ho'], g['Wo'], Cout), torch.float32)
        _lib.call('tgan_copy_channels', y.data_ptr(), BF16, ldo, yf.data_ptr(), 0, Cout, rows, Cout, _st())
        return yf
    return y


def _deconv_bwd_skinny(x, w, g, dy):
    from .core import add_grad
    Cin, Cout, kh, kw, pt, pl = g['Cin'], g['Cout'], g['kh'], g['kw'], g['pt'], g['pl']
    rows_out, rows_in = g['N'] * g['Ho'] * g['Wo'], g['N'] * g['h'] * g['w']
    dyb, ldy = _bf16_padded(dy, rows_out, Cout, Cout)
    K = kh * kw * Cout
    Kc = (K + 7) // 8 * 8
    col = _new((rows_in, Kc), torch.bfloat16)      # col[(n,i,j), (r,c,co)] = dy[n, 2i + r - pt, 2j + c - pl, co]
    _lib.call('tgan_im2col_bf16', dyb.data_ptr(), BF16, g['N'], g['Ho'], g['Wo'], Cout, ldy, kh, kw, 2, 2, pt, pl,
              g['h'], g['w'], col.data_ptr(), Kc, _st())
    if w.requires_grad:
        xd, ld = g.get('_x') or _bf16_padded(x.data, x.rows, Cin, x.ld)
        # dW[(r,c,co), ci] = col^T x: M side = x channels, N side = the im2col columns -> [K][Cin] == [kh,kw,Cout,Cin]
        _wgrad(xd, 1, 1, rows_in, Cin, ld, col, 1, rows_in, Kc, Kc, [(0, 0)], 1, w.grad_target(), cin_store=K)
    tgt, Cg = _grad_target(x, Cin)
    if tgt.requires_grad:
        # [ci][(r,c,co)]: k has stride Cin; a weight-normalised filter's per-channel scale is scale[k % Cout]
        wp, Kpad = _pack(w, ('ddgrad_col', Cg), 1, Cg, K, 0, 1, Cin, scale_mod=Cout)
        dx = _new(tuple(x.shape[:-1]) + (Cg,), torch.bfloat16)
        _igemm(col, 1, 1, rows_in, K, Kc, wp, Kpad, [(0, 0)], Cg, 1, rows_in, dx, 1, rows_in, Cg)
        add_grad(tgt, dx if dx.dtype == tgt.data.dtype else dx.to(tgt.data.dtype))
    g.pop('_x', None)


def deconv_bwd(x, w, g, dy):
    from .core import add_grad
    if _skinny(g):
        return _deconv_bwd_skinny(x, w, g, dy)
    Cin, Cout, kh, kw, pt, pl = g['Cin'], g['Cout'], g['kh'], g['kw'], g['pt'], g['pl']
    dyb, _ = _bf16_padded(dy, g['N'] * g['Ho'] * g['Wo'], Cout, Cout)
    taps = [(r - pt, c - pl) for r in range(kh) for c in range(kw)]
    tgt, Cg = _grad_target(x, Cin)
    par = w.requires_grad and tgt.requires_grad and _small(g['N'] * g['Ho'] * g['Wo'], max(Cin, Cout))
    if w.requires_grad:
        dw = w.grad_target()
        with (_ops().side_stream() if par else contextlib.nullcontext()):
            xd, ld = g.get('_x') or _bf16_padded(x.data, x.rows, Cin, x.ld)
            # roles exchanged: M side = x channels (ci), N side = strided dy window (co); the kernel's
            # [t][N-side][M-side] output is then exactly the filter layout [kh,kw,Cout,Cin]
            _wgrad(xd, g['N'], g['h'], g['w'], Cin, ld, dyb, g['Ho'], g['Wo'], Cout, Cout, taps, 2, dw)
    if tgt.requires_grad:
        wp, Kpad = _pack(w, ('ddgrad', Cg), kh * kw, Cg, Cout, Cout * Cin, 1, Cin)
        dx = _new(tuple(x.shape[:-1]) + (Cg,), torch.bfloat16)
        # few output tiles (4x4 / 8x8 images) x many taps (25): one CTA per tile would stream the whole 1.6 MB filter
        # through a single SM.  Split K: the tap list is cut into up to 4 groups that run as OUTPUT CLASSES of the same
        # launch, each writing its own bf16 partial slice (class offset = slice index * N * h rows), then one fold.
        tiles = -(-(g['N'] * g['h'] * g['w']) // 256) * -(-Cg // 128)
        parts = max(1, min(4, 148 // max(tiles, 1), len(taps) // 4))
        if parts > 1 and Cg % 8 == 0 and (g['N'] * g['h'] * g['w'] * Cg) % 8 == 0:
            per = -(-len(taps) // parts)
            cl = [(min(per, len(taps) - i * per), i * g['N'] * g['h'], 0) for i in range(parts)]
            part = _new((parts,) + tuple(x.shape[:-1]) + (Cg,), torch.bfloat16)
            _igemm(dyb, g['N'], g['Ho'], g['Wo'], Cout, Cout, wp, Kpad, taps, Cg, g['h'], g['w'], part, g['h'], g['w'], Cg,
                   s=2, classes=cl)
            _lib.call('tgan_sum_slices_bf16', part.data_ptr(), dx.data_ptr(), dx.numel(), parts, _st())
        else:
            _igemm(dyb, g['N'], g['Ho'], g['Wo'], Cout, Cout, wp, Kpad, taps, Cg, g['h'], g['w'], dx, g['h'], g['w'], Cg,
                   s=2)
        add_grad(tgt, dx if dx.dtype == tgt.data.dtype else dx.to(tgt.data.dtype))
    if par:
        _ops().defer_join([dyb, dy, g.get('_x'), x.data])
    g.pop('_x', None)
