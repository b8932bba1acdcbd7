"""TEST INFRASTRUCTURE: a PyTorch-CPU stand-in for the handful of C-ABI operators the S2M layer graph
uses (include/mivos_b200.h), written from the header's contracts.  Monkeypatched over
``mivos_b200.ops`` by tests/test_s2m_cpu.py so that the HOST side of ``engine.S2MEngine`` — weight
packing, channel windows of the concat buffers, dilation tables, buffer shapes, gather orders —
can be checked against the reference-generated golden vectors without a GPU.  It says nothing
about the kernels themselves (tests/test_gpu_s2m.py does, on a B200) and nothing in the product
package imports it."""
from __future__ import annotations

import torch
import torch.nn.functional as F


def halo_zeros(n, h, w, c, device, dtype=torch.float32):
    return torch.zeros((n, h + 2, w + 2, c), dtype=torch.float32)


def split_k_workspace(device):
    return torch.zeros(16, dtype=torch.uint8)


def _interior(n, h, w):
    m = torch.zeros((n, h + 2, w + 2), dtype=torch.bool)
    m[:, 1:-1, 1:-1] = True
    return m.reshape(-1)


def conv_gemm(x, pc, n, h, w, out, *, in_coff=0, out_coff=0, relu=False, residual=None, res_coff=0, out_relu=None,
              out_relu_coff=0, round_tf32=False, splitk_ws=None):
    rows = n * (h + 2) * (w + 2)
    X = x.reshape(-1, x.shape[-1])[:, in_coff:in_coff + pc.cin_pad].float()
    assert X.shape[0] >= rows and X.shape[1] == pc.cin_pad, (X.shape, rows, pc.cin_pad)
    W = pc.weight.float()  # [taps, cout_pad, cin_pad]
    acc = torch.zeros((rows, pc.cout_pad))
    if pc.taps == 1:
        acc = X[:rows] @ W[0].t()
    else:
        assert pc.taps == 9 and X.shape[0] == rows
        for t in range(9):
            off = (t // 3 - 1) * (w + 2) + (t % 3 - 1)
            sh = torch.zeros_like(X)
            if off >= 0:
                sh[:rows - off] = X[off:]
            else:
                sh[-off:] = X[:rows + off]
            acc = acc + sh @ W[t].t()
    v = acc[:, :pc.cout] + pc.bias.float()[:pc.cout]
    inner = _interior(n, h, w)
    if residual is not None:
        v = v + residual.reshape(-1, residual.shape[-1])[:rows, res_coff:res_coff + pc.cout]
    if relu:
        v = v.clamp_min(0)
    o = out.reshape(-1, out.shape[-1])
    o[:rows][inner, out_coff:out_coff + pc.cout] = v[inner]
    if out_relu is not None:
        r = out_relu.reshape(-1, out_relu.shape[-1])
        r[:rows][inner, out_relu_coff:out_relu_coff + pc.cout] = v[inner].clamp_min(0)
    return out


def _nchw(halo, n, h, w, c, coff=0):
    return halo[:, 1:-1, 1:-1, coff:coff + c].permute(0, 3, 1, 2).contiguous()


def _windows_to_rows(cols, n, k2, c, ho, wo, out):
    """F.unfold output [n, c*k2, ho*wo] (channel-major, tap-minor) -> matrix rows = HALO rows of the
    (ho, wo) map, columns (tap, channel); halo rows zero."""
    cols = cols.reshape(n, c, k2, ho, wo).permute(0, 3, 4, 2, 1).reshape(n, ho, wo, k2 * c)
    m = out.reshape(n, ho + 2, wo + 2, out.shape[-1])
    m.zero_()
    m[:, 1:-1, 1:-1, :k2 * c] = cols
    return out


def stem_gather_frames(frames, out):
    n, cin, h, w = frames.shape
    return _windows_to_rows(F.unfold(frames, 7, padding=3, stride=2), n, 49, cin, h // 2, w // 2, out)


def gather_s2(x, n, h, w, c, ks, out):
    return _windows_to_rows(F.unfold(_nchw(x, n, h, w, c), ks, padding=ks // 2, stride=2), n, ks * ks, c, h // 2, w // 2, out)


def gather_dilated(x, n, h, w, c, dilation, out):
    return _windows_to_rows(F.unfold(_nchw(x, n, h, w, c), 3, padding=dilation, dilation=dilation), n, 9, c, h, w, out)


def maxpool3x3s2(x, n, h, w, out):
    c = x.shape[-1]
    out[:, 1:-1, 1:-1, :] = F.max_pool2d(_nchw(x, n, h, w, c), 3, 2, 1).permute(0, 2, 3, 1)
    return out


def halo_avgpool_broadcast(x, n, h, w, c, out, *, in_coff=0, out_coff=0):
    m = _nchw(x, n, h, w, c, in_coff).mean(dim=(2, 3))  # [n, c]
    out[:, 1:-1, 1:-1, out_coff:out_coff + c] = m[:, None, None, :]
    return out


def upsample_bilinear(src, n, hs, ws, dst, h, w, c, *, src_coff=0, dst_coff=0):
    up = F.interpolate(_nchw(src, n, hs, ws, c, src_coff), size=(h, w), mode="bilinear", align_corners=False)
    dst[:, 1:-1, 1:-1, dst_coff:dst_coff + c] = up.permute(0, 2, 3, 1)
    return dst


def halo_upsample_to_plane(halo, n, hs, ws, out_h, out_w, *, coff=0, sigmoid=False, out=None):
    up = F.interpolate(_nchw(halo, n, hs, ws, 1, coff), size=(out_h, out_w), mode="bilinear", align_corners=False)
    up = torch.sigmoid(up) if sigmoid else up
    if out is not None:
        out.copy_(up)
        return out
    return up


OPS = ("halo_zeros", "split_k_workspace", "conv_gemm", "stem_gather_frames", "gather_s2", "gather_dilated", "maxpool3x3s2",
       "halo_avgpool_broadcast", "upsample_bilinear", "halo_upsample_to_plane")


def install(monkeypatch, ops_module):
    import sys
    me = sys.modules[__name__]
    for name in OPS:
        monkeypatch.setattr(ops_module, name, getattr(me, name))
