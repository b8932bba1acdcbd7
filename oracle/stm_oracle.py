"""CPU oracle for the mask-propagation hot path of hkchengrex/MiVOS.

TEST INFRASTRUCTURE ONLY.  Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import this module; the product package
``mivos_b200`` never does (its ops raise when the CUDA library is missing).

It is a functional restatement, in plain PyTorch-CPU fp32 ops driven by a reference-format
``state_dict``, of the algorithm the reference implements with nn.Modules.  Every function cites
the reference lines it follows (paths relative to the reference root).  The arithmetic itself
lives in un-vendored PyTorch (README pins torch 1.7.1; this image has 2.11) — see DESIGN.md.

Pinning: the reference ships no tests or golden vectors, so this oracle is pinned
against the reference ITSELF, imported from /root/reference in the build container by
``oracle/gen_golden.py`` (seeded weights from ``oracle/weights.py``); the resulting fixtures live
in ``tests/golden/`` and ``tests/test_oracle_golden.py`` replays them here and on the GPU box.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]
BN_EPS = 1e-5  # nn.BatchNorm2d default, used by mod_resnet.py:80-87 and torchvision


# --------------------------------------------------------------------------------- util/tensor_util.py
def pad_divide_by(x: torch.Tensor, d: int, in_size: Optional[Sequence[int]] = None):
    """util/tensor_util.py:62-80 — symmetric zero pad of the last two dims to multiples of d."""
    h, w = x.shape[-2:] if in_size is None else in_size
    nh = h if h % d == 0 else h + d - h % d
    nw = w if w % d == 0 else w + d - w % d
    lh, lw = (nh - h) // 2, (nw - w) // 2
    pad = (lw, nw - w - lw, lh, nh - h - lh)
    return F.pad(x, pad), pad


def unpad(x: torch.Tensor, pad) -> torch.Tensor:
    """util/tensor_util.py:82-87."""
    if pad[2] + pad[3] > 0:
        x = x[:, :, pad[2]:x.shape[2] - pad[3], :]
    if pad[0] + pad[1] > 0:
        x = x[:, :, :, pad[0]:x.shape[3] - pad[1]]
    return x


# --------------------------------------------------------------------------------- model/aggregate.py
def aggregate_wbg(prob: torch.Tensor, keep_bg: bool = False, hard: bool = False) -> torch.Tensor:
    """model/aggregate.py:22-37 — bg = prod(1-p); clamp; logit; softmax over objects+bg."""
    bg = torch.prod(1 - prob, dim=0, keepdim=True)
    p = torch.cat([bg, prob], 0).clamp(1e-7, 1 - 1e-7)
    logits = torch.log(p / (1 - p))
    if hard:
        logits = logits * 1000
    sm = F.softmax(logits, dim=0)
    return sm if keep_bg else sm[1:]


# --------------------------------------------------------------------------------- conv building blocks
def _conv(sd: SD, name: str, x: torch.Tensor, stride: int = 1, padding: int = 0) -> torch.Tensor:
    return F.conv2d(x, sd[name + ".weight"], sd.get(name + ".bias"), stride=stride, padding=padding)


def _bn(sd: SD, name: str, x: torch.Tensor) -> torch.Tensor:
    """eval-mode BatchNorm2d (callers put the nets in .eval(): eval_interactive_davis.py:60,67)."""
    return F.batch_norm(x, sd[name + ".running_mean"], sd[name + ".running_var"], sd[name + ".weight"],
                        sd[name + ".bias"], False, 0.0, BN_EPS)


def _bottleneck(sd: SD, p: str, x: torch.Tensor, stride: int) -> torch.Tensor:
    """mod_resnet.py:76-112 (and torchvision's Bottleneck, stride on the 3x3): 1x1-bn-relu,
    3x3(stride)-bn-relu, 1x1-bn, + (downsampled) input, relu."""
    out = F.relu(_bn(sd, p + ".bn1", _conv(sd, p + ".conv1", x)))
    out = F.relu(_bn(sd, p + ".bn2", _conv(sd, p + ".conv2", out, stride=stride, padding=1)))
    out = _bn(sd, p + ".bn3", _conv(sd, p + ".conv3", out))
    if (p + ".downsample.0.weight") in sd:
        x = _bn(sd, p + ".downsample.1", _conv(sd, p + ".downsample.0", x, stride=stride))
    return F.relu(out + x)


def _layer(sd: SD, p: str, x: torch.Tensor, blocks: int, stride: int) -> torch.Tensor:
    """mod_resnet.py:133-148 — first block carries the stride and the downsample branch."""
    for i in range(blocks):
        x = _bottleneck(sd, f"{p}.{i}", x, stride if i == 0 else 1)
    return x


def _stem(sd: SD, p: str, x: torch.Tensor) -> torch.Tensor:
    """modules.py:56-59 / 80-83: 7x7/2 conv, bn, relu, 3x3/2 maxpool."""
    x = F.relu(_bn(sd, p + ".bn1", _conv(sd, p + ".conv1", x, stride=2, padding=3)))
    return F.max_pool2d(x, 3, 2, 1)


def rgb_encoder(sd: SD, frame: torch.Tensor):
    """RGBEncoder.forward, modules.py:79-89 -> (f16, f8, f4)."""
    x = _stem(sd, "rgb_encoder", frame)
    f4 = _layer(sd, "rgb_encoder.res2", x, 3, 1)
    f8 = _layer(sd, "rgb_encoder.layer2", f4, 4, 2)
    f16 = _layer(sd, "rgb_encoder.layer3", f8, 6, 2)
    return f16, f8, f4


def mask_rgb_encoder(sd: SD, frame: torch.Tensor, masks: torch.Tensor, others: torch.Tensor) -> torch.Tensor:
    """MaskRGBEncoder.forward, modules.py:52-64 (5-channel input, biased convs)."""
    x = _stem(sd, "mask_rgb_encoder", torch.cat([frame, masks, others], 1))
    x = _layer(sd, "mask_rgb_encoder.layer1", x, 3, 1)
    x = _layer(sd, "mask_rgb_encoder.layer2", x, 4, 2)
    return _layer(sd, "mask_rgb_encoder.layer3", x, 6, 2)


def key_value(sd: SD, p: str, x: torch.Tensor):
    """KeyValue.forward, modules.py:113-114."""
    return _conv(sd, p + ".key_proj", x, padding=1), _conv(sd, p + ".val_proj", x, padding=1)


def _resblock(sd: SD, p: str, x: torch.Tensor) -> torch.Tensor:
    """ResBlock.forward, modules.py:28-35 (pre-activation, optional 3x3 downsample on the skip)."""
    r = _conv(sd, p + ".conv1", F.relu(x), padding=1)
    r = _conv(sd, p + ".conv2", F.relu(r), padding=1)
    if (p + ".downsample.weight") in sd:
        x = _conv(sd, p + ".downsample", x, padding=1)
    return x + r


def _upsample_block(sd: SD, p: str, skip_f: torch.Tensor, up_f: torch.Tensor) -> torch.Tensor:
    """UpsampleBlock.forward, modules.py:100-104."""
    x = _resblock(sd, p + ".skip_conv2", _conv(sd, p + ".skip_conv1", skip_f, padding=1))
    x = x + F.interpolate(up_f, scale_factor=2, mode="bilinear", align_corners=False)
    return _resblock(sd, p + ".out_conv", x)


def decoder(sd: SD, f16: torch.Tensor, f8: torch.Tensor, f4: torch.Tensor) -> torch.Tensor:
    """Decoder.forward, prop_net.py:23-31 -> logits at full resolution."""
    x = _resblock(sd, "decoder.compress", f16)
    x = _upsample_block(sd, "decoder.up_16_8", f8, x)
    x = _upsample_block(sd, "decoder.up_8_4", f4, x)
    x = _conv(sd, "decoder.pred", F.relu(x), padding=1)
    return F.interpolate(x, scale_factor=4, mode="bilinear", align_corners=False)


# --------------------------------------------------------------------------------- memory read
def memory_read(mk: torch.Tensor, mv: torch.Tensor, qk: torch.Tensor, top_k: Optional[int]) -> torch.Tensor:
    """EvalMemoryReader.forward + softmax_w_g_top, prop_net.py:47-73,81-108 (km=None path).
    mk [B,CK,T,H,W], mv [B,CV,T,H,W], qk [1,CK,H,W] -> [B,CV,H,W]."""
    B, CK, T, H, W = mk.shape
    CV = mv.shape[1]
    mi = mk.reshape(B, CK, T * H * W).transpose(1, 2)
    qi = qk.reshape(1, CK, H * W).expand(B, -1, -1) / math.sqrt(CK)
    aff = torch.bmm(mi, qi)  # B, THW, HW
    if top_k is not None:
        vals, idx = torch.topk(aff, k=top_k, dim=1)
        e = torch.exp(vals - vals[:, 0])
        e = e / torch.sum(e, dim=1, keepdim=True)
        aff = torch.zeros_like(aff).scatter_(1, idx, e.to(aff.dtype))  # x_exp.type(x.dtype), prop_net.py:61
    else:
        aff = F.softmax(aff, dim=1)
    return torch.bmm(mv.reshape(B, CV, T * H * W), aff).view(B, CV, H, W)


def memory_read_f64(mk: np.ndarray, mv: np.ndarray, qk: np.ndarray, top_k: int):
    """Independent float64 numpy statement of the same read, used to arbitrate near-ties
    (SURVEY.md §8c-iii).  Inputs are the fp32 arrays; q is divided by sqrt(CK) in fp32 first,
    exactly like prop_net.py:86.  Returns (readout [B,CV,HW], idx [B,k,HW], sorted scores
    [B,THW-sorted-desc top k+1,...]) so callers can measure the k/(k+1) gap."""
    B, CK = mk.shape[:2]
    CV = mv.shape[1]
    THW = int(np.prod(mk.shape[2:]))
    HW = int(np.prod(qk.shape[2:]))
    qs = (qk.reshape(CK, HW).astype(np.float32) / np.float32(math.sqrt(CK))).astype(np.float64)
    out = np.zeros((B, CV, HW))
    idx_out = np.zeros((B, top_k, HW), dtype=np.int64)
    gap = np.zeros((B, HW))
    for b in range(B):
        aff = mk[b].reshape(CK, THW).astype(np.float64).T @ qs  # THW, HW
        order = np.argsort(-aff, axis=0, kind="stable")[: top_k + 1]
        top = np.take_along_axis(aff, order, axis=0)
        gap[b] = top[top_k - 1] - top[top_k] if THW > top_k else np.inf
        idx = order[:top_k]
        v = top[:top_k]
        e = np.exp(v - v[0:1])
        wgt = e / e.sum(0, keepdims=True)
        vals = mv[b].reshape(CV, THW).astype(np.float64)
        for q in range(HW):
            out[b, :, q] = vals[:, idx[:, q]] @ wgt[:, q]
        idx_out[b] = idx
    return out, idx_out, gap


def aggregate_wbg_f64(prob: np.ndarray, keep_bg: bool = False, hard: bool = False) -> np.ndarray:
    """Independent float64 numpy statement of aggregate_wbg (model/aggregate.py:22-37):
    prob [K,1,H,W] fp32 -> [(K+1)|K,1,H,W].  The clamp bounds are the fp32 values the reference uses."""
    p = prob.astype(np.float64)
    bg = np.prod(1.0 - p, axis=0, keepdims=True)
    q = np.clip(np.concatenate([bg, p], 0), np.float64(np.float32(1e-7)), np.float64(np.float32(1 - 1e-7)))
    lg = np.log(q / (1.0 - q)) * (1000.0 if hard else 1.0)
    lg = lg - lg.max(axis=0, keepdims=True)
    e = np.exp(lg)
    sm = e / e.sum(axis=0, keepdims=True)
    return sm if keep_bg else sm[1:]


def get_attention_f64(mk16: np.ndarray, pos_mask: np.ndarray, neg_mask: np.ndarray, qk16: np.ndarray) -> np.ndarray:
    """Independent float64 numpy statement of get_attention (prop_net.py:115-129,187-200; SURVEY.md
    §8c-iii) for one object: mk16 [1,128,1|,h,w], masks [1,1,H,W], qk16 [1,128,h,w] -> [1,2,H,W].
    Area pooling by 16 is a block mean; the final resize is bilinear with align_corners=False."""
    H, W = pos_mask.shape[-2:]
    h, w = H // 16, W // 16
    hw = h * w
    m = mk16.reshape(128, hw).astype(np.float64)                                  # [C, memory pixel]
    q = (qk16.reshape(128, hw).astype(np.float32) / np.float32(math.sqrt(128))).astype(np.float64)
    aff = m.T @ q                                                                 # [memory, query]
    aff = aff - aff.max(axis=0, keepdims=True)
    Wm = np.exp(aff)
    Wm = Wm / Wm.sum(axis=0, keepdims=True)                                       # softmax over the memory axis
    rows = []
    for mask in (pos_mask, neg_mask):
        pooled = mask.reshape(h, 16, w, 16).astype(np.float64).mean(axis=(1, 3)).reshape(1, hw)
        rows.append(pooled @ Wm)
    am = np.concatenate(rows, 0).reshape(2, h, w)

    def axis_taps(n_out, n_in):
        src = np.maximum((np.arange(n_out) + 0.5) * (n_in / n_out) - 0.5, 0.0)
        i0 = np.floor(src).astype(np.int64)
        i1 = np.minimum(i0 + 1, n_in - 1)
        return i0, i1, src - i0

    y0, y1, ly = axis_taps(H, h)
    x0, x1, lx = axis_taps(W, w)
    top = am[:, y0][:, :, x0] * (1 - lx) + am[:, y0][:, :, x1] * lx
    bot = am[:, y1][:, :, x0] * (1 - lx) + am[:, y1][:, :, x1] * lx
    return (top * (1 - ly)[None, :, None] + bot * ly[None, :, None])[None]


# --------------------------------------------------------------------------------- PropagationNetwork
def memorize(sd: SD, frame: torch.Tensor, masks: torch.Tensor):
    """PropagationNetwork.memorize, prop_net.py:144-162."""
    k, _, h, w = masks.shape
    fr = frame.view(1, 3, h, w).repeat(k, 1, 1, 1)
    if k != 1:
        others = torch.cat([torch.sum(masks[[j for j in range(k) if i != j]], dim=0, keepdim=True)
                            for i in range(k)], 0)
    else:
        others = torch.zeros_like(masks)
    f16 = mask_rgb_encoder(sd, fr, masks, others)
    k16, v16 = key_value(sd, "kv_m_f16", f16)
    return k16.unsqueeze(2), v16.unsqueeze(2)


def get_query_values(sd: SD, frame: torch.Tensor):
    """PropagationNetwork.get_query_values, prop_net.py:164-168."""
    f16, f8, f4 = rgb_encoder(sd, frame)
    k16, v16 = key_value(sd, "kv_q_f16", f16)
    return f16, f8, f4, k16, v16


def segment_with_query(sd: SD, keys, values, f16, f8, f4, k16, v16, top_k: int) -> torch.Tensor:
    """PropagationNetwork.segment_with_query, prop_net.py:170-181 (one object per read)."""
    k = keys.shape[0]
    m4 = torch.cat([memory_read(keys[i:i + 1], values[i:i + 1], k16, top_k) for i in range(k)], 0)
    m4 = torch.cat([m4, v16.expand(k, -1, -1, -1)], 1)
    return torch.sigmoid(decoder(sd, m4, f8, f4))


def get_attention(sd_unused, mk16: torch.Tensor, pos_mask: torch.Tensor, neg_mask: torch.Tensor,
                  qk16: torch.Tensor) -> torch.Tensor:
    """PropagationNetwork.get_attention + AttentionMemory.forward, prop_net.py:115-129,187-200."""
    b, _, h, w = pos_mask.shape
    nh, nw = h // 16, w // 16
    B, CK = mk16.shape[:2]
    m = mk16.reshape(B, CK, nh * nw).transpose(1, 2)
    q = qk16.reshape(1, CK, nh * nw).expand(B, -1, -1) / math.sqrt(CK)
    Wm = F.softmax(torch.bmm(m, q), dim=1)
    pm = F.interpolate(pos_mask, size=(nh, nw), mode="area").view(b, 1, nh * nw) @ Wm
    nm = F.interpolate(neg_mask, size=(nh, nw), mode="area").view(b, 1, nh * nw) @ Wm
    am = torch.cat([pm, nm], 1).reshape(b, 2, nh, nw)
    return F.interpolate(am, mode="bilinear", size=(h, w), align_corners=False)


def attention_read_network(sd: SD, image, mask11, mask21, mask12, mask22, query_image):
    """AttentionReadNetwork.forward, model/attn_network.py:48-80 (the batch-B, two-object twin of
    get_attention that FusionNet training calls, model/fusion_model.py:81-85): memory keys are
    encoded on the fly from (image, mask2x, the other object's mask2x), the query key from
    query_image, W = softmax over the memory axis (attn_network.py:17-28), area-pooled positive /
    negative mask differences are propagated through W and resized bilinearly."""
    b, _, h, w = mask11.shape
    nh, nw = h // 16, w // 16
    pos1, neg1 = (mask21 - mask11).clamp(0, 1), (mask11 - mask21).clamp(0, 1)
    pos2, neg2 = (mask22 - mask12).clamp(0, 1), (mask12 - mask22).clamp(0, 1)
    f16_1 = mask_rgb_encoder(sd, image, mask21, mask22)
    f16_2 = mask_rgb_encoder(sd, image, mask22, mask21)
    qf16, _, _ = rgb_encoder(sd, query_image)
    qk16, _ = key_value(sd, "kv_q_f16", qf16)

    def attn(f16, pos, neg):
        k16, _ = key_value(sd, "kv_m_f16", f16)
        m = k16.reshape(b, 128, nh * nw).transpose(1, 2)
        q = qk16.reshape(b, 128, nh * nw) / math.sqrt(128)
        Wm = F.softmax(torch.bmm(m, q), dim=1)
        pm = F.interpolate(pos, size=(nh, nw), mode="area").view(b, 1, nh * nw) @ Wm
        nm = F.interpolate(neg, size=(nh, nw), mode="area").view(b, 1, nh * nw) @ Wm
        am = torch.cat([pm, nm], 1).reshape(b, 2, nh, nw)
        return F.interpolate(am, mode="bilinear", size=(h, w), align_corners=False)

    return attn(f16_1, pos1, neg1), attn(f16_2, pos2, neg2)


# --------------------------------------------------------------------------------- FusionNet
def fusion_net(sd: SD, im, seg1, seg2, attn, time) -> torch.Tensor:
    """FusionNet.forward, model/fusion_net.py:32-50 -> logit."""
    h, w = im.shape[-2:]
    t = time.unsqueeze(2).unsqueeze(2).expand(-1, -1, h, w)
    x = torch.cat([im, seg1, seg2, attn, t], 1)
    x = F.relu(_conv(sd, "conv1.0", x, padding=1))
    r = _conv(sd, "conv2.2", F.relu(_conv(sd, "conv2.0", x, padding=1)), padding=1)
    x = F.relu(x + r)
    r = _conv(sd, "conv3.2", F.relu(_conv(sd, "conv3.0", x, padding=1)), padding=1)
    x = F.relu(x + r)
    return _conv(sd, "final_conv", x, padding=1)


# --------------------------------------------------------------------------------- InferenceCore
class OracleInferenceCore:
    """Restatement of InferenceCore (inference_core.py:35-292) with mem_profile 0 on one device.
    State layout and bank bookkeeping follow the reference exactly; see the line references."""

    def __init__(self, prop_sd: SD, fuse_sd: Optional[SD], images: torch.Tensor, num_objects: int,
                 mem_freq: int = 5, top_k: int = 50, device="cpu"):
        """`device`: where the reference's `device=` argument puts everything (inference_core.py:36,
        mem_profile 0).  'cpu' is the oracle proper; a CUDA device runs the SAME eager PyTorch ops
        through cuDNN / cuBLAS — the "reference kernels on the same box" comparator of bench.py."""
        self.device = torch.device(device)
        dev = self.device
        if dev.type != "cpu":
            prop_sd = {k: v.to(dev) for k, v in prop_sd.items()}
            fuse_sd = None if fuse_sd is None else {k: v.to(dev) for k, v in fuse_sd.items()}
            images = images.to(dev)
        self.sd, self.fsd = prop_sd, fuse_sd
        self.mem_freq, self.top_k = mem_freq, top_k
        self.t = images.shape[1]
        self.h, self.w = images.shape[-2:]
        self.k = num_objects
        self.images, self.pad = pad_divide_by(images, 16, images.shape[-2:])  # :71
        self.nh, self.nw = self.images.shape[-2:]
        self.masks = torch.zeros((self.t, 1, self.nh, self.nw), dtype=torch.uint8, device=dev)  # :77
        self.np_masks = np.zeros((self.t, self.h, self.w), dtype=np.uint8)
        self.prob = torch.zeros((self.k + 1, self.t, 1, self.nh, self.nw), dtype=torch.float32, device=dev)  # :81
        self.prob[0] = 1e-7  # :82
        self.query_buf: Dict[int, tuple] = {}
        self.interacted = set()
        self.certain_mem_k = None
        self.certain_mem_v = None
        self.bank_trace: List[Tuple[int, int]] = []  # (frame, slots visible) — for plumbing tests

    def _query(self, ti: int):
        if ti not in self.query_buf:  # :110-120 (q_buf_size 105 never exceeded in tests)
            self.query_buf[ti] = get_query_values(self.sd, self.images[:, ti])
        return self.query_buf[ti]

    def do_pass(self, key_k, key_v, idx: int, forward: bool = True, step_cb=None):
        """inference_core.py:122-200."""
        nck = self.certain_mem_k.shape[2]
        m_front = nck
        if forward:
            closest = min([ti for ti in self.interacted if ti > idx] + [self.t])
            total_m = (closest - idx - 1) // self.mem_freq + 1 + nck
        else:
            closest = max([ti for ti in self.interacted if ti < idx] + [-1])
            total_m = (idx - closest - 1) // self.mem_freq + 1 + nck
        K, CK, _, H, W = key_k.shape
        CV = key_v.shape[1]
        keys = torch.empty((K, CK, total_m, H, W), device=self.device)  # fp32 even under autocast (:146-147)
        values = torch.empty((K, CV, total_m, H, W), device=self.device)
        keys[:, :, :nck] = self.certain_mem_k
        values[:, :, :nck] = self.certain_mem_v
        prev_in_mem, last_ti = True, idx
        if forward:
            rng, end = range(idx + 1, closest), closest - 1
        else:
            rng, end = range(idx - 1, closest, -1), closest + 1
        for ti in rng:
            vis = m_front if prev_in_mem else m_front + 1  # :166-171
            self.bank_trace.append((ti, vis))
            q = self._query(ti)
            out = segment_with_query(self.sd, keys[:, :, :vis], values[:, :, :vis], *q, top_k=self.top_k)
            out = aggregate_wbg(out, keep_bg=True)
            if ti != end:  # :177-186
                k_, v_ = memorize(self.sd, self.images[:, ti], out[1:])
                keys[:, :, m_front:m_front + 1], values[:, :, m_front:m_front + 1] = k_, v_
                if abs(ti - last_ti) >= self.mem_freq:
                    m_front += 1
                    last_ti = ti
                    prev_in_mem = True
                else:
                    prev_in_mem = False
            if closest != self.t and closest != -1:  # :190-194
                self.prob[:, ti] = self.fuse_one_frame(closest, idx, ti, self.prob[:, ti], out, key_k, q[3])
            else:
                self.prob[:, ti] = out
            if step_cb is not None:
                step_cb()
        return closest

    def fuse_one_frame(self, tc, tr, ti, prev_mask, curr_mask, mk16, qk16):
        """inference_core.py:202-217."""
        assert tc < ti < tr or tr < ti < tc
        prob = torch.zeros((self.k, 1, self.nh, self.nw), device=self.device)
        nc = abs(tc - ti) / abs(tc - tr)
        nr = abs(tr - ti) / abs(tc - tr)
        dist = torch.tensor([nc, nr], dtype=torch.float32, device=self.device).unsqueeze(0)
        for k in range(1, self.k + 1):
            attn = get_attention(None, mk16[k - 1:k], self.pos_mask_diff[k:k + 1], self.neg_mask_diff[k:k + 1], qk16)
            prob[k - 1] = torch.sigmoid(fusion_net(self.fsd, self.images[:, ti], prev_mask[k:k + 1],
                                                   curr_mask[k:k + 1], attn, dist))
        return aggregate_wbg(prob, keep_bg=True)

    def interact(self, mask: torch.Tensor, idx: int, total_cb=None, step_cb=None) -> np.ndarray:
        """inference_core.py:219-271."""
        self.interacted.add(idx)
        mask, _ = pad_divide_by(mask.to(self.device), 16, mask.shape[-2:])
        self.mask_diff = mask - self.prob[:, idx]
        self.pos_mask_diff = self.mask_diff.clamp(0, 1)
        self.neg_mask_diff = (-self.mask_diff).clamp(0, 1)
        self.prob[:, idx] = mask
        key_k, key_v = memorize(self.sd, self.images[:, idx], mask[1:])
        if self.certain_mem_k is None:
            self.certain_mem_k, self.certain_mem_v = key_k, key_v
        else:
            self.certain_mem_k = torch.cat([self.certain_mem_k, key_k], 2)
            self.certain_mem_v = torch.cat([self.certain_mem_v, key_v], 2)
        if total_cb is not None:
            front = min([ti for ti in self.interacted if ti > idx] + [self.t])
            back = max([ti for ti in self.interacted if ti < idx] + [-1])
            if front - back - 2 > 0:
                total_cb(front - back - 2)
        self.do_pass(key_k, key_v, idx, True, step_cb)
        self.do_pass(key_k, key_v, idx, False, step_cb)
        for ti in range(self.t):
            self.masks[ti] = torch.argmax(self.prob[:, ti], dim=0)
        out = unpad(self.masks, self.pad)
        self.np_masks = out.cpu().numpy()[:, 0].astype(np.uint8)
        return self.np_masks
