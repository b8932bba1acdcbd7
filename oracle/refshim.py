"""Import the UNMODIFIED reference from /root/reference (build container only — the GPU box has
no /root/reference, so nothing under tests -m gpu / smoke() / bench.py imports this module).

The only host-side patch replaces the ImageNet weight download that the reference performs while
constructing its encoders (model/propagation/modules.py:42,70 -> mod_resnet.py:153-157 and
torchvision.models.resnet50(pretrained=True)): there is no network here, and every weight is
overwritten by a strict ``load_state_dict`` of our seeded state dict right after construction.
"""
from __future__ import annotations

import contextlib
import os
import sys

REF_ROOT = os.environ.get("MIVOS_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "model", "propagation"))


_REF_PACKAGES = ("model", "util", "inference_core", "interact", "dataset")


def _is_ref_name(k: str) -> bool:
    return k in _REF_PACKAGES or any(k.startswith(p + ".") for p in _REF_PACKAGES)


@contextlib.contextmanager
def reference_on_path():
    """Temporarily put the reference root first on sys.path and hide same-named local shims
    (model/, util/, interact/, inference_core.py); on exit the reference's modules are moved aside
    (``_ref_<name>``) and the local ones restored, so later imports resolve to this repository again."""
    shadow = {k: sys.modules.pop(k) for k in list(sys.modules) if _is_ref_name(k)}
    sys.path.insert(0, REF_ROOT)
    try:
        yield
    finally:
        sys.path.remove(REF_ROOT)
        for k in list(sys.modules):
            if _is_ref_name(k):
                m = sys.modules[k]
                if getattr(m, "__file__", "") and str(m.__file__).startswith(REF_ROOT):
                    sys.modules["_ref_" + k] = sys.modules.pop(k)
        sys.modules.update(shadow)


def load_reference():
    """Returns a namespace with the reference's PropagationNetwork, FusionNet, InferenceCore,
    aggregate_wbg, pad_divide_by classes/functions."""
    import types

    import torch
    import torchvision

    if not available():
        raise RuntimeError(f"reference not found at {REF_ROOT}")
    with reference_on_path():
        import torch.utils.model_zoo as mz

        def _no_download(*a, **k):
            return torchvision.models.resnet50(weights=None).state_dict()

        orig_load, orig_r50 = mz.load_url, torchvision.models.resnet50
        mz.load_url = _no_download
        torchvision.models.resnet50 = lambda pretrained=False, **k: orig_r50(weights=None)
        try:
            import model.propagation.mod_resnet as mod_resnet

            mod_resnet.model_zoo.load_url = _no_download
            from model.propagation.prop_net import PropagationNetwork
            from model.fusion_net import FusionNet
            from model.attn_network import AttentionReadNetwork
            from model.aggregate import aggregate_wbg, aggregate_sbg
            from util.tensor_util import pad_divide_by, unpad
            from inference_core import InferenceCore

            def build_prop(state_dict, top_k):
                net = PropagationNetwork(top_k=top_k).eval()
                net.load_state_dict(state_dict, strict=True)
                return net

            def build_fusion(state_dict):
                net = FusionNet().eval()
                net.load_state_dict(state_dict, strict=True)
                return net

            def build_attn(state_dict):
                net = AttentionReadNetwork().eval()
                net.load_state_dict(state_dict, strict=False)  # fusion_model.py:187: prop checkpoint, decoder keys unused
                return net

            ns = types.SimpleNamespace(AttentionReadNetwork=AttentionReadNetwork, build_attn=build_attn,
                
                PropagationNetwork=PropagationNetwork, FusionNet=FusionNet, InferenceCore=InferenceCore,
                aggregate_wbg=aggregate_wbg, aggregate_sbg=aggregate_sbg, pad_divide_by=pad_divide_by, unpad=unpad,
                build_prop=build_prop, build_fusion=build_fusion)
            # building needs the patched constructors, so keep them patched while callers build;
            # callers run under torch.no_grad()
            ns._restore = lambda: (setattr(mz, "load_url", orig_load), setattr(torchvision.models, "resnet50", orig_r50))
            return ns
        except Exception:
            mz.load_url, torchvision.models.resnet50 = orig_load, orig_r50
            raise
