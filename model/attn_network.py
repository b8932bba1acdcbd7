"""``from model.attn_network import AttentionReadNetwork`` (model/fusion_model.py:12)."""
from mivos_b200.attn_network import AttentionReadNetwork  # noqa: F401
