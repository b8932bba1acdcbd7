"""``from model.fusion_net import FusionNet`` (eval_interactive_davis.py:12, interactive_gui.py:32)."""
from mivos_b200.fusion_net import FusionNet  # noqa: F401
