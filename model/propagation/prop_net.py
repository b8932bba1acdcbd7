"""``from model.propagation.prop_net import PropagationNetwork`` (eval_interactive_davis.py:11,
interactive_gui.py:29, generate_fusion.py:15) -> B200-native implementation."""
from mivos_b200.prop_net import PropagationNetwork  # noqa: F401
