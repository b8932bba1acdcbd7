"""``from model.s2m.s2m_network import deeplabv3plus_resnet50 as S2M`` (davis_processor.py:8,
interactive_gui.py:34, interact/s2m_controller.py:3)."""
from mivos_b200.s2m import S2MNetwork, deeplabv3plus_resnet50  # noqa: F401
