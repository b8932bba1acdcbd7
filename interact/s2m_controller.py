"""``from interact.s2m_controller import S2MController`` (interactive_gui.py:30)."""
from mivos_b200.s2m import S2MController  # noqa: F401
