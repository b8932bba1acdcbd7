"""``from util.palette import pal_color_map`` (interactive_gui.py:36)."""
from mivos_b200.egress import davis_color_map as get_color_map  # noqa: F401

color_map = get_color_map()


def pal_color_map():
    return color_map
