"""``from model.aggregate import aggregate_wbg, aggregate_sbg`` (davis_processor.py:6,
inference_core.py:13, generation/fusion_generator.py:8)."""
from mivos_b200.aggregate import aggregate_sbg, aggregate_wbg  # noqa: F401
