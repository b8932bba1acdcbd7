"""``from model.aggregate import aggregate_wbg, aggregate_sbg`` (davis_processor.py:6,
inference_core.py:13, generation/fusion_generator.py:8); ``aggregate_wbg_channel`` for
model/fusion_model.py:10 (training, torch-only)."""
from mivos_b200.aggregate import aggregate_sbg, aggregate_wbg, aggregate_wbg_channel  # noqa: F401
