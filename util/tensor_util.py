"""``from util.tensor_util import pad_divide_by, unpad`` (davis_processor.py, inference_core.py:15)."""
from mivos_b200.tensor_util import pad_amounts, pad_divide_by, unpad, unpad_3dim  # noqa: F401
