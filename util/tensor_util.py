"""``from util.tensor_util import pad_divide_by, unpad`` (inference_core.py:15, interact/interaction.py:15),
``compute_tensor_iou`` (davis_processor.py:9), ``unpad_3dim`` (interactive_gui.py:35)."""
from mivos_b200.tensor_util import (compute_multi_class_iou, compute_multi_class_iou_both_idx,  # noqa: F401
                                    compute_multi_class_iou_idx, compute_np_iou, compute_np_iu, compute_tensor_iou,
                                    compute_tensor_iu, pad_amounts, pad_divide_by, unpad, unpad_3dim)
