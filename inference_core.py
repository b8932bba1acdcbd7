"""Drop-in for the reference's top-level ``inference_core`` module (inference_core.py:17):
``from inference_core import InferenceCore`` (davis_processor.py:10, interactive_gui.py:33,
eval_interactive_davis.py via DAVISProcessor) resolves to the B200-native engine."""
from mivos_b200.inference_core import InferenceCore  # noqa: F401
