"""Re-hosted pipeline drivers with the reference's names (``inference/model_init.py``,
``inference/inference_utils.py``): ``from genvc_b200.inference.model_init import model_init``."""
