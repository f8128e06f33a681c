"""Alias of deepcubea_b200.utils.pytorch_models (reference import path)."""
from deepcubea_b200.utils.pytorch_models import *  # noqa: F401,F403
