"""Alias of deepcubea_b200.utils.data_utils (reference import path)."""
from deepcubea_b200.utils.data_utils import *  # noqa: F401,F403
