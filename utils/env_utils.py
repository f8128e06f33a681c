"""Alias of deepcubea_b200.utils.env_utils (reference import path)."""
from deepcubea_b200.utils.env_utils import *  # noqa: F401,F403
