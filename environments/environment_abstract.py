"""Alias of deepcubea_b200.environments.environment_abstract (reference import path, used by pickles)."""
from deepcubea_b200.environments.environment_abstract import *  # noqa: F401,F403
from deepcubea_b200.environments.environment_abstract import Environment, State  # noqa: F401
