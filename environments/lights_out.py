"""Alias of deepcubea_b200.environments.lights_out (reference import path, used by pickles)."""
from deepcubea_b200.environments.lights_out import *  # noqa: F401,F403
from deepcubea_b200.environments.lights_out import LightsOut, LOState  # noqa: F401
