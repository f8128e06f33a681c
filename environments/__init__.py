"""Reference-compatible import path: `environments.*` resolves to deepcubea_b200.environments.*"""
