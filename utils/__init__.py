"""Reference-compatible import path: `utils.*` resolves to deepcubea_b200.utils.*"""
