"""B200-native PinSage inference + retrieval hot path behind the reference's Python API.

Sub-packages mirror the reference's module layout so call sites keep their imports:
``utils.random_walk``, ``utils.nearest_neighbors``, ``utils.evaluation``,
``model.pinsage``, ``model.layers``, ``model.aggregators``.
All compute goes through ``libpinsage_b200.so`` (include/pinsage_b200.h); there is no
CPU or PyTorch fallback: a missing library or device raises.
"""
__version__ = "0.1.0"
