"""B200-native engine for irrotational_warp's energy-condition evaluators.

Import as ``irrotational_warp_lab_b200`` (see the shim package of that name).
"""
__version__ = "0.1.0"
