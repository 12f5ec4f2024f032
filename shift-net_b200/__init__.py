"""shift-net_b200: B200-native (sm_100a) implementation of Shift-Net's GShiftNet forward hot path.

The directory name carries a hyphen, so import it with
``importlib.import_module("shift-net_b200")`` (or through the ``basicsr.models.archs.*`` shims,
which is what the reference's entry points do).
"""
