"""Drop-in for the reference's basicsr/models/archs/gshift_denoise1.py: exports ``GShiftNet`` and ``make_model(opt)``
(reference: gshift_denoise1.py:9-16 and the GShiftNet class) backed by the B200 kernels of shift-net_b200/."""
import importlib

GShiftNet = importlib.import_module("shift-net_b200.host.gshift").make_arch("gshift_denoise1")


def make_model(opt):
    # the reference reads opt['pretrain_models_dir'] but ignores it and returns GShiftNet() with defaults
    _ = opt.get("pretrain_models_dir") if hasattr(opt, "get") else None
    return GShiftNet()
