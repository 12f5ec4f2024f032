"""Model registry of the hot path: ``create_video_model(opt)`` resolves ``opt['model']`` to an arch module and calls
its ``make_model`` (reference: basicsr/models/image_restoration_model.py:22-25)."""
import importlib


def create_video_model(opt):
    module = importlib.import_module("basicsr.models.archs." + opt["model"].lower())
    return module.make_model(opt)
