"""Drop-in for the reference's inference/test_deblur_small.py (same CLI: --default_data {GOPRO,DVD} --one_len N --save_image), running
gshift_deblur2 on the B200 kernels.  Multi-GPU: launch with torchrun --nproc-per-node N (clips shard across ranks, one final
all_gather of the PSNR/SSIM records).  --synthetic V evaluates on V generated videos when no dataset is present."""
import argparse
import importlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from basicsr.models.archs.gshift_deblur2 import GShiftNet  # noqa: E402

infer = importlib.import_module("shift-net_b200.host.infer")

if __name__ == "__main__":
    parser = infer.add_common_args(argparse.ArgumentParser(description="Shift-Net deblur inference (shiftnet_b200)"))
    parser.add_argument("--one_len", type=int, default=96)
    args = parser.parse_args()
    defaults = {"DVD": ("./dataset/DVD/test/", "pretrained_models/net_dvd_deblur_small.pth", "infer_results/DVD"),
                "GOPRO": ("./dataset/GOPRO/test/", "pretrained_models/net_gopro_deblur_small.pth", "infer_results/gopro")}
    d = defaults.get(args.default_data, (".", None, "infer_results/custom"))
    args.data_path = args.data_path or d[0]
    args.model_path = args.model_path or d[1]
    args.result_path = args.result_path or d[2]
    infer.run(GShiftNet, "deblur", args)
