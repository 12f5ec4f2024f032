"""Drop-in for the reference's inference/test_denoise_small.py (same CLI: --default_data {DAVIS,Set8} --sigma S --save_image), running
gshift_denoise2 on the B200 kernels: AWGN sigma/255, constant noise map, 2x2 overlapped spatial tiling.  Multi-GPU: torchrun
--nproc-per-node N (clips shard across ranks, one final all_gather).  --synthetic V for generated videos."""
import argparse
import importlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from basicsr.models.archs.gshift_denoise2 import GShiftNet  # noqa: E402

infer = importlib.import_module("shift-net_b200.host.infer")

if __name__ == "__main__":
    parser = infer.add_common_args(argparse.ArgumentParser(description="Shift-Net denoise inference (shiftnet_b200)"))
    parser.add_argument("--sigma", type=int, default=10)
    parser.add_argument("--one", type=int, default=10)
    args = parser.parse_args()
    defaults = {"DAVIS": ("./dataset/DAVIS-test", "pretrained_models/net_denoise_small.pth", "infer_results/DAVIS_2/sigma%d" % args.sigma),
                "Set8": ("./dataset/Set8", "pretrained_models/net_denoise_small.pth", "infer_results/Set8_2/sigma%d" % args.sigma)}
    d = defaults.get(args.default_data, (".", None, "infer_results/custom"))
    args.data_path = args.data_path or d[0]
    args.model_path = args.model_path or d[1]
    args.result_path = args.result_path or d[2]
    infer.run(GShiftNet, "denoise", args)
