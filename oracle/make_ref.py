"""Recipe for oracle/_ref/: copies the reference's four GShiftNet arch files, unmodified, from /root/reference (read-only,
present in the build container only) so that bench.py can time the reference's OWN implementation on the GPU box's host cores
(`--impl reference`, kind "reference") and as PyTorch-eager fp16 on the same B200 (`gpu_eager_baseline`).

TEST / MEASUREMENT INFRASTRUCTURE ONLY: nothing under shift-net_b200/, basicsr/ or inference/ reads oracle/_ref.  The directory is
git-ignored (reference sources never enter this repository's history) but not gpurun-ignored, so the copies travel with the
snapshot.  The files import only torch / numpy (gshift_deblur2.py:1-7), so they load by path without the basicsr package."""
import os
import shutil
import sys

SRC = "/root/reference/basicsr/models/archs"
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
ARCHS = ["gshift_deblur1", "gshift_deblur2", "gshift_denoise1", "gshift_denoise2"]


def main():
    if not os.path.isdir(SRC):
        have = [a for a in ARCHS if os.path.exists(os.path.join(DST, a + ".py"))]
        print(f"oracle/make_ref.py: {SRC} not present; keeping {len(have)} prebuilt file(s) in oracle/_ref")
        return 0
    os.makedirs(DST, exist_ok=True)
    for a in ARCHS:
        shutil.copyfile(os.path.join(SRC, a + ".py"), os.path.join(DST, a + ".py"))
    print(f"oracle/make_ref.py: copied {len(ARCHS)} arch files to {DST}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
