"""Alternative builds of the CUDA library for kernel experiments: tools/bin/lib_<name>.so (git-ignored; they travel to the
GPU box with the snapshot) -- compared by tools/gpu_variants.sh through V2GT_LIB_CUDA.  usage: build_variants.py name=-DFLAG=1,-DOTHER=2 ..."""
import os
import shutil
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vicon2gt_b200 import build

for spec in sys.argv[1:]:
    name, _, flags = spec.partition("=")
    out = os.path.join(build.ROOT, "tools", "bin", f"lib_{name}.so")
    build.build_cuda(force=True, extra=[f for f in flags.split(",") if f], out=out, objdir_name="_obj_" + name)
    shutil.rmtree(os.path.join(build.ROOT, "vicon2gt_b200", "_obj_" + name), ignore_errors=True)  # the objects would travel with every gpurun snapshot
    print("built", out)
