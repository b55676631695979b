"""How many pixels do the fast stack kernels leave to the generic routine, and why? (dev tool)
Runs with APGPU_SKIP_MARKED=1 so that the marks stay in the output image.
Usage: APGPU_SKIP_MARKED=1 python tools/count_marks.py [N ...]"""
import os
import sys

os.environ["APGPU_SKIP_MARKED"] = "1"
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench                                                     # noqa: E402
from astrophotography_b200 import kernels                        # noqa: E402
from tools.time_round2 import timeit, KAPPA, MEDMAD              # noqa: E402

WHY = {0: "unspecified", 1: "non-finite", 2: "bounds degenerate / pivot outside", 3: "guard band", 4: "all rejected"}


def main():
    ns = [int(x) for x in sys.argv[1:] if x.isdigit()] or [30, 100, 256, 512]
    dev = torch.device("cuda", 0)
    h, w = 2048, 2048
    for n in ns:
        cube = bench.synth_cube_device(torch, n, h, w, dev, seed=1000)
        for name, kw in (("kappa-sigma", KAPPA), ("medmad", MEDMAD)):
            out = {}
            ms = timeit(lambda: kernels.stack_reduce(cube, out=out, **kw))
            bits = out["data"].view(torch.int32)
            marked = (bits & ~15) == 0x7fc5a5a0
            cnt = int(marked.sum())
            reasons = torch.bincount((bits[marked] & 15).to(torch.int64), minlength=5).tolist()
            print(f"N={n} {name:12s} {kernels.stack_kernel_name(n, **kw):22s} {ms:7.3f} ms without the cleanup launch; "
                  f"marked {cnt} of {h * w} ({cnt / (h * w):.2e}): " +
                  ", ".join(f"{WHY[i]}={c}" for i, c in enumerate(reasons) if c), flush=True)
        del cube
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
