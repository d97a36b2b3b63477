"""Stages the handful of reference files the checkers need under baseline/_ref/ -- TEST INFRASTRUCTURE ONLY.

/root/reference exists in the build container only; the GPU box receives a snapshot of /root/repo.  SURVEY.md section 0
/ 7 therefore plan copies of the hot path's reference files under baseline/_ref/ (git-ignored: they never enter the
history; NOT gpurun-ignored: they travel with the snapshot).  They are used by
  * bench.py --impl reference       the reference's own MonoSDFNetwork + MonoSDFLoss + torch.optim.Adam on the host cores
  * tests/test_gpu_dropin.py        utils.general.get_class, model.loss.MonoSDFLoss and a reference-made checkpoint
                                    against the drop-in module
  * oracle/ref_shim.py              falls back to this tree when /root/reference is absent
Nothing under monosdf_b200/ reads them.  Run by __graft_entry__.build(); a no-op where /root/reference is absent.
"""
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.environ.get("MSDF_REFERENCE_SOURCE", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")

FILES = [
    "code/model/network.py", "code/model/ray_sampler.py", "code/model/density.py", "code/model/embedder.py",
    "code/model/loss.py", "code/utils/general.py", "code/utils/rend_util.py", "code/utils/plots.py",
    "code/hashencoder/__init__.py", "code/hashencoder/hashgrid.py", "code/hashencoder/backend.py",
]


def stage():
    if not os.path.isfile(os.path.join(SRC, FILES[0])):
        return None
    for rel in FILES:
        dst = os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(SRC, rel), dst)
    return DST


if __name__ == "__main__":
    print(stage())
