"""In-tree build of the CUDA library (nvcc, sm_100a only). `python -m vif_b200.build`."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "lib", "libvif_b200_qxmatch.so")


def build(verbose=False, force=False):
    csrc = os.path.join(HERE, "csrc")
    cmd = ["make", "-C", csrc, "-j", "6"]
    if force:
        subprocess.check_call(["make", "-C", csrc, "clean"], stdout=subprocess.DEVNULL)
    out = None if verbose else subprocess.DEVNULL
    subprocess.check_call(cmd, stdout=out)
    if not os.path.exists(LIB):
        raise RuntimeError("build finished but %s is missing" % LIB)
    return LIB


if __name__ == "__main__":
    print(build(verbose=True, force="--force" in sys.argv))
