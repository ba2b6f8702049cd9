"""Place an unmodified copy of the reference under baseline/_ref.

The prescribed ``pip install --no-index --no-build-isolation --find-links /opt/wheelhouse --target baseline/_ref
/root/reference`` fails ("Neither 'setup.py' nor 'pyproject.toml' found" -- the reference is a flat directory of
scripts), so the install is a plain copy of ``src/*.py`` and the two PNG assets.  Outcome is
recorded in DESIGN.md.  baseline/_ref is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
import os
import shutil
import sys


def install(src="/root/reference", dst=None):
    dst = dst or os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
    if not os.path.isdir(os.path.join(src, "src")):
        return False
    os.makedirs(dst, exist_ok=True)
    shutil.copytree(os.path.join(src, "src"), os.path.join(dst, "src"), dirs_exist_ok=True)
    for f in os.listdir(src):
        if f.endswith(".png"):
            shutil.copy(os.path.join(src, f), os.path.join(dst, f))
    return True


if __name__ == "__main__":
    print("installed" if install(*(sys.argv[1:2])) else "reference not found")
