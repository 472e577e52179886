"""Stage the UNMODIFIED reference under baseline/_ref/ (git-ignored, shipped to the GPU box by gpurun).

The reference (SHI-Labs/Prompt-Free-Diffusion) is an un-packaged Python app: no setup.py / pyproject, it runs
only from its own checkout with `lib/` and `configs/` resolved relative to the CWD (lib/cfg_helper.py:104), so
`pip install --target baseline/_ref /root/reference` has nothing to build.  The install step is
therefore a verbatim copy of the two directories the model_zoo needs (0.8 MB; assets/ and app.py are not needed:
app.py cannot be imported without gradio / pretrained weights).  Nothing under baseline/_ref is ever edited,
imported by the product, or committed.

    python tools/install_reference.py            # no-op when /root/reference is absent (GPU box)
"""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get("PFD_REFERENCE_SRC", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")


def install(verbose=True) -> bool:
    if not os.path.isdir(os.path.join(SRC, "lib", "model_zoo")):
        if verbose:
            print(f"[install_reference] {SRC} not present: keeping whatever is in {DST}")
        return os.path.isdir(os.path.join(DST, "lib", "model_zoo"))
    for sub in ("lib", "configs"):
        d = os.path.join(DST, sub)
        if os.path.isdir(d):
            shutil.rmtree(d)
        shutil.copytree(os.path.join(SRC, sub), d, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    for f in ("LICENSE", "README.md", "requirements.txt"):
        if os.path.exists(os.path.join(SRC, f)):
            shutil.copy2(os.path.join(SRC, f), os.path.join(DST, f))
    if verbose:
        print(f"[install_reference] copied lib/ and configs/ of {SRC} to {DST}")
    return True


if __name__ == "__main__":
    sys.exit(0 if install() else 1)
