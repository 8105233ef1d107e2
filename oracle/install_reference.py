"""TEST / MEASUREMENT INFRASTRUCTURE ONLY - places an UNMODIFIED copy of the reference's Python package under the
git-ignored `baseline/_ref/` so that the reference sampler itself can be timed on the GPU box, where /root/reference
does not exist (`bench.py --impl reference`, kind "reference").

    python -m oracle.install_reference        # /root/reference/hqvae/**/*.py -> baseline/_ref/hqvae/

The reference has no setup.py / pyproject (nothing `pip install` could build), so the "install" is a verbatim copy of
its .py files; nothing is edited and nothing enters git history (`baseline/_ref/` is in .gitignore, not in
.gpurunignore: it travels with the snapshot like the built .so).  `oracle/ref_shim.py` imports from /root/reference when
it exists and from this copy otherwise.  Run by `__graft_entry__.build()` whenever /root/reference is present.
"""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get("HQ_REFERENCE_SRC", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")


def install(verbose: bool = False) -> bool:
    src_pkg = os.path.join(SRC, "hqvae")
    if not os.path.isdir(src_pkg):
        return False
    n = 0
    for dirpath, _, files in os.walk(src_pkg):
        rel = os.path.relpath(dirpath, SRC)
        for f in files:
            if not f.endswith(".py"):
                continue
            os.makedirs(os.path.join(DST, rel), exist_ok=True)
            shutil.copyfile(os.path.join(dirpath, f), os.path.join(DST, rel, f))
            n += 1
    with open(os.path.join(DST, "SOURCE.txt"), "w") as fh:
        fh.write(f"verbatim copy of {src_pkg}/**/*.py ({n} files) made by oracle/install_reference.py; not product source\n")
    if verbose:
        print(f"installed {n} reference files under {DST}")
    return True


if __name__ == "__main__":
    sys.exit(0 if install(verbose=True) else 1)
