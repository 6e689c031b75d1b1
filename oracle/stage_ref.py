"""TEST / BASELINE INFRASTRUCTURE ONLY -- stage the UNMODIFIED reference under oracle/_ref/.

    python oracle/stage_ref.py            # copies /root/reference/src -> oracle/_ref/src (git-ignored)

The reference is a pure-Python source tree, not an installable package (no setup.py / pyproject), so the
"install" of bench.py's reference arm is a verbatim copy of its `src/` package next to the import shim
(oracle/refshim.py supplies the three absent third-party imports and redirects the hard-coded .cuda() calls to
the CPU).  oracle/_ref/ is listed in .gitignore -- reference sources never enter the history -- but not in
.gpurunignore, so the copy travels to the GPU box, where /root/reference does not exist.  Nothing under
kelpie_b200/ may import it; only `bench.py --impl reference` does."""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")


def stage(reference_root="/root/reference", quiet=False):
    src = os.path.join(reference_root, "src")
    if not os.path.isdir(src):
        return False  # not in the build container: keep whatever was staged before
    dst = os.path.join(DEST, "src")
    if os.path.isdir(dst):
        shutil.rmtree(dst)
    shutil.copytree(src, dst, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    with open(os.path.join(DEST, "STAGED_FROM"), "w") as f:
        f.write(reference_root + "\n")
    if not quiet:
        n = sum(len(files) for _, _, files in os.walk(dst))
        print(f"staged {n} files of the unmodified reference under {dst}")
    return True


if __name__ == "__main__":
    sys.exit(0 if stage(*(sys.argv[1:2])) else 1)
