"""Recipe for oracle/_ref/reference_src.zip: a byte-for-byte archive of the reference's pure-Python package and experiment scripts.

TEST / BENCHMARK INFRASTRUCTURE ONLY.  The reference (MScherbela/LaplaceSolverMPS) is four numpy modules with no build
system; `/root/reference` exists in the build container but not on the GPU box.  This script archives

    /root/reference/src/laplace_mps/*.py, src/experiments/*.py, src/scratchpad/test_amen.py   -> oracle/_ref/reference_src.zip

(the built artefact of a pure-Python reference: imported straight from the archive by oracle/ref_loader.py, never unpacked) so that (i) `bench.py --impl reference` times the UNMODIFIED reference code on the box's host cores
(`cpu_baseline.kind = "reference"`) and (ii) `tests/test_experiments.py` runs the reference's own experiment
scripts - the acceptance harness - on top of the `laplace_mps` alias package.  oracle/_ref/ is git-ignored (the
sources never enter this repository's history) but travels to the GPU box with the working tree.  The three
in-memory shims the reference needs on numpy 2.x live in oracle/ref_loader.py; nothing is edited on disk.

    python -m oracle.build_ref
"""
import os
import sys
import zipfile

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/src"
ARCHIVE = os.path.join(HERE, "_ref", "reference_src.zip")


def build(verbose=True):
    if not os.path.isdir(os.path.join(SRC, "laplace_mps")):
        if verbose:
            print("oracle/build_ref: /root/reference not mounted, keeping whatever oracle/_ref holds")
        return os.path.isfile(ARCHIVE)
    os.makedirs(os.path.dirname(ARCHIVE), exist_ok=True)
    tmp = ARCHIVE + f".tmp{os.getpid()}"
    n = 0
    with zipfile.ZipFile(tmp, "w", zipfile.ZIP_DEFLATED) as z:
        for sub, names in (("laplace_mps", None), ("experiments", None), ("scratchpad", ["test_amen.py"])):
            for name in sorted(os.listdir(os.path.join(SRC, sub))):
                if not name.endswith(".py") or (names is not None and name not in names):
                    continue
                z.write(os.path.join(SRC, sub, name), arcname=f"{sub}/{name}")
                n += 1
    os.replace(tmp, ARCHIVE)
    if verbose:
        print(f"oracle/build_ref: {n} files -> {ARCHIVE}")
    return True


if __name__ == "__main__":
    sys.exit(0 if build() else 1)
