"""Recipe for oracle/_ref/: the UNMODIFIED reference package, made importable on the GPU box.

TEST / BENCH INFRASTRUCTURE ONLY.  The reference (dsilvestro/npBNN) is pure Python; /root/reference exists in the build
container but not on the GPU box, while untracked files of the working tree do travel there.  This script installs the
reference package into the git-ignored directory oracle/_ref/ (pip --target from a scratch copy, because the source tree
is read-only; a plain copy of the np_bnn/ package directory if pip is not usable) so that

    bench.py --impl reference        and        bench.py's cpu_baseline leg

time the reference's own code (np_bnn.MCMC.mh_step, np_bnn.MC3.run_mcmc) on the box's host cores.  Nothing under
oracle/_ref/ is committed, nothing in the product path (npbnn_b200/, np_bnn/) reads it, and the -m gpu tests / smoke()
do not need it.

    python oracle/make_ref.py [--force]
"""
import filecmp
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference"
DEST = os.path.join(HERE, "_ref")


def available():
    return os.path.isfile(os.path.join(DEST, "np_bnn", "__init__.py"))


def _same_tree():
    """True when oracle/_ref/np_bnn holds byte-identical copies of every reference module."""
    src = os.path.join(REF_SRC, "np_bnn")
    names = [n for n in os.listdir(src) if n.endswith(".py")]
    match, mismatch, errors = filecmp.cmpfiles(src, os.path.join(DEST, "np_bnn"), names, shallow=False)
    return not mismatch and not errors


def make(force=False, verbose=True):
    """Returns 'present' | 'pip' | 'copy' | 'no-source'."""
    if not os.path.isdir(os.path.join(REF_SRC, "np_bnn")):
        return "present" if available() else "no-source"
    if available() and not force and _same_tree():
        return "present"
    shutil.rmtree(DEST, ignore_errors=True)
    os.makedirs(DEST, exist_ok=True)
    how = "copy"
    with tempfile.TemporaryDirectory() as tmp:
        work = os.path.join(tmp, "src")
        shutil.copytree(REF_SRC, work, ignore=shutil.ignore_patterns("example_files", ".git", "__pycache__"))
        cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps", "--quiet",
               "--find-links", "/opt/wheelhouse", "--target", DEST, work]
        try:
            r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
            if r.returncode == 0 and available():
                how = "pip"
            elif verbose:
                sys.stderr.write("pip install of the reference failed (%s); copying the package directory\n"
                                 % (r.stderr.strip().splitlines()[-1:] or ["?"])[0])
        except Exception as e:                                  # pip missing / timeout: fall through to the copy
            if verbose:
                sys.stderr.write("pip not usable (%s); copying the package directory\n" % e)
    if how == "copy":
        shutil.rmtree(os.path.join(DEST, "np_bnn"), ignore_errors=True)
        shutil.copytree(os.path.join(REF_SRC, "np_bnn"), os.path.join(DEST, "np_bnn"),
                        ignore=shutil.ignore_patterns("__pycache__"))
    assert available() and _same_tree(), "oracle/_ref/np_bnn is not a faithful copy of the reference"
    with open(os.path.join(DEST, "PROVENANCE.txt"), "w") as f:
        f.write("unmodified copy of %s/np_bnn (dsilvestro/npBNN), made by oracle/make_ref.py (%s); not committed\n"
                % (REF_SRC, how))
    return how


if __name__ == "__main__":
    print("oracle/_ref:", make(force="--force" in sys.argv))
