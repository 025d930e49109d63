"""oracle/_ref: the UNMODIFIED reference package, installed from /root/reference  --  TEST INFRASTRUCTURE.

Recipe (build container only; the GPU box has no /root/reference and uses the prebuilt oracle/_ref that travels with
the snapshot -- it is git-ignored, never committed):

    python -m pip install --no-index --no-build-isolation --no-deps --target oracle/_ref <copy of /root/reference>

(installed from a copy under /tmp because setuptools writes build/ and *.egg-info into the source tree, and
/root/reference is read-only; --no-deps because pyDOE2 / matplotlib are not in the image -- oracle/ref_runner.py
stubs them at import time).  No reference source is edited; the files under oracle/_ref/ODElib are byte-identical to
/root/reference/ODElib (checked below).  Used by bench.py's CPU legs (`--impl reference`, cpu_baseline, the CPU
chain-steps/s figure) and by tests that compare the oracle restatement with the live reference.
"""
import filecmp
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference"
REF_DST = os.path.join(HERE, "_ref")
FILES = ("__init__.py", "Framework.py", os.path.join("Statistics", "__init__.py"), os.path.join("Statistics", "Samplers.py"),
         os.path.join("Statistics", "stats.py"), os.path.join("Statistics", "distributions.py"))


def available():
    return all(os.path.exists(os.path.join(REF_DST, "ODElib", f)) for f in FILES)


def build(force=False):
    """-> one line describing what oracle/_ref holds."""
    if not os.path.isdir(os.path.join(REF_SRC, "ODElib")):
        return "prebuilt copy present" if available() else "absent (/root/reference not on this box; run in the build container)"
    if available() and not force and all(
            filecmp.cmp(os.path.join(REF_SRC, "ODElib", f), os.path.join(REF_DST, "ODElib", f), shallow=False) for f in FILES):
        return "up to date (byte-identical to /root/reference/ODElib)"
    shutil.rmtree(REF_DST, ignore_errors=True)
    how = "pip install --target"
    with tempfile.TemporaryDirectory() as tmp:
        src = os.path.join(tmp, "reference")
        shutil.copytree(REF_SRC, src, ignore=shutil.ignore_patterns("*.pdf", "demo", ".git"))
        cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps", "--quiet",
               "--find-links", "/opt/wheelhouse", "--target", REF_DST, src]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0 or not available():
            # pip could not build it here: the package is pure Python, a plain copy of the package directory is the
            # same install
            how = "copytree (pip failed: {})".format((res.stderr or res.stdout).strip().splitlines()[-1][:120] if (res.stderr or res.stdout).strip() else "no output")
            shutil.rmtree(REF_DST, ignore_errors=True)
            os.makedirs(REF_DST)
            shutil.copytree(os.path.join(REF_SRC, "ODElib"), os.path.join(REF_DST, "ODElib"))
    for f in FILES:
        if not filecmp.cmp(os.path.join(REF_SRC, "ODElib", f), os.path.join(REF_DST, "ODElib", f), shallow=False):
            raise RuntimeError("oracle/_ref/ODElib/{} differs from the reference".format(f))
    return "installed by {} (byte-identical to /root/reference/ODElib)".format(how)


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
