#!/usr/bin/env python
"""Stage the reference's own Python programs for this path under oracle/_ref/ (git-ignored, NOT gpurun-ignored).

TEST / BASELINE INFRASTRUCTURE ONLY.  /root/reference exists in the build container but not on the GPU box; the files
this recipe stages travel to the box with the snapshot, exactly like the built libatlj.so does, so that

  * bench.py's cpu_baseline leg can time the reference's NumPy routines on the box's own host cores
    (oracle/time_reference_python.py; BASELINE.json north_star, SURVEY.md 8d "CPU baseline beside it" item 1), and
  * tests/test_reference_driver.py can execute the UNMODIFIED driver scripts (md_nve_lj.py, bd_nvt_lj.py, ...) with
    examples_b200/dropin first on PYTHONPATH.

Nothing is edited: the files are byte copies, and MANIFEST.json records the sha256 of every one so that a test can
verify the staged tree still is the reference.  Nothing under examples_b200/ reads oracle/_ref.

    python oracle/make_ref.py            (called by __graft_entry__.build() when /root/reference is present)
"""
import hashlib
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("ATLJ_REFERENCE", "/root/reference")
DST = os.path.join(ROOT, "oracle", "_ref", "python_examples")

# the drivers of SURVEY.md 8b, the force modules they import, and the helper modules those drivers import
FILES = [
    "md_nve_lj.py", "md_nvt_lj.py", "bd_nvt_lj.py", "md_nvt_lj_le.py", "md_npt_lj.py", "smc_nvt_lj.py",
    "md_lj_module.py", "md_lj_ll_module.py", "md_lj_le_module.py", "md_lj_llle_module.py", "smc_lj_module.py",
    "averages_module.py", "config_io_module.py", "lrc_module.py", "maths_module.py", "initialize.py",
]


def sha256(path):
    h = hashlib.sha256()
    with open(path, "rb") as fh:
        h.update(fh.read())
    return h.hexdigest()


def stage(verbose=True):
    src_dir = os.path.join(REF, "python_examples")
    if not os.path.isdir(src_dir):
        if verbose:
            print("make_ref: %s not present (GPU box?): keeping whatever oracle/_ref holds" % src_dir)
        return False
    os.makedirs(DST, exist_ok=True)
    manifest = {}
    for name in FILES:
        src = os.path.join(src_dir, name)
        if not os.path.exists(src):
            continue
        shutil.copyfile(src, os.path.join(DST, name))
        manifest[name] = sha256(src)
    with open(os.path.join(DST, "MANIFEST.json"), "w") as fh:
        json.dump({"source": src_dir, "sha256": manifest}, fh, indent=1, sort_keys=True)
    if verbose:
        print("make_ref: staged %d reference files under %s" % (len(manifest), DST))
    return True


def staged_dir():
    """-> oracle/_ref/python_examples if it holds the staged reference (manifest verified), else None."""
    mf = os.path.join(DST, "MANIFEST.json")
    if not os.path.exists(mf):
        return None
    man = json.load(open(mf))["sha256"]
    for name, digest in man.items():
        p = os.path.join(DST, name)
        if not os.path.exists(p) or sha256(p) != digest:
            return None
    return DST


if __name__ == "__main__":
    sys.exit(0 if stage() else 1)
