#!/usr/bin/env python
"""Install the UNMODIFIED reference under baseline/_ref (git-ignored, shipped to the GPU box by gpurun).

    python tools/install_reference.py

Used by `bench.py --impl reference` (the CPU arm drives the reference's own engine) and by the GPU seam tests that run
the reference's CUDA forward on this repo's `curope` shim.  Only runs where /root/reference exists (the build
container).  Recipe, as the bench contract asks: `pip install --no-index --no-build-isolation --no-deps --find-links
/opt/wheelhouse --target baseline/_ref <copy of /root/reference>` (a /tmp copy because the build writes egg-info into
the source tree and /root/reference is read-only; --no-deps because its requirements - gradio, open3d, viser, roma,
git+https dependencies - cannot be resolved offline).  setup.py only packages `must3r`; its `dust3r` / `croco`
dependencies are git submodules of the checkout, which must3r finds by relative path (must3r/tools/path_to_dust3r.py),
so their python packages are placed next to it with the same layout.  Nothing is edited; INSTALL.json records a
sha256 per file so tests can prove the copy is byte-identical to what this script read.
"""
import hashlib
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference"
DST = os.path.join(ROOT, "baseline", "_ref")
SUBMODULE_DIRS = ["dust3r/dust3r", "dust3r/croco/models", "dust3r/croco/utils"]


def main():
    if not os.path.isdir(SRC):
        print(f"{SRC} not present: nothing to install (baseline/_ref ships prebuilt to the GPU box)")
        return 0
    shutil.rmtree(DST, ignore_errors=True)
    os.makedirs(DST)
    tmp = "/tmp/m3r_ref_src"
    shutil.rmtree(tmp, ignore_errors=True)
    os.makedirs(tmp)
    shutil.copytree(os.path.join(SRC, "must3r"), os.path.join(tmp, "must3r"))
    shutil.copy(os.path.join(SRC, "setup.py"), tmp)
    cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps", "--find-links",
           "/opt/wheelhouse", "--target", DST, tmp]
    r = subprocess.run(cmd, capture_output=True, text=True)
    pip_ok = r.returncode == 0 and os.path.isdir(os.path.join(DST, "must3r"))
    note = "pip install --target ok" if pip_ok else f"pip failed (rc {r.returncode}): {(r.stderr or r.stdout).strip().splitlines()[-1:]}; plain copy used"
    if not pip_ok:
        shutil.rmtree(os.path.join(DST, "must3r"), ignore_errors=True)
        shutil.copytree(os.path.join(SRC, "must3r"), os.path.join(DST, "must3r"))
    for d in SUBMODULE_DIRS:
        shutil.copytree(os.path.join(SRC, d), os.path.join(DST, d), ignore=shutil.ignore_patterns("*.so", "build", "__pycache__"))
    files = {}
    for base, _, names in os.walk(DST):
        for n in names:
            if n.endswith((".py", ".cu", ".cpp")) and "dist-info" not in base:
                p = os.path.join(base, n)
                rel = os.path.relpath(p, DST)
                src = os.path.join(SRC, rel)
                h = hashlib.sha256(open(p, "rb").read()).hexdigest()
                same = os.path.exists(src) and hashlib.sha256(open(src, "rb").read()).hexdigest() == h
                files[rel] = {"sha256": h, "identical_to_reference": same}
    bad = [k for k, v in files.items() if not v["identical_to_reference"]]
    json.dump({"source": SRC, "method": note, "n_files": len(files), "modified": bad, "files": files},
              open(os.path.join(DST, "INSTALL.json"), "w"), indent=1)
    print(f"baseline/_ref: {len(files)} source files, {note}; files differing from {SRC}: {bad or 'none'}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
