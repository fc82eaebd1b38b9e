"""Installs the UNMODIFIED reference package into the git-ignored ``baseline/_ref`` (base contract's reference arm).

    python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse --target baseline/_ref <copy of /root/reference>

(from a copy under /tmp because /root/reference is read-only; --no-deps because deepspeed / open3d / pytorch3d / pypcd are
not in the offline wheelhouse).  The reference's setup.cfg declares ``packages = dprt`` only, so the wheel pip builds holds
the four top-level modules and none of the sub-packages (dprt.models, dprt.datasets, ...): the install is then completed
with the sub-package files taken verbatim from the same source tree.  Nothing here enters git history; the directory
travels to the GPU box with gpurun, where /root/reference does not exist.  Called by ``__graft_entry__.build()``.
"""
import json
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE = "/root/reference"
TARGET = os.path.join(ROOT, "baseline", "_ref")


def install(force: bool = False) -> dict:
    marker = os.path.join(TARGET, "INSTALL.json")
    if not os.path.isdir(os.path.join(REFERENCE, "src", "dprt")):
        return {"status": "reference source absent (GPU box): using the prebuilt baseline/_ref" if os.path.exists(marker)
                else "reference source absent and nothing installed"}
    if os.path.exists(marker) and not force:
        with open(marker) as f:
            return json.load(f)
    shutil.rmtree(TARGET, ignore_errors=True)
    os.makedirs(TARGET, exist_ok=True)
    tmp = tempfile.mkdtemp(prefix="dpft_ref_")
    src = os.path.join(tmp, "reference")
    shutil.copytree(REFERENCE, src)
    cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps", "--find-links",
           "/opt/wheelhouse", "--target", TARGET, src]
    r = subprocess.run(cmd, capture_output=True, text=True)
    info = {"pip_cmd": " ".join(cmd[:-1]) + " <copy of /root/reference>", "pip_rc": r.returncode, "pip_tail": r.stdout[-300:] + r.stderr[-300:]}
    # complete the install: sub-packages the reference's setup.cfg (packages = dprt) leaves out of its own wheel
    added = []
    pkg_src, pkg_dst = os.path.join(REFERENCE, "src", "dprt"), os.path.join(TARGET, "dprt")
    for dirpath, dirnames, filenames in os.walk(pkg_src):
        dirnames[:] = [d for d in dirnames if d != "__pycache__"]
        rel = os.path.relpath(dirpath, pkg_src)
        for fn in filenames:
            if not fn.endswith(".py"):
                continue
            dst = os.path.join(pkg_dst, rel, fn)
            if not os.path.exists(dst):
                os.makedirs(os.path.dirname(dst), exist_ok=True)
                shutil.copyfile(os.path.join(dirpath, fn), dst)
                added.append(os.path.normpath(os.path.join(rel, fn)))
    shutil.rmtree(tmp, ignore_errors=True)
    info.update(status="installed", files_from_wheel=sorted(f for f in os.listdir(pkg_dst) if f.endswith(".py")),
                subpackage_files_added=len(added),
                note="setup.cfg of the reference lists packages = dprt only; sub-packages copied verbatim from /root/reference/src/dprt")
    with open(marker, "w") as f:
        json.dump(info, f, indent=1)
    return info


if __name__ == "__main__":
    print(json.dumps(install(force="--force" in sys.argv), indent=1))
