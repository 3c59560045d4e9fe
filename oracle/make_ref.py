"""Stage the reference's python sources for the GPU box: /root/reference/**/*.py -> oracle/_ref/.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

The reference (y2w-oc/CMR-Agent) is pure Python with no build system, so "building" it means making
its files importable where the GPU tests run.  `/root/reference` exists only in the build container;
`oracle/_ref/` is git-ignored (nothing of the reference enters the history) but NOT gpurun-ignored, so
it travels to the GPU box next to libcmr_b200.so.  What uses it there:

  * tests/test_gpu_reference_callers.py - the reference's own CMRAgent / PointNN / Buffer / Test_Agent-
    and Train_Agent-style loops running unchanged on the drop-ins,
  * tests/test_oracle_vs_reference.py    - the oracle pinned against the real reference,
  * bench.py's informational `gpu_torch_baseline` - the reference's environment.py on CUDA tensors.

Only `*.py` files are staged (no checkpoints, label maps or byte-code); files are copied verbatim and a
MANIFEST with their sha256 is written, so a test can tell which reference it ran against.

    python oracle/make_ref.py            # stage (idempotent)
    python oracle/make_ref.py --check    # exit 1 if oracle/_ref is missing or stale
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("CMR_REFERENCE_SRC", "/root/reference")
DST = os.path.join(HERE, "_ref")
PACKAGES = ("config", "dataset", "environment", "models", "utils")


def _sources():
    out = []
    for name in sorted(os.listdir(SRC)):
        p = os.path.join(SRC, name)
        if os.path.isfile(p) and name.endswith(".py"):
            out.append(name)
    for pkg in PACKAGES:
        root = os.path.join(SRC, pkg)
        for dirpath, dirnames, filenames in os.walk(root):
            dirnames[:] = sorted(d for d in dirnames if d != "__pycache__")
            for f in sorted(filenames):
                if f.endswith(".py"):
                    out.append(os.path.relpath(os.path.join(dirpath, f), SRC))
    return out


def _sha(path):
    with open(path, "rb") as fh:
        return hashlib.sha256(fh.read()).hexdigest()


def available():
    return os.path.isfile(os.path.join(SRC, "environment", "environment.py"))


def stage(check_only=False):
    if not available():
        if check_only:
            return os.path.isfile(os.path.join(DST, "MANIFEST.json"))
        raise SystemExit(f"{SRC} is not present: oracle/_ref can only be staged in the build container")
    manifest = {rel: _sha(os.path.join(SRC, rel)) for rel in _sources()}
    mpath = os.path.join(DST, "MANIFEST.json")
    if os.path.isfile(mpath):
        with open(mpath) as fh:
            if json.load(fh).get("files") == manifest and all(
                    os.path.isfile(os.path.join(DST, rel)) for rel in manifest):
                return True
    if check_only:
        return False
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    for rel in manifest:
        dst = os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(SRC, rel), dst)
    with open(mpath, "w") as fh:
        json.dump({"source": "y2w-oc/CMR-Agent (" + SRC + ")", "files": manifest}, fh, indent=1, sort_keys=True)
    return True


if __name__ == "__main__":
    ok = stage(check_only="--check" in sys.argv)
    print("oracle/_ref:", "up to date" if ok else "missing or stale")
    sys.exit(0 if ok else 1)
