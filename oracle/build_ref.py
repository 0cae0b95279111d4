"""Recipe for oracle/_ref: the reference's OWN implementation of the POD path, for the CPU arm of bench.py.

TEST / MEASUREMENT INFRASTRUCTURE.  The reference (pierremtb/POD-UQNN) is pure Python + numba: its hot-path
files `poduqnn/pod.py` (perform_pod :7-47, perform_fast_pod :51-68) and `poduqnn/acceleration.py` (loop_u :13-30,
loop_u_t :34-70) import only numpy and numba, both present in the image.  "Building" the reference therefore means
copying those two files, byte for byte, from where they lie under /root/reference into oracle/_ref/ (git-ignored,
NOT gpurun-ignored, so the copies travel to the GPU box like the built .so files; /root/reference itself does not
exist there).  A manifest with the sha256 of every source and copy is written next to them; `load()` refuses a copy
whose hash no longer matches its manifest entry.

    python oracle/build_ref.py          # run by __graft_entry__.build() when /root/reference is present

`bench.py --impl reference` and the `cpu_baseline` leg import the copies through `load()` (kind "reference") and fall
back to the numpy port in oracle/pod_oracle.py (kind "port") when oracle/_ref was never built.
"""
import hashlib
import importlib.util
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = "/root/reference"
OUT = os.path.join(HERE, "_ref")
FILES = ["poduqnn/pod.py", "poduqnn/acceleration.py"]


def _sha(path):
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def build():
    """Copy the reference's hot-path files verbatim into oracle/_ref/.  Returns the manifest, or None when the
    reference tree is not present (GPU box: the prebuilt copies are used)."""
    if not os.path.isdir(os.path.join(REF_ROOT, "poduqnn")):
        return None
    os.makedirs(OUT, exist_ok=True)
    manifest = {"source_root": REF_ROOT, "files": {}}
    for rel in FILES:
        src = os.path.join(REF_ROOT, rel)
        dst = os.path.join(OUT, "ref_" + os.path.basename(rel))
        shutil.copyfile(src, dst)
        manifest["files"][os.path.basename(dst)] = {"source": rel, "sha256": _sha(src), "bytes": os.path.getsize(src)}
        assert _sha(dst) == manifest["files"][os.path.basename(dst)]["sha256"]
    with open(os.path.join(OUT, "MANIFEST.json"), "w") as f:
        json.dump(manifest, f, indent=1)
    return manifest


def available():
    return os.path.exists(os.path.join(OUT, "MANIFEST.json"))


def load():
    """Import the copied reference modules.  Returns (pod_module, acceleration_module)."""
    if not available():
        raise RuntimeError("oracle/_ref has not been built (python oracle/build_ref.py in the build container)")
    with open(os.path.join(OUT, "MANIFEST.json")) as f:
        manifest = json.load(f)
    mods = []
    for name in ("ref_pod.py", "ref_acceleration.py"):
        path = os.path.join(OUT, name)
        if _sha(path) != manifest["files"][name]["sha256"]:
            raise RuntimeError(f"{path} differs from the reference file it was copied from")
        modname = "oracle_ref_" + name[4:-3]
        if modname in sys.modules:
            mods.append(sys.modules[modname]); continue
        spec = importlib.util.spec_from_file_location(modname, path)
        mod = importlib.util.module_from_spec(spec)
        sys.modules[modname] = mod
        spec.loader.exec_module(mod)
        mods.append(mod)
    return tuple(mods)


if __name__ == "__main__":
    m = build()
    print(json.dumps(m, indent=1) if m else "reference tree not present: nothing built")
