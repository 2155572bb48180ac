#!/usr/bin/env python
"""Recipe for `oracle/_ref/`: the UNMODIFIED reference, placed where the GPU box can import it.

The reference (Nobita421/signature-Gan) is plain Python under `src/` with no build system, so "building" it is
copying its source tree, byte for byte, from where it lies (`/root/reference/src`, or $SIGGAN_REFERENCE_SRC) into
`oracle/_ref/src/` — a git-ignored output directory that is NOT gpurun-ignored, so it travels to the GPU box next to the
built `.so` files while the repository's history stays free of reference sources. A manifest with the SHA-256 of every
file is written beside it; `tests/` and `bench.py` check it before using the copy.

Used by: `bench.py --impl reference` and the `cpu_baseline` legs (the reference's own `VanillaGAN(device='cpu')`
timed on the host cores, kind "reference"), and the "runs unchanged" tests that drive the reference's caller code
(`utils/inference.py`, `train_vanilla_gan_signatures.py`) over this repository's drop-in modules.

    python oracle/make_ref.py            # (re)build oracle/_ref from /root/reference/src
    python oracle/make_ref.py --check    # verify an existing copy against its manifest
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
SRC_DEFAULT = "/root/reference/src"


def _sha(path: str) -> str:
    h = hashlib.sha256()
    with open(path, "rb") as f:
        h.update(f.read())
    return h.hexdigest()


def _files(root: str):
    for d, _, names in sorted(os.walk(root)):
        if "__pycache__" in d:
            continue
        for n in sorted(names):
            if n.endswith(".py"):
                yield os.path.relpath(os.path.join(d, n), root)


def build(src: str = None) -> bool:
    """Copy the reference's `src/` tree into oracle/_ref/src. Returns False (and leaves any existing copy alone) when
    the reference tree is not mounted — the GPU box uses the copy made in the build container."""
    src = src or os.environ.get("SIGGAN_REFERENCE_SRC", SRC_DEFAULT)
    if not os.path.isdir(src):
        return False
    out = os.path.join(DEST, "src")
    if os.path.isdir(DEST):
        shutil.rmtree(DEST)
    os.makedirs(out)
    manifest = {}
    for rel in _files(src):
        dst = os.path.join(out, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(src, rel), dst)
        manifest[rel] = _sha(dst)
    with open(os.path.join(DEST, "MANIFEST.json"), "w") as f:
        json.dump({"source": src, "files": manifest}, f, indent=1, sort_keys=True)
    return True


def check() -> bool:
    """True when oracle/_ref exists and every file still has the hash recorded when it was copied."""
    try:
        with open(os.path.join(DEST, "MANIFEST.json")) as f:
            manifest = json.load(f)["files"]
    except OSError:
        return False
    return bool(manifest) and all(
        os.path.exists(os.path.join(DEST, "src", rel)) and _sha(os.path.join(DEST, "src", rel)) == h
        for rel, h in manifest.items())


def src_dir() -> str:
    """Directory to put on sys.path to import the reference's modules; raises when the copy is missing or altered."""
    if not check():
        raise FileNotFoundError(
            "oracle/_ref is missing or was modified: run `python oracle/make_ref.py` where /root/reference is mounted")
    return os.path.join(DEST, "src")


if __name__ == "__main__":
    if "--check" in sys.argv:
        ok = check()
        print("oracle/_ref:", "ok" if ok else "missing or modified")
        sys.exit(0 if ok else 1)
    if not build():
        print("reference tree not found; oracle/_ref left as it is", file=sys.stderr)
        sys.exit(1)
    print("oracle/_ref built:", len(list(_files(os.path.join(DEST, "src")))), "files")
