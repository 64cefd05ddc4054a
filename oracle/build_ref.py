"""TEST INFRASTRUCTURE -- recipe that stages the UNMODIFIED reference sources of the hot path under oracle/_ref/.

The reference is pure Python, so "building" it is a verbatim copy of the few source files the CLIP training step lives
in (no edits; sha256 recorded in oracle/_ref/MANIFEST.json).  oracle/_ref/ is git-ignored (reference sources never enter
this repository's history) but not gpurun-ignored, so the copy travels to the GPU box, where /root/reference does not
exist, and bench.py's `--impl reference` arm / `cpu_baseline` leg can time the reference's own `training_step` there.
Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import anything under oracle/.

Run here (build container):  python oracle/build_ref.py      (also called by __graft_entry__.build())
"""
import hashlib
import json
import os
import shutil

REF = os.environ.get("MAVEN_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
# what `from src.models_multimodal import LightCurveImageCLIP` needs: the three files of the path + the module its
# validation hook imports get_AUC from (src/models_multimodal.py:16-18)
FILES = ["src/__init__.py", "src/loss.py", "src/transformer_utils.py", "src/models_multimodal.py", "src/utils.py"]


def build(verbose: bool = False) -> bool:
    """Returns True when oracle/_ref is present and current (False: no reference tree here, nothing staged)."""
    if not os.path.isdir(os.path.join(REF, "src")):
        return os.path.exists(os.path.join(DST, "MANIFEST.json"))
    manifest = {}
    for rel in FILES:
        src, dst = os.path.join(REF, rel), os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        manifest[rel] = hashlib.sha256(open(src, "rb").read()).hexdigest()
    with open(os.path.join(DST, "MANIFEST.json"), "w") as f:
        json.dump({"source": REF, "files": manifest}, f, indent=1)
    if verbose:
        print("staged", len(FILES), "reference files under", DST)
    return True


if __name__ == "__main__":
    print("ok" if build(verbose=True) else "no reference tree at " + REF)
