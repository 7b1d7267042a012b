"""Stage the UNMODIFIED reference under baseline/_ref/StyleSinger (git-ignored, NOT gpurun-ignored) so that it travels to
the GPU box with the snapshot, like the built .so does.

    python tools/stage_reference.py [--src /root/reference]

Called by __graft_entry__.build() whenever /root/reference is present (the build container).  Only Python sources and
small config / JSON files are copied, byte for byte; audio and binary artefacts (infer_out/, test/*.wav) stay behind.
The staged tree is used ONLY by the reference arms (bench.py --impl reference, tools/baseline_arms.py) and by the
tests that run this repo's drop-ins inside the reference's own registries (tests/test_gpu_reference_dropin.py); the
product package never imports it.
"""
import argparse
import filecmp
import os
import shutil
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DST = os.path.join(REPO, "baseline", "_ref", "StyleSinger")
KEEP_EXT = (".py", ".yaml", ".yml", ".json", ".txt", ".md")
SKIP_DIRS = {"infer_out", "test", ".git", "__pycache__"}


def stage(src="/root/reference", dst=DST, verbose=True):
    if not os.path.isdir(src):
        return None
    n = 0
    for root, dirs, files in os.walk(src):
        dirs[:] = [d for d in dirs if d not in SKIP_DIRS]
        rel = os.path.relpath(root, src)
        for f in files:
            if not f.endswith(KEEP_EXT) and f != "LICENSE":
                continue
            s = os.path.join(root, f)
            d = os.path.join(dst, rel, f)
            os.makedirs(os.path.dirname(d), exist_ok=True)
            if not (os.path.exists(d) and filecmp.cmp(s, d, shallow=False)):
                shutil.copyfile(s, d)
            n += 1
    with open(os.path.join(dst, "STAGED_FROM"), "w") as fh:
        fh.write(f"{src}\n{n} files, copied unmodified by tools/stage_reference.py\n")
    if verbose:
        print(f"staged {n} reference files under {os.path.relpath(dst, REPO)}")
    return dst


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--src", default="/root/reference")
    a = ap.parse_args()
    sys.exit(0 if stage(a.src) else 1)
