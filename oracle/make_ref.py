#!/usr/bin/env python3
"""Recipe for oracle/_ref: an UNMODIFIED copy of the reference's Python files, made from
/root/reference (or $DSP_REFERENCE_SRC) where it exists -- the build container.  The copy is
git-ignored (never part of this repository's history) but travels to the GPU box with the
snapshot, so that

  * `bench.py --impl reference` / `cpu_baseline` time the reference's own code (kind "reference"),
  * tests/test_scripts_gpu.py runs the reference's run.py / ablation_study.py unchanged on the
    CUDA drop-in modules.

Only tests/, bench.py's CPU legs and __graft_entry__ use it; nothing under
dsp_audioreclabs_b200/ reads oracle/_ref.  Without the source tree this script does nothing.
"""
import os
import shutil
import sys

SRC = os.environ.get("DSP_REFERENCE_SRC", "/root/reference")
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
KEEP_DIRS = ("src", "experiments", "tests")


def main():
    if not os.path.isdir(SRC):
        print(f"make_ref: {SRC} not present; keeping whatever oracle/_ref holds")
        return 0
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    os.makedirs(DST)
    n = 0
    for name in sorted(os.listdir(SRC)):
        p = os.path.join(SRC, name)
        if os.path.isfile(p) and (name.endswith(".py") or name == "requirements.txt"):
            shutil.copyfile(p, os.path.join(DST, name))
            n += 1
        elif os.path.isdir(p) and name in KEEP_DIRS:
            for dirpath, _dirs, files in os.walk(p):
                rel = os.path.relpath(dirpath, SRC)
                os.makedirs(os.path.join(DST, rel), exist_ok=True)
                for f in files:
                    if f.endswith(".py"):
                        shutil.copyfile(os.path.join(dirpath, f), os.path.join(DST, rel, f))
                        n += 1
    with open(os.path.join(DST, "COPIED_FROM"), "w") as f:
        f.write(f"{SRC}\nunmodified copy of {n} Python files made by oracle/make_ref.py; not part of the repository\n")
    print(f"make_ref: {n} files -> {DST}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
