#!/usr/bin/env python
"""Recipe for oracle/_ref/: the reference's OWN implementation of the path, for timing beside the GPU and for checking the
restatement (TEST INFRASTRUCTURE; never imported by the product path).

cyliustack/sofa is pure Python - there is nothing to compile.  `build_ref()` copies the modules `bin/sofa` imports
(`/root/reference/bin/*.py`, 230 KB) into oracle/_ref/bin/ where they lie unmodified.  oracle/_ref/ is git-ignored (no
reference source enters the history) but not gpurun-ignored, so it travels to the GPU box like the built .so files, where
`bench.py --impl reference` and the cpu_baseline leg run it (oracle/ref_runner.py) on the host cores; /root/reference
itself does not exist there.  Run by `__graft_entry__.build()` whenever /root/reference is present.

    python oracle/build_ref.py
"""
import glob
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/bin"
DST = os.path.join(HERE, "_ref", "bin")


def build_ref(src=SRC, dst=DST):
    """-> number of files copied (0 when the reference is not there: the GPU box only uses what was built here)"""
    if not os.path.isdir(src):
        return 0
    os.makedirs(dst, exist_ok=True)
    n = 0
    for p in sorted(glob.glob(os.path.join(src, "*.py"))):
        q = os.path.join(dst, os.path.basename(p))
        if not os.path.isfile(q) or open(p, "rb").read() != open(q, "rb").read():
            shutil.copyfile(p, q)
            os.chmod(q, 0o644)
        n += 1
    return n


if __name__ == "__main__":
    print("oracle/_ref/bin: %d modules" % build_ref())
