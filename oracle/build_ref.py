"""TEST INFRASTRUCTURE ONLY.  Byte-compiles the reference's three hot-path modules
from where they lie under /root/reference into oracle/_ref/*.refbc (sourceless CPython
byte-code, git-ignored; no reference SOURCE is copied into the repo; the extension is not
.pyc because the gpurun snapshot drops *.pyc files).  The GPU box has the same
image (same CPython magic), so the byte-code files import there via oracle/ref_shim.py and
let `bench.py --impl reference` time the reference's own get_tree_score on the box's
host cores.  No-op when /root/reference is absent.
"""
import os
import py_compile
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference"
DST = os.path.join(HERE, "_ref")
MODULES = ("clustering", "tree_tools", "tree_io")


def build(verbose=True):
    if not os.path.isfile(os.path.join(SRC, "clustering.py")):
        if verbose:
            print("oracle/build_ref: %s absent, keeping prebuilt oracle/_ref" % SRC)
        return False
    os.makedirs(DST, exist_ok=True)
    for m in MODULES:
        py_compile.compile(
            os.path.join(SRC, m + ".py"),
            cfile=os.path.join(DST, m + ".refbc"),
            dfile=m + ".py",
            doraise=True,
            optimize=0,
            invalidation_mode=py_compile.PycInvalidationMode.UNCHECKED_HASH,
        )
    with open(os.path.join(DST, "PYTHON_VERSION"), "w") as f:
        f.write(sys.version.split()[0] + "\n")
    if verbose:
        print("oracle/build_ref: wrote %s" % DST)
    return True


if __name__ == "__main__":
    build()
