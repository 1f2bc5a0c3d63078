"""Build recipe for ``oracle/_ref/`` (TEST INFRASTRUCTURE ONLY -- never imported by the product package).

The reference is pure Python, so "compiling it from its own sources where they lie" means byte-compiling the
two modules that own rows of the hot path,

    /root/reference/kt_service/ai_tools/utils.py                     (a3, a5-a10, a15-a20)
    /root/reference/kt_service/ai_tools/mesh_tools/femm_generator.py (a21-a24; imports gmsh / shapely, stubbed)

into ``oracle/_ref/*.pyc``.  ``oracle/_ref/`` is git-ignored (no reference source enters the history) but not
gpurun-ignored, so the byte code travels to the GPU box like the repo's own built ``.so`` files; there the
parity tests and ``bench.py --impl reference`` run the ACTUAL reference functions for the rows they own.

    python -m oracle.build_ref            # run by __graft_entry__.build() when /root/reference is present
"""
from __future__ import annotations

import os
import py_compile
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REF_ROOT = os.environ.get("EITB_REFERENCE_ROOT", "/root/reference")
MODULES = {
    "ref_utils": os.path.join("kt_service", "ai_tools", "utils.py"),
    "ref_femm_generator": os.path.join("kt_service", "ai_tools", "mesh_tools", "femm_generator.py"),
}


def build() -> list:
    """Byte-compile the reference modules; returns the files written ([] when the reference tree is absent)."""
    done = []
    if not os.path.isdir(REF_ROOT):
        return done
    os.makedirs(OUT, exist_ok=True)
    for name, rel in MODULES.items():
        src = os.path.join(REF_ROOT, rel)
        if not os.path.isfile(src):
            continue
        dst = os.path.join(OUT, f"{name}.cpython-{sys.version_info.major}{sys.version_info.minor}.pyc")
        py_compile.compile(src, cfile=dst, dfile=rel, doraise=True, optimize=0)
        done.append(dst)
    with open(os.path.join(OUT, "README"), "w") as f:
        f.write("byte code of the unmodified reference modules, built by oracle/build_ref.py; not tracked by git\n")
    return done


if __name__ == "__main__":
    for p in build():
        print(p)
