#!/usr/bin/env python
"""Build ``oracle/_ref/``: the UNMODIFIED reference, byte-compiled from the sources where they lie (TEST INFRASTRUCTURE).

The reference is pure Python, so its "binary" is CPython bytecode: every ``*.py`` of /root/reference that the image -> 3D-head
path (and its demo / benchmark callers) can import is compiled with ``py_compile`` straight from /root/reference into
``oracle/_ref/<same relative path>.pyc`` (sourceless-import layout), and the reference's DATA assets the code opens relative
to ``__file__`` (flame.pkl, index sets, yaml configs, the demo image) are copied next to them.  No reference source file is
copied anywhere; ``oracle/_ref/`` is git-ignored (never enters history) but not gpurun-ignored, so -- like the repo's own
built ``.so`` -- it travels to the GPU box, where /root/reference does not exist.  Same interpreter on both sides
(the image's python 3.12), so the bytecode loads there.

Run by ``__graft_entry__.build()`` when /root/reference is present; a no-op otherwise (the prebuilt tree is used).
"""
import os
import py_compile
import shutil
import sys

REF = os.environ.get("DAD3D_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")

# python trees compiled (relative to REF); Sim3DR (Cython build tree) and nothing else is skipped
PY_ROOTS = ["predictor.py", "demo.py", "demo_utils.py", "utils.py", "visualize.py", "__init__.py", "model_training",
            "inference", "dad_3dheads_benchmark"]
# data the compiled code opens relative to __file__ / cwd
DATA = ["dad_3dnet.yaml", "model_training/model/backbone.yaml", "model_training/model/static", "model_training/config",
        "images/demo_heads/1.jpeg", "dad_3dheads_benchmark/data/static"]
SKIP_DATA_SUFFIX = (".py", ".pyc")


def _compile(src, rel):
    dst = os.path.join(OUT, rel + "c")
    os.makedirs(os.path.dirname(dst), exist_ok=True)
    # dfile = the reference-relative name, so tracebacks cite the reference file, not this machine's path
    py_compile.compile(src, cfile=dst, dfile=os.path.join("<reference>", rel), doraise=True,
                       invalidation_mode=py_compile.PycInvalidationMode.UNCHECKED_HASH)


def build(ref: str = REF, out: str = OUT) -> bool:
    if not os.path.isdir(ref):
        return False
    n_py = n_data = 0
    for root in PY_ROOTS:
        p = os.path.join(ref, root)
        if os.path.isfile(p):
            _compile(p, root)
            n_py += 1
            continue
        for d, _, files in os.walk(p):
            for f in files:
                if f.endswith(".py"):
                    full = os.path.join(d, f)
                    _compile(full, os.path.relpath(full, ref))
                    n_py += 1
    for item in DATA:
        p = os.path.join(ref, item)
        if os.path.isfile(p):
            dst = os.path.join(out, item)
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            shutil.copyfile(p, dst)
            n_data += 1
            continue
        for d, _, files in os.walk(p):
            for f in files:
                if f.endswith(SKIP_DATA_SUFFIX):
                    continue
                full = os.path.join(d, f)
                dst = os.path.join(out, os.path.relpath(full, ref))
                os.makedirs(os.path.dirname(dst), exist_ok=True)
                if not (os.path.isfile(dst) and os.path.getsize(dst) == os.path.getsize(full)):
                    shutil.copyfile(full, dst)
                n_data += 1
    # the reference's one native component on the path's "next" rows: the Sim3DR z-buffer rasteriser (C++), compiled from the
    # source file where it lies (same flags as its setup.py: -std=c++11, default -O2, no -march) plus our extern "C" shim
    import subprocess
    so = os.path.join(out, "libsim3dr_ref.so")
    src = os.path.join(ref, "Sim3DR", "lib", "rasterize_kernel.cpp")
    shim = os.path.join(os.path.dirname(os.path.abspath(__file__)), "sim3dr_ref_shim.cpp")
    if os.path.isfile(src):
        subprocess.run(["g++", "-O2", "-std=c++11", "-fPIC", "-shared", "-I", os.path.dirname(src), src, shim, "-o", so],
                       check=True)
    with open(os.path.join(out, "BUILD_INFO.txt"), "w") as f:
        f.write(f"byte-compiled from {ref} by oracle/build_ref.py with python {sys.version.split()[0]}: "
                f"{n_py} modules, {n_data} data files\n")
    return True


if __name__ == "__main__":
    ok = build()
    print("oracle/_ref built" if ok else f"{REF} absent: nothing to do")
