"""Build the reference's OWN Cython/OpenMP hot path into oracle/_ref/ (test infrastructure only).

This compiles, unchanged and from where they lie under /root/reference, the two Cython modules that
ARE the reference hot path, and the evaluation module that drives `recommend` in batches:

  * implicit/cpu/_als.pyx    -> oracle/_ref/_als.*.so        (least_squares, least_squares_cg, calculate_loss)
  * implicit/cpu/topk.pyx    -> oracle/_ref/topk.*.so        (topk + implicit/cpu/select.h)
  * implicit/evaluation.pyx  -> oracle/_ref/evaluation.*.so  (ranking_metrics_at_k, train_test_split)

Nothing from the reference is copied into this repository: the .pyx files are cythonized into a
scratch directory under /tmp and only the resulting shared objects are written to oracle/_ref/
(git-ignored; they travel to the GPU box with the gpurun snapshot, like our own built .so files).

The compiler must be /usr/bin/g++ (the image default /opt/gcc/bin/g++ cannot find libgomp.spec).
Flags follow implicit/CMakeLists.txt:1-5 (cython directives) and implicit/cpu/CMakeLists.txt (OpenMP).

Usage:  python oracle/build_ref.py [--force]
"""
import glob
import os
import shutil
import subprocess
import sys
import sysconfig
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REFERENCE = os.environ.get("ALS_B200_REFERENCE", "/root/reference")

MODULES = {
    "_als": "implicit/cpu/_als.pyx",
    "topk": "implicit/cpu/topk.pyx",
    "evaluation": "implicit/evaluation.pyx",
}


def ref_available():
    return all(os.path.exists(os.path.join(REFERENCE, p)) for p in MODULES.values())


def built():
    return all(glob.glob(os.path.join(OUT, f"{m}*.so")) for m in MODULES)


def build(force=False, verbose=False):
    """Returns True if oracle/_ref holds both extension modules after the call."""
    if built() and not force:
        return True
    if not ref_available():
        return False
    import numpy

    os.makedirs(OUT, exist_ok=True)
    ext_suffix = sysconfig.get_config_var("EXT_SUFFIX")
    py_inc = sysconfig.get_paths()["include"]
    tmp = tempfile.mkdtemp(prefix="als_b200_ref_")
    try:
        for mod, rel in MODULES.items():
            src = os.path.join(REFERENCE, rel)
            cxx = os.path.join(tmp, mod + ".cpp")
            cmd = [
                sys.executable, "-m", "cython", "--cplus", "-3",
                "-X", "always_allow_keywords=True,binding=True,embedsignature=True",
                "-I", REFERENCE, "-o", cxx, src,
            ]
            subprocess.run(cmd, check=True, capture_output=not verbose)
            so = os.path.join(OUT, mod + ext_suffix)
            cmd = [
                "/usr/bin/g++", "-O3", "-fopenmp", "-std=c++17", "-fPIC", "-shared", "-w",
                "-DNPY_NO_DEPRECATED_API=NPY_1_7_API_VERSION",
                "-I", py_inc, "-I", numpy.get_include(), "-I", REFERENCE,
                cxx, "-o", so,
            ]
            subprocess.run(cmd, check=True, capture_output=not verbose)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return built()


if __name__ == "__main__":
    ok = build(force="--force" in sys.argv, verbose=True)
    print("oracle/_ref built" if ok else "oracle/_ref NOT built (reference not present)")
    sys.exit(0 if ok else 1)
