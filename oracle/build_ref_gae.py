"""TEST INFRASTRUCTURE.  Compile the reference's own native GAE (puffer_phc/c_gae.pyx) into oracle/_ref/.

Usage: python build_ref_gae.py <reference_root> <out_dir>

The .pyx is cythonized and compiled from where it lies under the reference checkout; only generated
outputs (c_gae.c, c_gae*.so) are written, and only into <out_dir> (git-ignored).  No reference source is
copied into the repository.  The reference builds the same file at import time through pyximport with
default flags (puffer_phc/clean_pufferl/core.py:33-36); we use the same default flags (no -march, so
no FMA contraction).
"""
import os
import subprocess
import sys
import sysconfig


def main(ref_root: str, out_dir: str) -> int:
    import numpy as np
    from Cython.Build.Dependencies import cythonize  # noqa: F401  (checks availability)
    from Cython.Compiler.Main import CompilationOptions, compile as cy_compile

    pyx = os.path.join(ref_root, "puffer_phc", "c_gae.pyx")
    if not os.path.isfile(pyx):
        print(f"build_ref_gae: {pyx} not found; skipping", file=sys.stderr)
        return 0
    os.makedirs(out_dir, exist_ok=True)
    c_file = os.path.join(out_dir, "c_gae.c")
    opts = CompilationOptions(output_file=c_file, language_level=3)
    res = cy_compile(pyx, opts)
    if res.num_errors:
        return 1
    ext = sysconfig.get_config_var("EXT_SUFFIX")
    so = os.path.join(out_dir, "c_gae" + ext)
    cmd = [
        os.environ.get("CC", "gcc"), "-O2", "-fPIC", "-shared", "-fwrapv", "-fno-strict-aliasing",
        "-DNPY_NO_DEPRECATED_API=NPY_1_7_API_VERSION",
        "-I" + sysconfig.get_paths()["include"], "-I" + np.get_include(),
        c_file, "-o", so,
    ]
    subprocess.check_call(cmd)
    print("built", so)
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv[1], sys.argv[2]))
