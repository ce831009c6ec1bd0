"""TEST INFRASTRUCTURE.  Stage the reference's OWN implementation of the hot path under oracle/_ref/ so that it can run on the
GPU box (which has no /root/reference) as the second oracle and as the baseline arm of bench.py.

Usage: python build_ref.py <reference_root> <out_dir>

oracle/_ref/ is git-ignored (nothing of the reference enters the repository's history) but NOT gpurun-ignored, so what this script
writes travels to the GPU box with the snapshot, exactly like the compiled c_gae extension that build_ref_gae.py puts there
(SURVEY.md section 8c).  Staged, unmodified, at their package-relative paths:

    puffer_phc/__init__.py, torch_utils.py, motion_lib.py, poselib_skeleton.py      (query + math + loader)
    puffer_phc/envs/common.py                                                     (obs / reward / reset functions)
    puffer_phc/policies/running_norm.py                                           (RunningNorm; loaded by file path)
    puffer_phc/assets/smpl_humanoid.xml, sample_data/cmu_mocap_05_06.pkl         (BASELINE config 1 inputs)

plus the cythonized + compiled puffer_phc/c_gae.pyx (build_ref_gae.py).  A MANIFEST.json with the sha256 of every staged file is
written so tests can state exactly which reference bytes they ran.
"""
import hashlib
import json
import os
import shutil
import sys

FILES = (
    "puffer_phc/__init__.py",
    "puffer_phc/torch_utils.py",
    "puffer_phc/motion_lib.py",
    "puffer_phc/poselib_skeleton.py",
    "puffer_phc/envs/common.py",
    "puffer_phc/policies/running_norm.py",
    "puffer_phc/assets/smpl_humanoid.xml",
    "sample_data/cmu_mocap_05_06.pkl",
)


def main(ref_root: str, out_dir: str) -> int:
    if not os.path.isdir(os.path.join(ref_root, "puffer_phc")):
        print(f"build_ref: {ref_root}/puffer_phc not found; skipping", file=sys.stderr)
        return 0
    manifest = {}
    for rel in FILES:
        src, dst = os.path.join(ref_root, rel), os.path.join(out_dir, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        with open(dst, "rb") as f:
            manifest[rel] = hashlib.sha256(f.read()).hexdigest()
    with open(os.path.join(out_dir, "MANIFEST.json"), "w") as f:
        json.dump({"source": ref_root, "files": manifest}, f, indent=1)
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, here)
    import build_ref_gae
    return build_ref_gae.main(ref_root, out_dir)


if __name__ == "__main__":
    sys.exit(main(sys.argv[1], sys.argv[2]))
