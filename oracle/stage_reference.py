"""TEST / MEASUREMENT INFRASTRUCTURE -- stages the reference's hot-path files, UNMODIFIED, from the read-only
checkout (/root/reference or $MVS_REFERENCE_SRC) into the git-ignored oracle/_ref/reference/, so that the reference
itself travels to the GPU box with `gpurun` (ignored files ship; /root/reference does not exist there).

Used by bench.py --impl reference (the reference arm: the reference's own closure + LBFGS on the box's host cores, and
optionally on torch-CUDA) and by the -m gpu parity tests that run the unmodified reference closure with its SDF term.
Nothing is copied into git history and nothing under mvsmplfitting_b200/ reads this tree.

    python -m oracle.stage_reference            # copy + build oracle/_ref/libsdf_refcuda.so
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref", "reference")
SRC = os.environ.get("MVS_REFERENCE_SRC", "/root/reference")

# the path named by BASELINE.json north_star + the data files it loads by relative path
FILES = [
    "code/camera.py", "code/prior.py", "code/model/VPoser.py",
    "code/smplx/__init__.py", "code/smplx/body_models_scale.py", "code/smplx/body_models.py", "code/smplx/lbs.py",
    "code/smplx/utils.py", "code/smplx/vertex_ids.py", "code/smplx/vertex_joint_selector.py", "code/smplx/joint_names.py",
    "code/optimizers/__init__.py", "code/optimizers/lbfgs_ls.py", "code/optimizers/optim_factory.py",
    "code/utils/__init__.py", "code/utils/fitting.py", "code/utils/utils.py", "code/utils/prior.py",
    "code/utils/init_guess.py", "code/utils/recompute3D.py", "code/utils/umeyama.py", "code/utils/non_linear_solver.py",
    "code/utils/data_parser.py", "code/utils/FileLoaders.py", "code/utils/module_utils.py", "code/utils/rotation_conversions.py",
    "sdf/sdf/__init__.py", "sdf/sdf/sdf.py", "sdf/sdf/sdf_loss.py",
    "data/J_regressor_lsp.npz", "data/3DOH50K_Parameters.txt",
    "priors/snapshots/poser_epoch091.pkl", "cfg_files/fit_smpl.yaml",
]
TREES = ["data/keypoints"]


def stage(verbose: bool = True) -> str:
    if not os.path.isdir(os.path.join(SRC, "code", "smplx")):
        raise RuntimeError("reference checkout not found at %s" % SRC)
    manifest = {}
    for rel in FILES:
        s, d = os.path.join(SRC, rel), os.path.join(DST, rel)
        if not os.path.exists(s):
            continue
        os.makedirs(os.path.dirname(d), exist_ok=True)
        shutil.copyfile(s, d)
        with open(s, "rb") as f:
            manifest[rel] = hashlib.sha256(f.read()).hexdigest()
    for rel in TREES:
        s, d = os.path.join(SRC, rel), os.path.join(DST, rel)
        if os.path.isdir(s):
            shutil.copytree(s, d, dirs_exist_ok=True)
    with open(os.path.join(DST, "MANIFEST.json"), "w") as f:
        json.dump(dict(source=SRC, files=manifest), f, indent=1)
    if verbose:
        print("staged %d reference files under %s" % (len(manifest), DST))
    return DST


def build_sdf() -> str:
    out = subprocess.check_output([os.path.join(HERE, "build_ref_sdf.sh")], env=dict(os.environ, MVS_REFERENCE_ROOT=SRC))
    return out.decode().strip().splitlines()[-1]


def main(verbose: bool = True) -> None:
    stage(verbose)
    so = os.path.join(HERE, "_ref", "libsdf_refcuda.so")
    src = os.path.join(SRC, "sdf", "sdf", "csrc", "sdf_cuda_kernel.cu")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(os.path.join(HERE, "ref_sdf_wrap.cu"))):
        out = build_sdf()
        if verbose:
            print(out)


if __name__ == "__main__":
    main()
