import sys, os, numpy as np, torch
sys.path.insert(0, os.getcwd())
from types import SimpleNamespace
from puffer_phc_b200 import synth
from puffer_phc_b200.motion_file import RawClips
from puffer_phc_b200.motion_lib import MotionLibSMPL
from puffer_phc_b200.skeleton import SkeletonTree
DEV="cuda:0"
T = synth.make_motion_library(11313, seed=0, device=DEV)
nf = T["num_frames"].cpu().numpy(); fps = np.round(1.0 / T["motion_dt"].cpu().numpy()).astype(np.int32)
raw = RawClips.from_device([f"c{i}" for i in range(len(nf))], nf, fps, T["gts"][:, 0].double().contiguous(), T["motion_aa"].double().contiguous(), T["grs"].double().contiguous())
parents = [-1, 0, 1, 2, 3, 0, 5, 6, 7, 0, 9, 10, 11, 12, 11, 14, 15, 16, 17, 11, 19, 20, 21, 22]
sk = SkeletonTree([f"b{j}" for j in range(24)], np.array(parents, np.int32), np.random.default_rng(0).normal(0, 0.15, (24, 3)).astype(np.float32))
del T
cfg = SimpleNamespace(motion_file=raw, device=DEV, min_length=-1, max_length=300, im_eval=False, is_deterministic=True, step_dt=1 / 30)
lib = MotionLibSMPL(cfg)
n=11313
for _ in range(3):
    lib.load_motions(skeleton_trees=[sk]*n, gender_betas=torch.zeros(n,17), limb_weights=np.zeros((n,10)), sample_idxes=torch.arange(n))
torch.cuda.synchronize()
