"""One batch, one mode: the smallest run that shows thread-per-world vs warp-per-world behaviour for ncu.
usage: profile_modes.py SCENE MODE WORLDS STEPS SEED"""
import sys
import torch
sys.path.insert(0, ".")
import fizz2d_b200 as fz
scene, mode, n, steps, seed = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
bw = fz.BatchedWorld.from_in(fz.scenes.scene_path(scene), n_worlds=n, perturb=True, seed=seed, max_calls=256)
bw.set_tuning(mode, 256)
bw.step(steps)
torch.cuda.synchronize()
print(scene, mode, "calls/world-step", bw.calls().sum() / max(1, bw.steps_done().sum()), "capped", int((bw.status() != 0).sum()))
