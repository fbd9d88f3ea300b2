"""GA demo: evolve convex polygons that dissipate the most energy in the PENTAGONVOID scene.
Every generation is one batched GPU run; fitness = world.heat after `steps` timesteps (the scene's PARAM
ELASTICITY is overridden to 0.8 so contacts dissipate).   usage: ga_demo.py [population] [generations] [steps]"""
import sys
import numpy as np
sys.path.insert(0, ".")
import fizz2d_b200 as fz

W = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
G = int(sys.argv[2]) if len(sys.argv) > 2 else 5
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 400
scene = fz.scenes.scene_path("PENTAGONVOID_1SQUARE")
text = open(scene).read() + "\nPARAM ELASTICITY .8\n"
path = "/tmp/ga_scene.in"
open(path, "w").write(text)
pop = fz.ga.random_convex_polygons(20260005, W)
for gen in range(G):
    bw = fz.BatchedWorld.from_ga(path, W, population=pop, max_calls=256)
    bw.step(steps)
    fit = bw.heat()
    st = bw.status()
    fit = np.where(st == 0, fit, -1.0)      # capped worlds are not candidates
    print("generation %d: best %.4g  mean %.4g  capped %d  vertices of best %d" % (
        gen, fit.max(), fit[st == 0].mean(), int((st != 0).sum()), pop[0][int(np.argmax(fit))]))
    bw.close()
    pop = fz.ga.next_generation(pop, fit, 20260005, gen)
