"""Drop-in timestep driver with the reference's CLI (main.py:9-13, :86-121; simulate.sh:12):

    python -m fizz2d_b200.main INPUT NSTEPS [VERBOSITY] [PNG] [ENERGYLOG] [--worlds N --seed S --view I]

ARGV[1] input file, ARGV[2] number of time steps, ARGV[3] verbosity (accepted, unused: the engine prints nothing),
ARGV[4] PNG (spawn ./draw like the reference, main.py:101-116), ARGV[5] energy log.  Per step it writes
`./simulations/plane_%03d.txt` (positions AFTER step t) and, with ENERGYLOG, appends one row
`total,kinetic,potential,heat` taken BEFORE step t to `./simulations/energy.txt` -- the formats and the
off-by-one of the reference.  The extra flags run N perturbed copies of the scene in lockstep on the
GPU and print world `--view`; without them the run is the reference run.
"""
import os
import subprocess
import sys

from . import compat


def _draw(start, end, wait):
    """./draw START END (main.py:106,111) when the reference's rasteriser has been built next to us, else the same frames
    from the GPU rasteriser (fizz2d_b200.raster, pixel-identical to draw.c)."""
    if os.path.isfile("./draw") and os.access("./draw", os.X_OK):
        if wait:
            subprocess.run(["./draw", str(start), str(end)])
            return None
        return subprocess.Popen(["./draw", str(start), str(end), "&"])
    from . import raster
    raster.main([str(start), str(end)])
    return None


def main(argv=None):
    argv = list(sys.argv if argv is None else argv)
    extra = {"--worlds": 1, "--seed": 0, "--view": 0}
    pos = [argv[0]]
    i = 1
    while i < len(argv):
        if argv[i] in extra:
            extra[argv[i]] = int(argv[i + 1])
            i += 2
        else:
            pos.append(argv[i])
            i += 1
    num_iters = int(pos[2])
    png = int(pos[4]) if len(pos) >= 5 else 0          # the reference reads a missing config.PNG here
    energylog = int(pos[5]) if len(pos) >= 6 else 0    # config.ENERGYLOG = 0
    n = extra["--worlds"]
    plane = compat.read_input(pos[1], energylog=energylog, n_worlds=n, perturb=n > 1, seed=extra["--seed"],
                              view=extra["--view"])
    processes = []
    t = 0
    for t in range(num_iters):
        plane.update()
        with open(compat.SIMULATION_DIR + "plane_%03d.txt" % t, "w") as f:
            f.write(str(plane))
        if png and t and t % 126 == 0:
            p = _draw(t - 126, t, wait=False)
            if p is not None:
                processes.append(p)
    if png:
        _draw(t - (t % 126), t, wait=True)
        for p in processes:
            p.wait()
    if energylog:
        plane.log.close()
    return plane


if __name__ == '__main__':
    main()
