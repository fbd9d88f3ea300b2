#!/bin/bash
# Drop-in for the reference's CLI wrapper (/root/reference/simulate.sh:1-35), with its line 12 -- `python ./src/main.py
# $1 $2 $3 $4 $5` -- switched to the batched B200 engine.  Same arguments, same outputs:
#
#   ./simulate.sh INPUT NSTEPS [VERBOSITY] [PNG] [ENERGYLOG] [--worlds N --seed S --view I]
#
#   simulations/plane_%03d.txt   positions AFTER every step          (main.py:99-100)
#   simulations/energy.txt       one row BEFORE every step           (physics.py:81-82), with ENERGYLOG
#   simulations/plane_%03d.png   frames, with PNG                    (draw.c); ./draw if it exists, else the GPU rasteriser
#   simulations/simul.gif        when ffmpeg is installed            (simulate.sh:31-35)
# The extra flags run N perturbed copies of the scene in lockstep on the GPU; the files show world --view.
cd "$(dirname "$0")"
mkdir -p simulations
count_pngs=`ls -1 simulations/*.png 2>/dev/null | wc -l`
if [ $count_pngs != 0 ]
then
rm ./simulations/*png
fi
count_txts=`ls -1 simulations/*.txt 2>/dev/null | wc -l`
if [ $count_txts != 0 ]
then
rm ./simulations/*txt
fi
python -m fizz2d_b200.main "$@" || exit $?
count_pngs=`ls -1 simulations/*.png 2>/dev/null | wc -l`
if [ $count_pngs != 0 ] && command -v ffmpeg >/dev/null 2>&1
then
  yes | ffmpeg -i simulations/plane_%03d.png simulations/simul.gif
fi
