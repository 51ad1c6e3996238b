#!/bin/sh
# End-to-end demo of the job farm on N GPUs (default 2): preprocess -> run (torchrun, lines sharded over the ranks) -> postprocess.
#   scripts/farm_demo.sh [N] [workdir]
set -e
N=${1:-2}
W=${2:-/tmp/cyl_farm_demo}
ROOT=$(cd "$(dirname "$0")/.." && pwd)
rm -rf "$W" && mkdir -p "$W" && cd "$W"
cat > infile.txt <<'EOT'
--n_steps 40
--field_type lattice
--method sequential
--alpha -1
--C 1
--u 1
--n 2
--kappa 0
--gamma 1
--temp .05
--intrinsic_curvature 0
--amplitude 0
--radius 1
--wavenumber 1
--measure_every 1
--fieldsteps_per_ampstep 50
--dims 50 50
EOT
PYTHONPATH=$ROOT python -m cylinder_b200.farm preprocess -n "demo scan" --var1name alpha --var1range -2 0.1 .5 --var2name wavenumber --var2range .6 1.3 .1
PYTHONPATH=$ROOT python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 -m cylinder_b200.farm run --input infile.txt --varfile varfile.txt --outdir out
PYTHONPATH=$ROOT python -m cylinder_b200.farm postprocess --outdir out
ls out | wc -l
head -3 out/abs_amplitude.csv
