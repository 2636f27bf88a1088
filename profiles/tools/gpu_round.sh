#!/bin/bash
# One GPU-box round: parity tests, the bench line, the serialised launch list and one `ncu --set full` capture per hot kernel.
# Usage (under gpurun):  bash profiles/tools/gpu_round.sh <tag> [tests|notests] [kernels-regex]
# Everything lands in gpurun_out/<tag>_*; copy the summaries you want judged into profiles/.
tag=${1:-r02}
tests=${2:-tests}
kern=${3:-"k_advance|k_observe|k_rs_enumerate|k_rs_walk|k_rs_check"}
out=gpurun_out
mkdir -p $out
if [ "$tests" = "tests" ]; then
  python -m pytest tests -m gpu -x -q > $out/${tag}_pytest_gpu.log 2>&1
  tail -3 $out/${tag}_pytest_gpu.log
fi
python bench.py --steps 200 --warmup 5 > $out/${tag}_bench.json 2> $out/${tag}_bench.err
tail -c 600 $out/${tag}_bench.json
# serialised launch list of three steps after warm-up (shares only: cold cache, one kernel at a time)
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --device-only > /dev/null 2>&1
python profiles/tools/launch_shares.py $out/${tag}_launches.csv > $out/${tag}_launch_shares.txt 2>&1
cat $out/${tag}_launch_shares.txt
# one full capture per hot kernel, taken at its 5th launch (steady state)
for k in $(echo "$kern" | tr '|' ' '); do
  ncu --set full --clock-control none --import-source on -k regex:"^$k" -s 5 -c 1 -f -o $out/${tag}_full_$k \
      python bench.py --steps 3 --warmup 3 --no-cpu-baseline --device-only > /dev/null 2>&1
done
ls -la $out | tail -20
