#!/bin/bash
# Round-end evidence pass on one GPU box: the GPU suite, smoke(), both bench arms with the driver's arguments, the secondary
# configurations.  Everything lands in gpurun_out/<tag>_*.
tag=${1:-r02x}
out=gpurun_out
mkdir -p $out
python -m pytest tests -x -q -m gpu > $out/${tag}_pytest_gpu.txt 2>&1; tail -3 $out/${tag}_pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.txt 2>&1; tail -1 $out/${tag}_smoke.txt
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $out/${tag}_bench_reference.json 2> $out/${tag}_bench_reference.err; tail -c 600 $out/${tag}_bench_reference.json
python bench.py --gpus 1 --steps 20 --warmup 5 > $out/${tag}_bench_driver_args.json 2> $out/${tag}_bench_driver_args.err; tail -c 300 $out/${tag}_bench_driver_args.json
python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err
for c in image cfg2 dlp; do python bench.py --config $c --steps 50 --warmup 25 > $out/${tag}_bench_$c.json 2>/dev/null; done
python bench.py --config rollout --steps 50 --warmup 25 > $out/${tag}_bench_rollout.json 2>/dev/null
python bench.py --config rollout --image --steps 50 --warmup 25 > $out/${tag}_bench_rollout_img.json 2>/dev/null
python bench.py --config sac --steps 64 --warmup 16 > $out/${tag}_bench_sac.json 2>/dev/null
for f in $out/${tag}_bench*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    sec={k:round(v["value"]/1e6,2) for k,v in (d.get("secondary") or {}).items()}
    print(sys.argv[1].split("/")[-1], round(d["value"]/1e6,3), d.get("ms_per_step"), "e2e", (d.get("e2e") or {}).get("value"), sec)
except Exception as e:
    print(sys.argv[1], "unreadable", e)
PY
done
