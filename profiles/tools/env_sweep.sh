#!/bin/bash
# Device-resident step time of the cfg-3 workload under tuning switches (environment variables read by hope_create).
# Usage (under gpurun): bash profiles/tools/env_sweep.sh <out.jsonl> "VAR=a VAR2=b" "VAR=c" ...   (first entry "" = defaults)
out=$1; shift
: > $out
for combo in "$@"; do
  line=$(env $combo python bench.py --steps 100 --warmup 5 --no-cpu-baseline --device-only 2>/dev/null | tail -1)
  echo "{\"env\": \"$combo\", \"result\": $line}" >> $out
done
cat $out
