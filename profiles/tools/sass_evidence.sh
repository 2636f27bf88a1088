#!/bin/bash
# What the shipped library's machine code contains, kernel by kernel: resources, and counts of the SASS mnemonics that matter
# (HMMA = tensor-core MMA, LDSM = ldmatrix, LDGSTS = asynchronous global->shared copy, DFMA/DADD/DMUL = float64 pipe, MUFU.RCP64H,
# SHFL/VOTE = warp collectives, ATOM/RED).  Usage: bash profiles/tools/sass_evidence.sh > profiles/r02_sass_evidence.txt
lib=${1:-hope_b200/libhope_b200.so}
echo "library: $lib  ($(git rev-parse --short HEAD 2>/dev/null))"
cuobjdump -res-usage $lib 2>/dev/null | grep -A1 "Function" | grep -v "^--" | paste - - | sed 's/ Function /\n/; s/^ *//' | awk 'NF' | sed 's/^/  /'
echo
for k in k_advance k_observe k_rs_enumerate k_rs_walk k_rs_check k_rs_select k_pack_lidar k_render k_policy_forward k_norm_partial k_masked_sample k_planner; do
  cuobjdump -sass $lib 2>/dev/null | awk -v K="$k" '
    /Function :/ { f = index($0, K) > 0 }
    f && /^ +\/\*[0-9a-f]+\*\// {
      n++
      if ($0 ~ /HMMA/) hmma++; if ($0 ~ /LDSM/) ldsm++; if ($0 ~ /LDGSTS/) ldgsts++
      if ($0 ~ /DFMA/) dfma++; if ($0 ~ /DADD/) dadd++; if ($0 ~ /DMUL/) dmul++; if ($0 ~ /MUFU/) mufu++
      if ($0 ~ /SHFL/) shfl++; if ($0 ~ /VOTE|MATCH/) vote++; if ($0 ~ /ATOM|RED\./) atom++; if ($0 ~ /LDG/) ldg++; if ($0 ~ /STG/) stg++; if ($0 ~ /LDS/) lds++; if ($0 ~ /BAR\.SYNC/) bar++
    }
    END { printf "%-18s instr %6d  HMMA %4d LDSM %4d LDGSTS %3d | DFMA %5d DADD %5d DMUL %5d MUFU %4d | SHFL %4d VOTE %3d ATOM %3d | LDG %4d STG %4d LDS %4d BAR %3d\n", K, n, hmma, ldsm, ldgsts, dfma, dadd, dmul, mufu, shfl, vote, atom, ldg, stg, lds, bar }'
done
