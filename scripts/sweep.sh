#!/bin/bash
# usage: sweep.sh <preset[:layout]> "<ctas warps pf cpc>" ...   (dev tool)
p=$1; shift
for cfg in "$@"; do set -- $cfg; 
  r=$(SRFE_CTAS=$1 SRFE_WARPS=$2 SRFE_CPC=$4 python scripts/quick_bench.py 16384 $p 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['Mclips_s'])" 2>/dev/null)
  echo "$p ctas=$1 warps=$2 pf=$3 cpc=$4 -> $r"
done
