#!/bin/bash
# usage: sweep.sh <preset[:layout]> "<ctas warps pf cpc>" ...   (dev tool)
p=$1; shift
for cfg in "$@"; do set -- $cfg; 
  r=$(python scripts/quick_bench.py 16384 $p "ctas=$1,warps=$2,cpc=$4" 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['Mclips_s'])" 2>/dev/null)
  echo "$p ctas=$1 warps=$2 pf=$3 cpc=$4 -> $r"
done
