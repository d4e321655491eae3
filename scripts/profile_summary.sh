#!/bin/bash
# usage: scripts/profile_summary.sh <rep.ncu-rep> <out.txt>   -- text summary of one ncu capture for profiles/
set -e
rep=$1; out=$2
ncu -i "$rep" --page source --csv 2>/dev/null > /tmp/_src.csv
{
  echo "# ncu --set full --clock-control none, one launch; source: $(basename $rep)"
  echo "## raw metrics"; python scripts/ncu_raw.py "$rep"
  echo; echo "## instruction mix / stalls"; python scripts/ncu_src.py /tmp/_src.csv
  echo; echo "## instructions and stall samples by source line (nvdisasm line info)"; python scripts/ncu_lines.py /tmp/_src.csv speechrecognitionproject_b200/libsrfe.so 30 || true
} > "$out"
