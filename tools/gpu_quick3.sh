#!/bin/bash
# kernel + whole-parse time of the index-writing shapes (8 GiB)
for cfg in "0 1 150" "1 1 150" "0 0 150"; do
  set -- $cfg
  echo "hist=$1 index=$2 L=$3:"; timeout 300 python tools/prof_one.py 8.0 $1 $2 $3 3 2>&1 | tail -1 | sed -e 's/Outcome(.*line_phase=0)//'
done
