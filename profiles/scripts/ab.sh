#!/bin/bash
# usage: scratch/ab.sh "ENV1=..;ENV2=.." ...   -> one bench line summary per env setting
for cfg in "$@"; do
  echo "== $cfg"
  env $(echo "$cfg" | tr ';' ' ') timeout 300 python bench.py --no-gpu-reference --steps 30 --warmup 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['share_of_step'])"
done
