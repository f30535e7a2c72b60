#!/bin/bash
# development aid: per-stage device times for a list of "VAR=val,VAR=val" environment settings, one box:
#   gpurun -- 'tools/variants.sh X=0 TSP_SOME_KNOB=1 > /dev/null; cat gpurun_out/variants.log'
out=gpurun_out/variants.log
: > $out
for v in "$@"; do
  echo "== $v" >> $out
  env $(echo $v | tr ',' ' ') python bench.py --device-only --steps 20 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['single_stream_ms_per_step'], d['stage_ms'])" >> $out
done
cat $out
