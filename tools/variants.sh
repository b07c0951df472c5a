#!/bin/bash
# python tools/one_gesv.py runs for a list of option sets: tools/variants.sh n reps "opts1" "opts2" ...
n=$1; reps=$2; shift 2
for v in "$@"; do timeout -s KILL 180 python tools/one_gesv.py $n 0 $reps $v; done
