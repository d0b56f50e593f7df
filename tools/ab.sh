#!/bin/bash
# A/B of kernel variants on the bench step (C3, 1080p, depth 8): usage tools/ab.sh [spp] [reps] [variant dirs...]
SPP=${1:-64}; REPS=${2:-4}; shift; shift
run() { echo "== $1"; shift; env "$@" python tools/profile_step.py --spp $SPP --reps $REPS | tail -2; }
run "current build" VRJ_X=0
run "current build, list-staged rays" VRJ_RECORDS=0
for v in "$@"; do
  [ -d variants/$v ] && run "variants/$v" VRJ_LIBDIR=$PWD/variants/$v
  [ -d variants/$v ] && run "variants/$v, list-staged rays" VRJ_LIBDIR=$PWD/variants/$v VRJ_RECORDS=0
done
