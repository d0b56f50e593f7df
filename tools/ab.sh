#!/bin/bash
# A/B of kernel variants on the bench step (C3, 1080p, depth 8): usage tools/ab.sh [spp] [reps]
SPP=${1:-64}; REPS=${2:-4}
run() { echo "== $1"; shift; env "$@" python tools/profile_step.py --spp $SPP --reps $REPS | tail -2; }
[ -d variants/r1 ] && run "round-1 library" VRJ_LIBDIR=$PWD/variants/r1
run "current, general material kernels, list-staged rays" VRJ_MATERIAL_MASK=15 VRJ_RECORDS=0
run "current, Lambertian-only kernels, list-staged rays" VRJ_RECORDS=0
run "current, general material kernels, ray records" VRJ_MATERIAL_MASK=15
run "current (Lambertian-only kernels + ray records)" VRJ_X=0
for t in 12 8 4; do run "current, refill threshold $t" VRJ_TUNE=$t,2,4,64,262144,2; done
