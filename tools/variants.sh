for v in variants/*/; do echo "== $v"; VRJ_LIBDIR=$PWD/$v python tools/profile_step.py --spp 8 --reps 2 | tail -1; done
