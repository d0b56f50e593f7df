for v in variants/*/; do echo "== $v"; VRJ_LIBDIR=$PWD/$v python tools/profile_step.py --spp 16 --reps 3 | tail -1; done
