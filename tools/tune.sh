for t in 0 16384 65536 131072 262144 524288 1048576 4194304; do echo "tail_max=$t"; VRJ_TUNE=16,2,4,64,$t python tools/profile_step.py --spp 16 --reps 3 | tail -1; done
echo C5; for t in 0 131072 1048576; do VRJ_TUNE=16,2,4,64,$t python tools/run_config.py C5 --reps 2 | tail -1 | cut -c1-330; done
