set -x
mkdir -p gpurun_out
timeout 200 python scripts/ncu_constrained.py 1e8 2>&1 | tail -2 | cut -c1-300
COLIBRI_B200_MATCH_BPS=8 timeout 200 python scripts/ncu_constrained.py 1e8 2>&1 | tail -1 | cut -c1-300
COLIBRI_B200_MATCH_BPS=16 timeout 200 python scripts/ncu_constrained.py 1e8 2>&1 | tail -1 | cut -c1-300
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"constrained_match" -s 4 -c 4 -f -o gpurun_out/prof_constrained_r01c python scripts/ncu_constrained.py > gpurun_out/ncu_constrained2.log 2>&1
