. tools/gpu_ab.sh
timeout 900 python -m pytest tests -m gpu -x -q -k "golden or full_grids or c2_frame or c1_real or layouts or thin or c5" 2>&1 | tail -3
run trim c3 8 X=1
run trim c2 64 X=1
