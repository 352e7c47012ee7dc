mkdir -p gpurun_out/v3d
python tools/timeline.py --window 4 --nosync 1 --out gpurun_out/v3d/tl_nosync.json 2>&1 | tail -18
python tools/timeline.py --window 2 8 --out gpurun_out/v3d/tl_w.json 2>&1 | grep -E "step_us|weight wait"
