mkdir -p gpurun_out/r01c
python tools/timeline.py --window 6 --out gpurun_out/r01c/timeline_sync.json > gpurun_out/r01c/timeline_sync.txt 2>&1
python tools/timeline.py --window 6 --nosync 1 --out gpurun_out/r01c/timeline_nosync.json > gpurun_out/r01c/timeline_nosync.txt 2>&1
cat gpurun_out/r01c/timeline_sync.txt gpurun_out/r01c/timeline_nosync.txt
