mkdir -p gpurun_out/r01e
for w in 1 2 3 4; do
python tools/timeline.py --window $w --out gpurun_out/r01e/tl_w$w.json 2>&1 | grep -v "tiles cta"
done
