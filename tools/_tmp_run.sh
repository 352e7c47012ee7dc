mkdir -p gpurun_out/tc5
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/tc5/bench.json 2> gpurun_out/tc5/bench.err; tail -3 gpurun_out/tc5/bench.err; python -c "
import json; d=json.load(open('gpurun_out/tc5/bench.json')); print(d['value'], d['ms_per_step'], d['decode_ms_per_token'], d['first_chunk_ms'], d['roofline']['frac'])"
GENVC_TC=0 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/tc5/bench_notc.json 2> gpurun_out/tc5/bench.err; python -c "
import json; d=json.load(open('gpurun_out/tc5/bench_notc.json')); print('no-tc', d['value'], d['ms_per_step'], d['decode_ms_per_token'], d['first_chunk_ms'], d['roofline']['frac'])"
