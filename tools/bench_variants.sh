mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 --mixing swap-neighbors 2>&1 | tail -1 > gpurun_out/bench_r1_n1_swapneighbors.json
python bench.py --steps 20 --warmup 5 --replicas 64 2>&1 | tail -1 > gpurun_out/bench_r1_n1_k64.json
python -m pytest tests/test_gpu_sampler.py -m gpu -q -k locality 2>&1 | tail -2
python -c "
import json
for f in ('gpurun_out/bench_r1_n1_swapneighbors.json','gpurun_out/bench_r1_n1_k64.json'):
    d=json.load(open(f)); print(f, d['value'], d['phases_ms'], d['e2e']['value'], d['cpu_baseline']['value'], d['cpu_baseline']['phases_ms'], d['roofline']['frac'])
"
