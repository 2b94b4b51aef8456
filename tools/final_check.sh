timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2g_bench_ref.json 2> gpurun_out/r2g_bench.err
python bench.py --steps 10 --warmup 3 > gpurun_out/r2g_bench.json 2>> gpurun_out/r2g_bench.err
python - <<'P'
import json
d=json.load(open('gpurun_out/r2g_bench.json')); r=d['roofline']
print('value', round(d['value']), 'e2e', round(d['e2e']['value']), 'pageable', round(d['e2e']['pageable']['value']), 'frac', round(r['frac'],3), 'alone', round(r['alone']['frac'],3), 'launches', d['gpu_launches'])
print('parity', d['parity']); print('cpu', d['cpu_baseline']['value'], d['cpu_baseline']['cores']); print('clocks', d['clocks'])
for k,v in d['extras'].items(): print(k, v.get('value', v.get('searches_per_sec')), v.get('e2e'), (v.get('roofline') or {}).get('frac'), v.get('parity'), v.get('error'))
print('ref', json.load(open('gpurun_out/r2g_bench_ref.json'))['value'])
P
tail -3 gpurun_out/r2g_bench.err
for b in 6 48; do python tools/bench_kernels.py vitl $b 2>&1 | grep "^vit "; done
python bench.py --workload "configs[2]" --steps 10 --warmup 3 > gpurun_out/r2g_bench_configs2.json 2>> gpurun_out/r2g_bench.err; python -c "
import json; d=json.load(open('gpurun_out/r2g_bench_configs2.json')); print('configs[2]', d['metric'], round(d['value'],1), d['unit'], 'e2e', d.get('e2e',{}).get('value'), 'roofline', d.get('roofline'))"
