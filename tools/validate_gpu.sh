timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for tool in memcheck racecheck; do
  timeout 600 compute-sanitizer --tool $tool --error-exitcode 1 python tools/sanitize_smoke.py > gpurun_out/r2q_sanitize_$tool.txt 2>&1; echo "$tool rc=$?" >> gpurun_out/r2q_sanitize_$tool.txt
  tail -4 gpurun_out/r2q_sanitize_$tool.txt
done
for w in "configs[2]" "configs[3]" "configs[0]"; do python bench.py --workload "$w" > "gpurun_out/r2q_bench_$w.json" 2> gpurun_out/r2q_bench.err; python -c "
import json,sys
d=json.load(open('gpurun_out/r2q_bench_$w.json')); print('$w', d['value'], d['e2e']['value'], (d.get('roofline') or {}).get('frac'), d.get('parity'), d.get('vit_forward_ms'))"; done
