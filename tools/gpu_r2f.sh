mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_b_tc.py -q -m gpu -s -k "split_operand or stream_schedules" > gpurun_out/r2f_split.log 2>&1; echo "split exit $?"
grep -E "passed|failed|fwd" gpurun_out/r2f_split.log | cut -c1-200 | tail -12
timeout 900 python -m pytest tests/test_gpu_c_parity.py tests/test_gpu_e_training.py -q -m gpu -s -k "f32tc" > gpurun_out/r2f_parity.log 2>&1; echo "parity exit $?"
grep -E "passed|failed|^E  .*assert|FAILED" gpurun_out/r2f_parity.log | cut -c1-250 | tail -20
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2f_bench.json') if l.startswith('{')][-1])
print(d['value'], d['e2e']['value']); print(d.get('parity_mode')); print(d['cpu_baseline']['parity'])
print({k:v for k,v in d['configs']['config3_embedding_inference'].items() if k!='workload'})
PY
