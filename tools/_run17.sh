set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dense.py -x -q -m gpu -k "prefilter" > gpurun_out/t_r2_pf8.log 2>&1; tail -3 gpurun_out/t_r2_pf8.log
for sr in 0 65536 262144; do
QB_SAMPLE_ROWS=$sr timeout 600 python bench.py --config c2 --steps 20 --warmup 5 --no-cpu > gpurun_out/bench_r2_c2_pf8_s$sr.json 2> gpurun_out/bench_r2_c2_pf8_s$sr.err; tail -3 gpurun_out/bench_r2_c2_pf8_s$sr.err
python -c "import json; d=json.loads(open('gpurun_out/bench_r2_c2_pf8_s$sr.json').read().strip().splitlines()[-1]); print('c2 q8 sample $sr', d['value'], d['ms_per_step'], d['roofline']['avg_launch_ms'], d['roofline'].get('hbm_frac_of_bytes_moved'), d['e2e']['value'], d['parity'])"
done
