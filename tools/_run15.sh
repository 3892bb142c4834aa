set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dense.py -x -q -m gpu -k "prefilter" > gpurun_out/t_r2_pf.log 2>&1; tail -5 gpurun_out/t_r2_pf.log
timeout 600 python bench.py --config c2 --steps 20 --warmup 5 --no-cpu > gpurun_out/bench_r2_c2_pf.json 2> gpurun_out/bench_r2_c2_pf.err; tail -3 gpurun_out/bench_r2_c2_pf.err
python -c "import json; d=json.loads(open('gpurun_out/bench_r2_c2_pf.json').read().strip().splitlines()[-1]); print('c2 pf', d['value'], d['ms_per_step'], d['roofline']['avg_launch_ms'], d['roofline']['frac'], d['roofline'].get('hbm_frac_of_bytes_moved'), d['e2e']['value'], d['gpu_launches'], d['parity'])"
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_r2c.log 2>&1; tail -2 gpurun_out/smoke_r2c.log
