set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dense.py -x -q -m gpu -k "prefilter" > gpurun_out/t_r2_pf8.log 2>&1; tail -5 gpurun_out/t_r2_pf8.log
timeout 600 python bench.py --config c2 --steps 20 --warmup 5 --no-cpu > gpurun_out/bench_r2_c2_pf8.json 2> gpurun_out/bench_r2_c2_pf8.err; tail -3 gpurun_out/bench_r2_c2_pf8.err
python -c "import json; d=json.loads(open('gpurun_out/bench_r2_c2_pf8.json').read().strip().splitlines()[-1]); print('c2 q8', d['value'], d['ms_per_step'], d['roofline']['avg_launch_ms'], d['roofline']['frac'], d['roofline'].get('hbm_frac_of_bytes_moved'), d['e2e']['value'], d['gpu_launches'], d['parity'])"
QB_PREFILTER_PLANE=1 timeout 600 python bench.py --config c2 --steps 20 --warmup 5 --no-cpu > gpurun_out/bench_r2_c2_pf16.json 2> gpurun_out/bench_r2_c2_pf16.err; tail -3 gpurun_out/bench_r2_c2_pf16.err
python -c "import json; d=json.loads(open('gpurun_out/bench_r2_c2_pf16.json').read().strip().splitlines()[-1]); print('c2 bf16', d['value'], d['ms_per_step'], d['roofline']['avg_launch_ms'], d['roofline']['frac'], d['e2e']['value'], d['gpu_launches'], d['parity'])"
