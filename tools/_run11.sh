set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_quant.py -x -q -m gpu -k "pq" > gpurun_out/t_r2_pq8.log 2>&1; tail -5 gpurun_out/t_r2_pq8.log
timeout 600 python -m pytest tests/test_gpu_multi_local.py -x -q -m gpu > gpurun_out/t_r2_pipe.log 2>&1; tail -5 gpurun_out/t_r2_pipe.log
timeout 600 python bench.py --config c4 --steps 10 --warmup 3 > gpurun_out/bench_r2_c4_pq8.json 2> gpurun_out/bench_r2_c4_pq8.err; tail -3 gpurun_out/bench_r2_c4_pq8.err
python -c "import json; d=json.loads(open('gpurun_out/bench_r2_c4_pq8.json').read().strip().splitlines()[-1]); print('c4 pq8', d['value'], d['ms_per_step'], d['roofline']['avg_launch_ms'], d['roofline']['frac'], d['parity'])"
QB_PQ_QUERIES=4 timeout 600 python bench.py --config c4 --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_r2_c4_pq4b.json 2> gpurun_out/bench_r2_c4_pq4b.err
python -c "import json; d=json.loads(open('gpurun_out/bench_r2_c4_pq4b.json').read().strip().splitlines()[-1]); print('c4 pq4', d['value'], d['ms_per_step'], d['roofline']['avg_launch_ms'], d['roofline']['frac'])"
