set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_quant.py -x -q -m gpu -k "pq" > gpurun_out/t_r2_pq16.log 2>&1; tail -5 gpurun_out/t_r2_pq16.log
timeout 600 python bench.py --config c4 --steps 10 --warmup 3 > gpurun_out/bench_r2_c4_pq16.json 2> gpurun_out/bench_r2_c4_pq16.err; tail -3 gpurun_out/bench_r2_c4_pq16.err
python -c "import json; d=json.loads(open('gpurun_out/bench_r2_c4_pq16.json').read().strip().splitlines()[-1]); print('c4 pq16', d['value'], d['ms_per_step'], d['roofline']['avg_launch_ms'], d['roofline']['frac'], d['e2e']['value'], d['parity'])"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pq_scan16 --launch-skip 9 --launch-count 1 -o gpurun_out/ncu_pq16_r02 -f python bench.py --config c4 --rows 8000000 --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_pq16.log 2>&1; tail -2 gpurun_out/ncu_pq16.log
