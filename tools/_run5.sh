set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu > gpurun_out/t_r2_e.log 2>&1; echo "rc=$?" >> gpurun_out/t_r2_e.log
tail -12 gpurun_out/t_r2_e.log
timeout 600 python tools/hnsw_probe.py 500000 768 8192 128 > gpurun_out/hnsw_probe_c.json 2> gpurun_out/hnsw_probe_c.err; cat gpurun_out/hnsw_probe_c.json; tail -3 gpurun_out/hnsw_probe_c.err
for q in 4 2; do QB_PQ_QUERIES=$q timeout 300 python bench.py --config c4 --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_r2_c4_qpp$q.json 2> gpurun_out/bench_r2_c4_qpp$q.err; python -c "import json,sys; d=json.loads(open('gpurun_out/bench_r2_c4_qpp$q.json').read().strip().splitlines()[-1]); print('c4 qpp=$q', d['value'], d['roofline']['avg_launch_ms'], d['roofline']['frac'])"; tail -2 gpurun_out/bench_r2_c4_qpp$q.err; done
timeout 900 python tools/f32_batch_probe.py 10000000 1024 > gpurun_out/f32_batch_probe_b.json 2> gpurun_out/f32_batch_probe_b.err; cat gpurun_out/f32_batch_probe_b.json
( time timeout 1500 python bench.py --steps 20 --warmup 5 ) > gpurun_out/bench_r2_b.json 2> gpurun_out/bench_r2_b.err; tail -c 600 gpurun_out/bench_r2_b.json; tail -6 gpurun_out/bench_r2_b.err
