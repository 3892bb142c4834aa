set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_custom.py tests/test_gpu_hnsw_device.py tests/test_gpu_dense.py tests/test_gpu_formats.py tests/test_gpu_train.py tests/test_gpu_maxsim.py tests/test_gpu_edges.py tests/test_gpu_quant.py tests/test_gpu_multi_local.py -q -m gpu > gpurun_out/t_r2_d.log 2>&1; echo "rc=$?" >> gpurun_out/t_r2_d.log
tail -8 gpurun_out/t_r2_d.log
timeout 900 python tools/f32_batch_probe.py 10000000 1024 > gpurun_out/f32_batch_probe_a.json 2> gpurun_out/f32_batch_probe_a.err; cat gpurun_out/f32_batch_probe_a.json; tail -3 gpurun_out/f32_batch_probe_a.err
timeout 600 python tools/hnsw_probe.py 200000 768 4096 128 > gpurun_out/hnsw_probe_b.json 2> gpurun_out/hnsw_probe_b.err; cat gpurun_out/hnsw_probe_b.json; tail -3 gpurun_out/hnsw_probe_b.err
for q in 4 2; do QB_PQ_QUERIES=$q timeout 300 python bench.py --config c4 --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_r2_c4_qpp$q.json 2> gpurun_out/bench_r2_c4_qpp$q.err; python -c "import json,sys; d=json.loads(open('gpurun_out/bench_r2_c4_qpp$q.json').read().strip().splitlines()[-1]); print('c4 qpp=$q', d['value'], d['roofline']['avg_launch_ms'], d['roofline']['frac'])"; tail -2 gpurun_out/bench_r2_c4_qpp$q.err; done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pq_scan4 --launch-skip 14 --launch-count 1 -o gpurun_out/ncu_pq4_r02 -f python bench.py --config c4 --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_pq4.log 2>&1; tail -3 gpurun_out/ncu_pq4.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hnsw_search --launch-skip 1 --launch-count 1 -o gpurun_out/ncu_hnsw_r02 -f python tools/hnsw_probe.py 100000 768 2048 128 > gpurun_out/ncu_hnsw.log 2>&1; tail -3 gpurun_out/ncu_hnsw.log
ls -la gpurun_out/*.ncu-rep
