set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu -x > gpurun_out/t_r2_f.log 2>&1; echo "rc=$?" >> gpurun_out/t_r2_f.log
tail -12 gpurun_out/t_r2_f.log
timeout 300 python bench.py --config c2 --steps 50 --warmup 5 --no-cpu > gpurun_out/bench_r2_c2_lastcta.json 2> gpurun_out/bench_r2_c2_lastcta.err; tail -c 900 gpurun_out/bench_r2_c2_lastcta.json; tail -3 gpurun_out/bench_r2_c2_lastcta.err
timeout 600 python tools/hnsw_probe.py 500000 768 8192 128 > gpurun_out/hnsw_probe_d.json 2> gpurun_out/hnsw_probe_d.err; cat gpurun_out/hnsw_probe_d.json; tail -3 gpurun_out/hnsw_probe_d.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02_bench_c2.csv python bench.py --config c2 --steps 3 --warmup 3 --no-cpu > gpurun_out/launches_c2.log 2>&1; tail -2 gpurun_out/launches_c2.log; grep -c "dense_f32_stream" gpurun_out/launches_r02_bench_c2.csv
