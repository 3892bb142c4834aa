set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu > gpurun_out/t_r2_c.log 2>&1; echo "rc=$?" >> gpurun_out/t_r2_c.log
tail -25 gpurun_out/t_r2_c.log
timeout 600 python bench.py --config c4 --steps 5 --warmup 3 > gpurun_out/bench_r2_c4_pq4.json 2> gpurun_out/bench_r2_c4_pq4.err; tail -c 1200 gpurun_out/bench_r2_c4_pq4.json; tail -5 gpurun_out/bench_r2_c4_pq4.err
