set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -q -m gpu ) > gpurun_out/t_r2_final2.log 2>&1; tail -6 gpurun_out/t_r2_final2.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_r2_final2.log 2>&1; tail -2 gpurun_out/smoke_r2_final2.log
( time timeout 900 python bench.py --steps 20 --warmup 5 ) > gpurun_out/bench_r2_final2.json 2> gpurun_out/bench_r2_final2.err; tail -c 300 gpurun_out/bench_r2_final2.json; tail -6 gpurun_out/bench_r2_final2.err
