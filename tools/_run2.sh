set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/t_r2_b.log 2>&1; echo "rc=$?" >> gpurun_out/t_r2_b.log
tail -15 gpurun_out/t_r2_b.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_r2.log 2>&1; tail -3 gpurun_out/smoke_r2.log
( time timeout 1200 python bench.py --steps 20 --warmup 5 ) > gpurun_out/bench_r2_a.json 2> gpurun_out/bench_r2_a.err; tail -c 1500 gpurun_out/bench_r2_a.json; tail -8 gpurun_out/bench_r2_a.err
( time timeout 300 python bench.py --impl reference --steps 20 --warmup 5 ) > gpurun_out/bench_r2_ref.json 2> gpurun_out/bench_r2_ref.err; tail -c 800 gpurun_out/bench_r2_ref.json; tail -4 gpurun_out/bench_r2_ref.err
