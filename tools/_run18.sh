set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -q -m gpu ) > gpurun_out/t_r2_final.log 2>&1; tail -6 gpurun_out/t_r2_final.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_r2_final.log 2>&1; tail -2 gpurun_out/smoke_r2_final.log
( time timeout 900 python bench.py --steps 20 --warmup 5 ) > gpurun_out/bench_r2_final.json 2> gpurun_out/bench_r2_final.err; tail -c 300 gpurun_out/bench_r2_final.json; tail -6 gpurun_out/bench_r2_final.err
( time timeout 300 python bench.py --impl reference --steps 20 --warmup 5 ) > gpurun_out/bench_r2_final_ref.json 2> gpurun_out/bench_r2_final_ref.err; tail -c 300 gpurun_out/bench_r2_final_ref.json
timeout 400 ncu --set full --clock-control none --import-source on -k regex:dense_q8_filter --launch-skip 4 --launch-count 1 -o gpurun_out/ncu_q8_filter_r02 -f python bench.py --config c2 --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_q8.log 2>&1; tail -2 gpurun_out/ncu_q8.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sq8_mma_kernel --launch-skip 5 --launch-count 1 -o gpurun_out/ncu_f32mma_r02 -f python tools/f32_batch_probe.py 4000000 1024 > gpurun_out/ncu_f32mma.log 2>&1; tail -2 gpurun_out/ncu_f32mma.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:hnsw_search --launch-skip 8 --launch-count 1 -o gpurun_out/ncu_hnsw128_r02 -f python tools/hnsw_probe.py 200000 768 4096 128 > gpurun_out/ncu_hnsw128.log 2>&1; tail -2 gpurun_out/ncu_hnsw128.log
ls -la gpurun_out/*.ncu-rep
