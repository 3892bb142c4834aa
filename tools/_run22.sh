set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29722 bench.py --gpus 2 --config c2 --steps 40 --warmup 5 --no-cpu > gpurun_out/bench_r2_n2_final.json 2> gpurun_out/bench_r2_n2_final.err; tail -c 1800 gpurun_out/bench_r2_n2_final.json; tail -5 gpurun_out/bench_r2_n2_final.err
