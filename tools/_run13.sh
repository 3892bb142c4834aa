set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 300 python -m pytest tests/test_gpu_multi_local.py -x -q -m gpu > gpurun_out/t_r2_pipe8.log 2>&1; tail -3 gpurun_out/t_r2_pipe8.log
for n in 8 2; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2961$n bench.py --gpus $n --config c2 --steps 40 --warmup 5 > gpurun_out/bench_r2_pipe_n$n.json 2> gpurun_out/bench_r2_pipe_n$n.err; tail -c 1500 gpurun_out/bench_r2_pipe_n$n.json; tail -5 gpurun_out/bench_r2_pipe_n$n.err
done
