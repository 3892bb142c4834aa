set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_multi_local.py -q -m gpu > gpurun_out/t_r2_multi.log 2>&1; echo "rc=$?" >> gpurun_out/t_r2_multi.log; tail -8 gpurun_out/t_r2_multi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_r2_n2.json 2> gpurun_out/bench_r2_n2.err; tail -c 1500 gpurun_out/bench_r2_n2.json; tail -5 gpurun_out/bench_r2_n2.err
