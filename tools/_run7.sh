set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L | wc -l
for n in 8 4; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/bench_r2_n$n.json 2> gpurun_out/bench_r2_n$n.err; tail -c 1200 gpurun_out/bench_r2_n$n.json; tail -5 gpurun_out/bench_r2_n$n.err
done
