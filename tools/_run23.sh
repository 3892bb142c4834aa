cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 90 python -m pytest tests/test_gpu_multi_local.py -q -m gpu -k "1200001" > gpurun_out/t_r2_multi_big.log 2>&1; tail -4 gpurun_out/t_r2_multi_big.log
