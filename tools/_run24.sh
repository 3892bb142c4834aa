cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 100 python -m pytest tests/test_gpu_dense.py tests/test_gpu_multi_local.py -q -m gpu -k "prefilter or 1200001" -x > gpurun_out/t_r2_small_sample.log 2>&1; tail -4 gpurun_out/t_r2_small_sample.log
