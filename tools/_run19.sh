set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python tools/prefilter_probe.py 10000000 768 > gpurun_out/prefilter_probe.json 2> gpurun_out/prefilter_probe.err; cat gpurun_out/prefilter_probe.json; tail -3 gpurun_out/prefilter_probe.err
timeout 300 python -m pytest tests/test_gpu_multi_local.py tests/test_gpu_multi.py -q -m gpu > gpurun_out/t_r2_multi_final.log 2>&1; tail -3 gpurun_out/t_r2_multi_final.log
