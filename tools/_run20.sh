set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python tools/prefilter_probe.py 10000000 768 > gpurun_out/prefilter_probe2.json 2> gpurun_out/prefilter_probe2.err; cat gpurun_out/prefilter_probe2.json; tail -3 gpurun_out/prefilter_probe2.err
timeout 600 python -m pytest tests/test_gpu_dense.py -q -m gpu -k "prefilter" > gpurun_out/t_r2_pf_final.log 2>&1; tail -3 gpurun_out/t_r2_pf_final.log
