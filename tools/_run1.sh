set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_hnsw_device.py tests/test_gpu_edges.py -x -q -m gpu > gpurun_out/t_r2_a.log 2>&1; echo "rc=$?" >> gpurun_out/t_r2_a.log
tail -30 gpurun_out/t_r2_a.log
timeout 600 python tools/hnsw_probe.py 200000 768 4096 128 > gpurun_out/hnsw_probe_a.json 2> gpurun_out/hnsw_probe_a.err; tail -3 gpurun_out/hnsw_probe_a.json gpurun_out/hnsw_probe_a.err
