set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for v in "" _epi8 _epi12 _pf6; do
  QB_LIB_PATH=$GRAFT_REPO_ROOT/qdrant_b200/lib/libqdrant_b200$v.so timeout 300 python bench.py --config c3 --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_r2_c3$v.json 2> gpurun_out/bench_r2_c3$v.err
  python -c "import json; d=json.loads(open('gpurun_out/bench_r2_c3$v.json').read().strip().splitlines()[-1]); print('c3 variant [$v]', d['value'], d['ms_per_step'], d['roofline']['avg_launch_ms'], d['roofline']['frac'])"; tail -2 gpurun_out/bench_r2_c3$v.err
done
for v in "" _epi8 _epi12; do
  QB_LIB_PATH=$GRAFT_REPO_ROOT/qdrant_b200/lib/libqdrant_b200$v.so timeout 600 python tools/f32_batch_probe.py 10000000 1024 > gpurun_out/f32_batch_probe$v.json 2> gpurun_out/f32_batch_probe$v.err; cat gpurun_out/f32_batch_probe$v.json
done
for v in "" _nolnk; do
  QB_LIB_PATH=$GRAFT_REPO_ROOT/qdrant_b200/lib/libqdrant_b200$v.so timeout 600 python tools/hnsw_probe.py 500000 768 8192 128 > gpurun_out/hnsw_probe_e$v.json 2> gpurun_out/hnsw_probe_e$v.err; cat gpurun_out/hnsw_probe_e$v.json
done
for c in 64 128 256; do
  QB_MMA_SEG_CAP=$c timeout 300 python bench.py --config c3 --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_r2_c3_seg$c.json 2> gpurun_out/bench_r2_c3_seg$c.err
  python -c "import json; d=json.loads(open('gpurun_out/bench_r2_c3_seg$c.json').read().strip().splitlines()[-1]); print('c3 seg_cap $c', d['value'], d['ms_per_step'], d['roofline']['avg_launch_ms'], d['roofline']['frac'], d['parity'])"; tail -2 gpurun_out/bench_r2_c3_seg$c.err
done
