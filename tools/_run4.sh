set -x
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_reference_grids.py tests/test_gpu_inverse.py -m gpu -q > gpurun_out/r2_t4_grids.log 2>&1; echo "rc=$?" >> gpurun_out/r2_t4_grids.log
timeout 900 python -m pytest tests -m gpu -q --ignore tests/test_gpu_reference_grids.py --ignore tests/test_gpu_inverse.py > gpurun_out/r2_t4_full.log 2>&1; echo "rc=$?" >> gpurun_out/r2_t4_full.log
rm -f gpurun_out/r2_ab4.jsonl
timeout 300 python tools/ab_time.py --tag base cfg5 cfg4 >> gpurun_out/r2_ab4.jsonl 2>> gpurun_out/r2_ab4.err
timeout 120 python tools/cpu_scaling.py 1 8 16 32 64 128 > gpurun_out/r2_cpu_scaling.jsonl 2>&1
for th in 0 4 8 12 16 24; do
  B2L_H2D_THREADS=$th timeout 300 python - >> gpurun_out/r2_h2d_threads.txt 2>&1 <<PY
import os, time, sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np, bench, librosa_b200 as lb
w = bench.WORKLOADS["cfg2"]
ctx = lb.default_context()
lb.bind_host_to_device(0)
y = np.array(bench.make_batch(w, 0))
for _ in range(2): lb.feature.melspectrogram(y=y, sr=w["sr"], **w["kw"])
t0 = time.perf_counter()
for _ in range(4): out = lb.feature.melspectrogram(y=y, sr=w["sr"], **w["kw"])
ctx.synchronize()
print("threads", os.environ["B2L_H2D_THREADS"], "pageable e2e ms", (time.perf_counter() - t0) / 4 * 1e3, flush=True)
PY
done
timeout 600 python bench.py --steps 10 > gpurun_out/r2_bench4.json 2> gpurun_out/r2_bench4.err
tail -n 15 gpurun_out/r2_t4_grids.log; tail -n 4 gpurun_out/r2_t4_full.log; cat gpurun_out/r2_ab4.jsonl gpurun_out/r2_cpu_scaling.jsonl gpurun_out/r2_h2d_threads.txt
