"""How the CPU reference arm scales with the number of worker processes on this box (cgroup quotas show up as a
plateau far below the visible core count).  python tools/cpu_scaling.py [workers ...]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import multiprocessing as mp
import bench

w = bench.WORKLOADS["cfg2"]
counts = [int(a) for a in sys.argv[1:]] or [1, 4, 8, 16, 32, 64, 128]
print(json.dumps({"visible": len(os.sched_getaffinity(0)), "cpu_count": os.cpu_count(), "quota": bench.cpu_quota(),
                  "cpu.max": open("/sys/fs/cgroup/cpu.max").read().strip() if os.path.exists("/sys/fs/cgroup/cpu.max") else None}))
bench._CPU_BATCH = bench.make_batch(dict(w, clips=512), rank=0)
for n in counts:
    if n > len(os.sched_getaffinity(0)):
        continue
    pool = mp.get_context("fork").Pool(n, initializer=bench._cpu_worker_init)
    jobs = [(w["op"], w["kw"], w["sr"], i, i + 1) for i in range(max(64, 4 * n))]
    pool.map(bench._cpu_range_job, jobs[:n], chunksize=1)
    t0 = time.perf_counter()
    frames = sum(pool.map(bench._cpu_range_job, jobs, chunksize=1))
    dt = time.perf_counter() - t0
    pool.close(); pool.join()
    print(json.dumps({"workers": n, "frames_s": round(frames / dt), "per_worker": round(frames / dt / n)}), flush=True)
