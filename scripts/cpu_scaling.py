import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
print("cpu_count", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
try: print("cgroup cpu.max:", open("/sys/fs/cgroup/cpu.max").read().strip())
except Exception as e: print("cpu.max n/a", e)
for th in (1, 4, 16, 32, 64, 128):
    pool = bench.OraclePool("kuka", 1024, 32, th)
    pool.step()
    dt = min(pool.step() for _ in range(2))
    print("threads %3d: %.0f env-steps/s  (%.1f us/env-step/thread)" % (pool.threads, 1024 * 32 / dt, dt * pool.threads / (1024 * 32) * 1e6))
