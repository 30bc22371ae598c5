#!/bin/bash
# box facts the bench sizing rule depends on (host cores / memory limits, GPU topology)
mkdir -p gpurun_out
{
echo "== nproc $(nproc)"; free -g | head -2
echo "== cgroup"; cat /sys/fs/cgroup/cpu.max /sys/fs/cgroup/memory.max 2>/dev/null
grep -E 'MemTotal|MemAvailable' /proc/meminfo
lscpu | grep -E 'Model name|Socket|Core|Thread|NUMA'
ulimit -l
nvidia-smi --query-gpu=index,name,memory.total --format=csv
nvidia-smi topo -m 2>/dev/null | head -20
python - <<'PY'
import torch, time, numpy as np, ctypes
n = 32 << 20
a = np.random.rand(n // 8)
d = torch.empty(n // 8, dtype=torch.float64, device="cuda")
t = torch.from_numpy(a)
torch.cuda.synchronize()
for rep in range(3):
    t0 = time.time(); d.copy_(t); torch.cuda.synchronize(); t1 = time.time()
    print("pageable H2D 32MB: %.2f ms" % ((t1 - t0) * 1e3))
rt = torch.cuda.cudart()
t0 = time.time(); r = rt.cudaHostRegister(a.ctypes.data, n, 0); t1 = time.time()
print("cudaHostRegister 32MB: %.2f ms rc=%s" % ((t1 - t0) * 1e3, r))
for rep in range(3):
    t0 = time.time(); d.copy_(t, non_blocking=True); torch.cuda.synchronize(); t1 = time.time()
    print("registered H2D 32MB: %.2f ms" % ((t1 - t0) * 1e3))
t0 = time.time(); rt.cudaHostUnregister(a.ctypes.data); t1 = time.time()
print("cudaHostUnregister 32MB: %.2f ms" % ((t1 - t0) * 1e3))
PY
} > gpurun_out/sysinfo.txt 2>&1
cat gpurun_out/sysinfo.txt
