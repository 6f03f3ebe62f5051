"""Probe of the host side of a multi-GPU box (development tool): CPU / NUMA affinity of every GPU (NVML), the CPUs this
process may run on, and the aggregate pinned device->host bandwidth of all GPUs copying at once, with and without binding
each copying process to its GPU's CPUs.  usage: python tools/numa_probe.py"""
import os
import subprocess
import sys
import time


def worker(dev, bind, secs=2.0):
    import torch
    if bind:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(dev)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * w + b for w, v in enumerate(words) for b in range(64) if (v >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
    torch.cuda.set_device(dev)
    n = 1 << 29
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    h_ = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_.copy_(d, non_blocking=True)
    torch.cuda.synchronize()
    # crude start barrier: wait for the next multiple of 5 s of wall time
    time.sleep(5.0 - time.time() % 5.0)
    t0 = time.perf_counter()
    reps = 0
    while time.perf_counter() - t0 < secs:
        h_.copy_(d, non_blocking=True)
        torch.cuda.synchronize()
        reps += 1
    dt = time.perf_counter() - t0
    print(f"dev {dev} bind={bind}: {reps * n / dt / 1e9:.1f} GB/s", flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1:
        worker(int(sys.argv[1]), sys.argv[2] == "1")
        sys.exit(0)
    print("cpus allowed:", sorted(os.sched_getaffinity(0)), "of", os.cpu_count())
    print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout)
    try:
        print(open("/sys/devices/system/node/online").read().strip(), "NUMA nodes online")
    except OSError:
        pass
    import torch
    ng = torch.cuda.device_count()
    for bind in ("0", "1"):
        ps = [subprocess.Popen([sys.executable, __file__, str(g), bind]) for g in range(ng)]
        for p in ps:
            p.wait()
