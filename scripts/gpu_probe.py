"""First-contact probe of the GPU box: host cores, faiss presence,
topology, TF32 / bf16 cuBLAS peaks measured the way MEASURED_PEAKS.json was."""
import json
import os
import subprocess
import time

import torch

out = {}
out["nproc"] = os.cpu_count()
out["affinity"] = len(os.sched_getaffinity(0))
try:
    import faiss  # noqa: F401
    out["faiss"] = getattr(faiss, "__version__", "present")
except Exception as e:  # noqa: BLE001
    out["faiss"] = f"absent ({type(e).__name__})"
out["baseline_ref"] = os.path.isdir("baseline/_ref")
out["gpu"] = torch.cuda.get_device_name(0)
out["sm_count"] = torch.cuda.get_device_properties(0).multi_processor_count
try:
    out["lscpu"] = subprocess.run("lscpu | grep -E 'Model name|^CPU\\(s\\)|Thread|Socket'", shell=True,
                                  capture_output=True, text=True).stdout.strip().split("\n")
except Exception:  # noqa: BLE001
    pass


def peak(dtype, tf32):
    torch.backends.cuda.matmul.allow_tf32 = tf32
    n = 8192
    a = torch.randn(n, n, device="cuda", dtype=dtype)
    b = torch.randn(n, n, device="cuda", dtype=dtype)
    for _ in range(3):
        a @ b
    best = 1e9
    for _ in range(10):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        a @ b
        e.record()
        torch.cuda.synchronize()
        best = min(best, s.elapsed_time(e))
    burst = 2 * n ** 3 / best / 1e9
    t0 = time.time()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    it = 0
    while time.time() - t0 < 3.0:
        for _ in range(20):
            a @ b
        it += 20
        torch.cuda.synchronize()
    e.record()
    torch.cuda.synchronize()
    sus = 2 * n ** 3 * it / s.elapsed_time(e) / 1e9
    return round(burst, 1), round(sus, 1)


out["tf32_tflops_burst_sustained"] = peak(torch.float32, True)
out["bf16_tflops_burst_sustained"] = peak(torch.bfloat16, False)
out["fp32_tflops_burst_sustained"] = peak(torch.float32, False)
print(json.dumps(out, indent=1))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/probe.json", "w"), indent=1)
