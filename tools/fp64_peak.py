"""Measured FP64 peak of the device (tool; SURVEY §8d: the FP64 roofline denominator must be a
measurement, MEASURED_PEAKS.json has none).  ``measure()`` returns TFLOP/s of dependent-FMA chains
(2 flops per instruction) and instructions/s of the non-fused multiply+add form, best of ``reps``
launches timed with CUDA events, plus the SM clock seen by NVML right after.

    python tools/fp64_peak.py            # prints one JSON line
"""
from __future__ import annotations

import ctypes
import json
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "libprobe.so")
SRC = os.path.join(HERE, "fp64_peak.cu")


def build(force: bool = False) -> str:
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        subprocess.run([os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc"), "-gencode", "arch=compute_100a,code=sm_100a",
                        "-O3", "-lineinfo", "-Xcompiler", "-fPIC", "-shared", "-o", LIB, SRC], check=True,
                       capture_output=True)
    return LIB


def measure(reps: int = 5, iters: int = 20000) -> dict:
    import torch
    lib = ctypes.CDLL(build())
    lib.probe_fp64.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
    chains = lib.probe_chains()
    dev = torch.device("cuda", torch.cuda.current_device())
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    out = torch.zeros(8, dtype=torch.float64, device=dev)
    blocks, threads = sms * 8, 256
    stream = torch.cuda.current_stream().cuda_stream
    result = {}
    for kind, name in ((0, "fma"), (1, "mul_add")):
        best = None
        for rep in range(reps + 1):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            rc = lib.probe_fp64(kind, blocks, threads, iters, out.data_ptr(), stream)
            e1.record()
            torch.cuda.synchronize()
            assert rc == 0, rc
            ms = e0.elapsed_time(e1)
            if rep and (best is None or ms < best):
                best = ms
        insts = blocks * threads * chains * iters * (1 if kind == 0 else 2)
        result[name] = {"ms": best, "thread_instructions_per_s": insts / (best * 1e-3)}
    result["fp64_fma_tflops"] = 2 * result["fma"]["thread_instructions_per_s"] / 1e12
    # separately rounded multiply and add: 1 flop per instruction
    result["fp64_mul_add_tflops"] = result["mul_add"]["thread_instructions_per_s"] / 1e12
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(torch.cuda.current_device())
        result["sm_mhz_after"] = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
        result["sm_max_mhz"] = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
    except Exception:
        pass
    result["how"] = ("%d blocks x %d threads x %d independent chains x %d iterations, best of %d, CUDA events"
                     % (blocks, threads, chains, iters, reps))
    return result


if __name__ == "__main__":
    print(json.dumps(measure()))
