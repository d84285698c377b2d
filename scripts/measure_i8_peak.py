"""int8 tensor-core peak measured the way MEASURED_PEAKS.json measures bf16: a library GEMM
(torch._int_mm -> cuBLASLt s8 x s8 -> s32) at 8192^3, best of 10 (burst) and back to back for ~3 s (sustained),
CUDA events; fp16 / bf16 torch.matmul beside it on the same box.  Writes gpurun_out/i8_peak.json."""
import json
import os
import time

import torch


def bench(fn, flops, burst_reps=10, sustain_s=3.0):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(burst_reps):
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    n = max(10, int(sustain_s * 1e3 / best))
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    sus = e0.elapsed_time(e1) / n
    return round(flops / best / 1e9, 1), round(flops / sus / 1e9, 1)


def main():
    dev = "cuda"
    n = 8192
    out = {"gpu": torch.cuda.get_device_name(0), "n": n, "when": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime())}
    a8 = torch.randint(-128, 127, (n, n), dtype=torch.int8, device=dev)
    b8 = torch.randint(-128, 127, (n, n), dtype=torch.int8, device=dev).t()   # column-major B as cuBLASLt wants
    try:
        out["int8_tops"], out["int8_tops_sustained"] = bench(lambda: torch._int_mm(a8, b8), 2.0 * n ** 3)
    except Exception as e:  # noqa: BLE001
        out["int8_error"] = f"{type(e).__name__}: {e}"[:300]
    for name, dt in (("fp16", torch.float16), ("bf16", torch.bfloat16)):
        a = torch.randn(n, n, dtype=dt, device=dev)
        b = torch.randn(n, n, dtype=dt, device=dev)
        out[f"{name}_tflops"], out[f"{name}_tflops_sustained"] = bench(lambda: torch.matmul(a, b), 2.0 * n ** 3)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/i8_peak.json", "w"), indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
