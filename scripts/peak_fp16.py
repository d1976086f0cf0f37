"""cuBLAS fp16 / bf16 8192^3 throughput, burst (best of 10) and sustained (back to back for ~3 s): the tensor
roofline the fp16 retrieval kernel is up against on THIS box (MEASURED_PEAKS.json holds the bf16 numbers)."""
import json, sys, time
import torch
out = {}
for name, dt in (("fp16", torch.float16), ("bf16", torch.bfloat16)):
    a = torch.randn(8192, 8192, device="cuda").to(dt); b = torch.randn(8192, 8192, device="cuda").to(dt)
    for _ in range(3): a @ b
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); a @ b; e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    n = 0; e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); t0 = time.time()
    while time.time() - t0 < 3.0:
        for _ in range(50): a @ b
        n += 50
    e1.record(); torch.cuda.synchronize()
    fl = 2 * 8192 ** 3
    out[name] = {"burst_tflops": fl / best / 1e9, "sustained_tflops": fl * n / e0.elapsed_time(e1) / 1e9}
print(json.dumps(out))
