"""Measured L2 read bandwidth of the box: torch.sum over an L2-resident float32 buffer (32 MB and 64 MB of the 126 MB L2),
CUDA events, best of 20 after warm-up.  The denominator of bench.py's roofline_l2."""
import torch
for mb in (16, 32, 64, 96):
    x = torch.ones(mb * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")
    for _ in range(5): x.sum()
    best = 1e9
    for _ in range(20):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); x.sum(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print(f"{mb} MB resident buffer: {best*1e3:.1f} us  {mb*1.048576/best:.1f} GB/s")
