"""Reference points for HBM-bound kernels on this box: pure write (fill), pure read (sum), copy.  CUDA events, best of 10."""
import torch
n = 1200 * 1024 * 1024 // 4
a = torch.empty(n, dtype=torch.float32, device="cuda")
b = torch.empty(n, dtype=torch.float32, device="cuda")
def best(fn, reps=10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fn(); torch.cuda.synchronize()
    t = 1e9
    for _ in range(reps):
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        t = min(t, e0.elapsed_time(e1))
    return t
bytes_ = n * 4
print("fill  (write only): %.1f GB/s" % (bytes_ / best(lambda: a.fill_(1.0)) / 1e6))
print("sum   (read only) : %.1f GB/s" % (bytes_ / best(lambda: a.sum()) / 1e6))
print("copy  (read+write): %.1f GB/s" % (2 * bytes_ / best(lambda: b.copy_(a)) / 1e6))
