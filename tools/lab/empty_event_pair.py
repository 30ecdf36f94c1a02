import torch, statistics
x = torch.zeros(32, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ts = []
for i in range(30):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); x.add_(1.0); e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) * 1e3)
print("event pair around a 32-element add kernel after an L2 flush: median %.2f us, min %.2f us" % (statistics.median(ts[5:]), min(ts[5:])))
ts = []
for i in range(30):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) * 1e3)
print("event pair around nothing: median %.2f us" % statistics.median(ts[5:]))
