"""Experiment: HBM read bandwidth for row segments of w columns out of a [n, ld] float32 matrix (torch copy kernels)."""
import torch, time
n, ld = 700_000, 18016
X = torch.empty((n, ld), dtype=torch.float32, device="cuda").normal_()
def bw(fn, nbytes, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return nbytes * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9
for w in (32, 64, 128, 256, 512, 1024, 4096):
    nslab = max(1, 4096 // w)
    outs = [torch.empty((n, w), dtype=torch.float32, device="cuda") for _ in range(nslab)]
    def f():
        for s in range(nslab):
            c0 = (s * 37 * w) % (ld - w)
            c0 -= c0 % 32
            torch.sum(X[:, c0:c0 + w], dim=0, out=outs[s][0])
    print("segment %5d B: read %.0f GB/s (column-sum of %d slabs)" % (w * 4, bw(f, n * w * 4 * nslab), nslab), flush=True)
flat = X.view(-1)[: n * 4096]
print("contiguous: %.0f GB/s" % bw(lambda: torch.sum(flat.view(n, 4096), dim=0), n * 4096 * 4))
