"""HBM streaming rate of a [n, ld] float32 matrix read as column slabs (segment = bytes contiguous per row)."""
import ctypes, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from chrysalis_b200 import _lib
lib = _lib.load()
f = lib.chr_debug_segment_read
f.restype = ctypes.c_int
f.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
out = torch.zeros(4, device="cuda")
def run(n, ld, ncols, vec, rpc, label):
    X = torch.empty((n, ld), dtype=torch.float32, device="cuda")
    X.view(-1)[: 1 << 20].normal_()
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(2): f(X.data_ptr(), n, ld, ncols, vec, rpc, out.data_ptr(), st)
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(5): f(X.data_ptr(), n, ld, ncols, vec, rpc, out.data_ptr(), st)
    e1.record(); torch.cuda.synchronize()
    gbs = 5 * n * ncols * 4 / (e0.elapsed_time(e1) * 1e-3) / 1e9
    print(f"{label}: segment {vec*512} B, row stride {ld*4} B, {rpc} rows/CTA: {gbs:.0f} GB/s", flush=True)
    del X
n = 700_000
for rpc in (2048, 8192):
    for vec in (1, 2, 4):
        run(n, 18016, 17920, vec, rpc, "row-major 18016")
run(n * 8, 128, 128, 1, 8192, "slab (ld=128, contiguous)")
run(n * 4, 256, 256, 2, 8192, "slab (ld=256, contiguous)")
run(n * 2, 512, 512, 4, 8192, "slab (ld=512, contiguous)")
