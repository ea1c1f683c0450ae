import time, numpy as np, torch
torch.cuda.set_device(0)
def sync(): torch.cuda.synchronize(); return time.perf_counter()
x = torch.randn(700000, 12, dtype=torch.float64, device="cuda")
h32 = np.random.rand(700000, 30).astype(np.float32)
for rep in range(3):
    t0 = sync(); a = x.cpu().numpy(); t1 = sync()
    b = np.empty((700000, 12)); torch.from_numpy(b).copy_(x); t2 = sync()
    p = torch.empty((700000, 12), dtype=torch.float64, pin_memory=True); p.copy_(x, non_blocking=True); t3 = sync()
    c = p.numpy().copy(); t4 = sync()
    u1 = torch.from_numpy(h32).pin_memory().to("cuda", non_blocking=True); t5 = sync()
    u2 = torch.from_numpy(h32).to("cuda"); t6 = sync()
    big = np.random.rand(1 << 27).astype(np.float32) if rep == 0 else big   # 512 MB
    t6 = sync(); u3 = torch.from_numpy(big).to("cuda"); t7 = sync()
    u4 = torch.from_numpy(big).pin_memory().to("cuda", non_blocking=True); t8 = sync()
    print(f"rep{rep} d2h .cpu {1e3*(t1-t0):.1f}  np.empty+copy_ {1e3*(t2-t1):.1f}  pinned alloc+copy {1e3*(t3-t2):.1f} (+numpy copy {1e3*(t4-t3):.1f})  "
          f"h2d 84MB pin {1e3*(t5-t4):.1f} pageable {1e3*(t6-t5):.1f}  512MB pageable {1e3*(t7-t6):.1f} pin {1e3*(t8-t7):.1f}")
