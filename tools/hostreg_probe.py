import time, numpy as np, torch
rt = torch.cuda.cudart()
torch.cuda.init()
for mb in (32, 128):
    a = np.ones(mb << 20, dtype=np.uint8)
    ts = []
    for i in range(6):
        t0 = time.perf_counter(); r = rt.cudaHostRegister(a.ctypes.data, a.nbytes, 0); t1 = time.perf_counter(); rt.cudaHostUnregister(a.ctypes.data); t2 = time.perf_counter()
        ts.append(((t1 - t0) * 1e3, (t2 - t1) * 1e3))
    print(mb, "MB register/unregister ms:", [(round(x, 3), round(y, 3)) for x, y in ts], int(r))
    d = torch.empty(a.nbytes, dtype=torch.uint8, device="cuda")
    rt.cudaHostRegister(a.ctypes.data, a.nbytes, 0)
    t = torch.from_numpy(a)
    torch.cuda.synchronize(); t0 = time.perf_counter(); d.copy_(t, non_blocking=True); torch.cuda.synchronize(); print("  copy registered", (time.perf_counter() - t0) * 1e3)
    rt.cudaHostUnregister(a.ctypes.data)
    torch.cuda.synchronize(); t0 = time.perf_counter(); d.copy_(t); torch.cuda.synchronize(); print("  copy pageable", (time.perf_counter() - t0) * 1e3)
