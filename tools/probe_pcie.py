import ctypes as C, sys, os, time
sys.path.insert(0, "/root/repo")
from pod_uqnn_b200 import _lib
ctx = _lib.Context(0); L = _lib.lib()
nb = 1280000000
h = C.c_void_p(); ctx.check(L.pb_host_alloc(nb, C.byref(h)))
C.memset(h, 1, nb)
d = ctx.alloc(nb // 8)
for chunks in (1, 4, 16, 64):
    best = 1e9
    for _ in range(3):
        ctx.sync(); ctx.timer_start()
        sz = nb // chunks
        for c in range(chunks):
            ctx.check(L.pb_upload_async(ctx.h, C.c_void_p(d.ptr + c * sz), C.c_void_p(h.value + c * sz), sz))
        ms = ctx.timer_stop(); best = min(best, ms)
    print("H2D", chunks, "chunks:", round(best, 2), "ms", round(nb / best / 1e6, 1), "GB/s")
best = 1e9
for _ in range(3):
    ctx.sync(); ctx.timer_start()
    ctx.check(L.pb_download_async(ctx.h, h, d.ptr, nb))
    ms = ctx.timer_stop(); best = min(best, ms)
print("D2H:", round(best, 2), "ms", round(nb / best / 1e6, 1), "GB/s")
