"""C5-shaped run under torchrun: every rank owns a slab of the DCT known-answer matrix
(n_h total x 4096 fp64), row-sharded Gram + NCCL all-reduce, replicated eigensolve, basis, projection.
    python -m torch.distributed.run --nproc-per-node N tools/bench_c5_multi.py [n_h_total] [mode]
Prints one JSON line from rank 0 (also written to gpurun_out/c5_multi_N.json)."""
import ctypes as C, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from pod_uqnn_b200 import _lib, sharded, workloads as wl

world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
n_h = int(sys.argv[1]) if len(sys.argv) > 1 else 1_250_000 * world
mode = sys.argv[2] if len(sys.argv) > 2 else "gram2"
n_s, r, eps = 4096, 256, 1e-10
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = _lib.Context(local); L = _lib.lib(); be = sharded.CudaBackend(ctx)
lo, hi = sharded.row_range(n_h, rank, world)
U = ctx.alloc(hi - lo, n_s)
ctx.check(L.pb_dct_lowrank(ctx.h, n_h, n_s, r, 0.25, lo, hi - lo, U.ptr, U.ld)); ctx.sync()
pk = C.c_double(); ctx.check(L.pb_probe_fp64_peak(ctx.h, 1, C.byref(pk)))


def step():
    V, sig, n_L = sharded.pod_sharded(be, U, eps, 0, mode)
    v = sharded.project_sharded(be, V, U)
    return V, sig, n_L, v


def barrier():
    ctx.sync()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


step(); barrier()
times = []
for _ in range(2):
    barrier(); t = time.perf_counter(); V, sig, n_L, v = step(); barrier(); times.append(time.perf_counter() - t)
sec = min(times)
if world > 1:
    tt = torch.tensor([sec], dtype=torch.float64, device=f"cuda:{local}"); dist.all_reduce(tt, op=dist.ReduceOp.MAX); sec = float(tt.item())
s = 2. ** (-0.25 * np.arange(r))
flops = n_h * n_s * (n_s + 1) + 4 * n_h * n_s * n_L
if mode == "tsqr":      # Householder count of the local QRs + basis + projection
    flops = 2 * n_h * n_s ** 2 - world * (2 / 3) * n_s ** 3 + 4 * n_h * n_s * n_L
if rank == 0:
    out = {"n_gpus": world, "n_h": n_h, "n_s": n_s, "mode": mode, "n_L": n_L, "n_L_closed_form": wl.dct_rank(s, eps), "wall_s": sec,
           "pod_tflops": flops / sec / 1e12, "fp64_peak_per_gpu": pk.value, "frac_of_aggregate_peak": flops / sec / 1e12 / (pk.value * world),
           "sigma_err_rel_sigma_max": float(np.max(np.abs(sig[:n_L] - s[:n_L]))), "gram_kernel_ms": ctx.ms(_lib.T_TN_KERNEL)}
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open(f"gpurun_out/c5_multi_{world}.json", "w"), indent=1)
    print(json.dumps(out))
if world > 1:
    dist.destroy_process_group()
