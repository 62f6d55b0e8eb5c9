"""Where the `bin + all-reduce` stage of the slab pipeline goes (torchrun, G ranks): python profiles/bin_stage.py [N=2048]"""
import os, sys, time, torch, torch.distributed as dist
sys.path.insert(0, '.')
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=dev)
import pylians_b200
from pylians_b200 import dist as pdist, Pk_library as PKL
pylians_b200.set_verbose(False)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
eng = pdist.SlabPk(N, 1000.0, "PCS", 2, exchange="particles")
nyl, P = N // world, N // 2 + 2
gen = torch.Generator(device=dev); gen.manual_seed(3 + rank)
dk = torch.view_as_complex(torch.randn((N, nyl, P, 2), device=dev, generator=gen))
def sync():
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
mi = [PKL.MAS_function("PCS")]
for it in range(4):
    sync(); t0 = time.perf_counter()
    L, sums, counts = eng.ops.bin([dk], N, 2, mi, True, rank * nyl, nyl)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    raw = sums._pylb_raw
    tail = raw[L.n_doubles:]
    tail.copy_(counts.to(torch.float64))
    torch.cuda.synchronize(); t2 = time.perf_counter()
    dist.all_reduce(raw, op=dist.ReduceOp.SUM)
    torch.cuda.synchronize(); t3 = time.perf_counter()
    counts.copy_(tail.clone().round_().to(torch.int64))
    torch.cuda.synchronize(); t4 = time.perf_counter()
    sync(); t5 = time.perf_counter()
    b = eng.bin([dk], ["PCS"], True)
    sync(); t6 = time.perf_counter()
    if rank == 0 and it >= 2:
        print("N=%d G=%d: raw %d doubles (%.1f MB); pk_bin %.3f ms, counts->f64 %.3f, all_reduce %.3f, f64->counts %.3f; eng.bin as a whole %.3f ms" % (
            N, world, raw.numel(), raw.numel() * 8 / 1e6, (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, (t4 - t3) * 1e3, (t6 - t5) * 1e3), flush=True)
dist.destroy_process_group()
