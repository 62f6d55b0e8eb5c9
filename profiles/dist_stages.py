"""Per-stage wall times (with device syncs) of SlabPk in both exchange modes.  Launch with torchrun."""
import os, sys, time, torch, torch.distributed as dist
sys.path.insert(0, '.')
import pylians_b200
from pylians_b200 import dist as pdist, Pk_library as PKL
pylians_b200.set_verbose(False)
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank); dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
# python profiles/dist_stages.py [particles-per-GPU side = 1024] [MAS = PCS]
nside = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
MAS = sys.argv[2] if len(sys.argv) > 2 else "PCS"
N = int(round(nside * {1: 1.0, 2: 1.25, 4: 1.5625, 8: 2.0}[world])); box = 1000.0
gen = torch.Generator(device=dev); gen.manual_seed(1 + rank)
pos = torch.rand((nside ** 3, 3), device=dev, generator=gen) * box
def sync():
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
for mode in ("grid", "particles"):
    eng = pdist.SlabPk(N, box, MAS, 2, exchange=mode)
    for it in range(3):
        sync(); t0 = time.perf_counter()
        slab = eng.density_slab(pos); sync(); t1 = time.perf_counter()
        dk = eng.fft_slab(slab); sync(); t2 = time.perf_counter()
        b = eng.bin([dk], [MAS], True); sync(); t3 = time.perf_counter()
    if rank == 0:
        print("%-9s N=%d G=%d: density_slab %.2f ms  fft_slab %.2f ms  bin+allreduce %.2f ms  total %.2f ms" % (
            mode, N, world, (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, (t3 - t0) * 1e3))
# particle mode: pieces of the routed payload overlapped with the windowed deposit (exchange_chunks), and the local kernels
for chunks in (1, 2, 4):
    eng = pdist.SlabPk(N, box, MAS, 2, exchange="particles", exchange_chunks=chunks)
    for it in range(3):
        sync(); t0 = time.perf_counter(); slab = eng.density_slab(pos); sync(); t1 = time.perf_counter()
    if rank == 0:
        print("particles, %d piece(s): density_slab %.2f ms" % (chunks, (t1 - t0) * 1e3))
eng = pdist.SlabPk(N, box, MAS, 2, exchange="particles"); ops = eng.ops
for it in range(2):
    sync(); t0 = time.perf_counter()
    send, offsets = ops.partition(pos, None, box, MAS, world, N); sync(); t1 = time.perf_counter()
    off = offsets.to("cpu").tolist(); ss = [off[g + 1] - off[g] for g in range(world)]
    ts = torch.tensor(ss, dtype=torch.int64, device=dev); tr = torch.empty_like(ts); dist.all_to_all_single(tr, ts); rs = tr.to("cpu").tolist(); sync(); t2 = time.perf_counter()
    recv = send.new_empty((sum(rs), 4)); dist.all_to_all_single(recv, send, output_split_sizes=rs, input_split_sizes=ss); sync(); t3 = time.perf_counter()
    grid = ops.zeros((eng.nxl + {"NGP": 0, "CIC": 1, "TSC": 2, "PCS": 3}[MAS], N, N)); ops.deposit_window(recv, grid, rank * eng.nxl, box, MAS, False, N); sync(); t4 = time.perf_counter()
if rank == 0:
    print("particles, stages one after the other: partition %.2f  splits %.2f  all_to_all(%.2f GB) %.2f  window deposit %.2f ms" % (
        (t1 - t0) * 1e3, (t2 - t1) * 1e3, send.numel() * 4 / 1e9, (t3 - t2) * 1e3, (t4 - t3) * 1e3))
dist.destroy_process_group()
