"""Duration of the peer-memory data-parallel kernel alone (csrc/dp.cu), ranks aligned by a barrier before every launch;
debug variants isolate the peer loads / peer stores.  Run under torchrun."""
import os, sys, threading, torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from r2l_b200 import _lib, ops
from r2l_b200.parallel import PeerDataParallel
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
N = ops.NUM_PARAMS
dp = PeerDataParallel(N, dev)
dp.grads.normal_(); dp.params.normal_()
m, v = torch.zeros(N, device=dev), torch.zeros(N, device=dev)
hyper = torch.tensor([5e-4, 1.0, 5e-4, 0.0], device=dev)
L = _lib.lib()
for name, grid, variant in (("production", 0, 0), ("plain ld.cg", 0, 1), ("no peer stores", 0, 2), ("no peer loads", 0, 4), ("neither (local Adam only)", 0, 6),
                            ("grid 148", 148, 0), ("grid 444", 444, 0)):
    L.r2l_debug_set_dp_grid(grid, variant)
    ts = []
    for it in range(12):
        dist.barrier(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); dp.adam_step(m, v, 0.9, 0.999, 1e-8, hyper); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    t = torch.tensor([sorted(ts[2:])[len(ts[2:]) // 2]], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        mb = (world - 1) / world * N * 4 / 1e6
        print(f"world {world} {name:28s}: {float(t) * 1e3:7.1f} us  ({mb:.1f} MB in + {mb:.1f} MB out per GPU over NVLink -> {mb / float(t) / 1e3 * 1e3:.0f} GB/s per direction if both)", flush=True)
L.r2l_debug_set_dp_grid(0, 0)
sys.stdout.flush()
threading.Timer(15.0, lambda: os._exit(0)).start()
dist.barrier(); dist.destroy_process_group(); os._exit(0)
