"""Latency of the packed-statistics exchange on one NVSwitch box: peer-memory kernel vs NCCL.
torchrun --nproc-per-node N tools/exchange_latency.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from spectrogram_cube_clustering_b200.latent_buffer import PeerExchange

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
ex = PeerExchange(dist.group.WORLD, dev, 1024)
for length in (9, 74, 881):
    t = torch.ones(length, dtype=torch.float64, device=dev)
    res = {}
    for name, fn in (("peer", lambda: ex.all_reduce(t)), ("nccl", lambda: dist.all_reduce(t))):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            fn(); s.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=s):
                for _ in range(50):
                    fn()
        torch.cuda.synchronize(); dist.barrier()
        g.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            g.replay()
        e1.record(); torch.cuda.synchronize()
        res[name] = e0.elapsed_time(e1) / 500 * 1e3
        del g
    if rank == 0:
        print(f"world={world} len={length:4d} doubles: peer {res['peer']:.2f} us   nccl {res['nccl']:.2f} us (back-to-back, graph replay)", flush=True)
dist.barrier()
os._exit(0)
