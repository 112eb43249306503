"""Latency of the small exchanges: one-shot peer all-reduce vs NCCL (run under torchrun)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
from parapint_b200 import Communicator
comm = Communicator()
for n in (58, 2508, 16384):
    bufs = comm.exchange_buffers(n, dev)
    plain = torch.zeros(n, dtype=torch.float64, device=dev)
    for b in bufs:
        b.fill_(rank + 1.0)
    out = comm.allreduce_sum_(bufs[0])
    torch.cuda.synchronize()
    want = sum(range(1, comm.size + 1))
    ok = bool(torch.all(out == want))
    res = {}
    for name in ("peer", "nccl"):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        turn = 1
        for it in range(220):
            if it == 20:
                torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize(); ev0.record()
            if name == "peer":
                comm.allreduce_sum_(bufs[turn % len(bufs)]); turn += 1
            else:
                dist.all_reduce(plain)
        ev1.record(); torch.cuda.synchronize()
        res[name] = ev0.elapsed_time(ev1) / 200 * 1e3
    if rank == 0:
        print(f"n={n:6d} doubles  peer buffers: {len(bufs) == 2} (error: {comm.peer_error})  correct: {ok}  peer {res['peer']:.1f} us  nccl {res['nccl']:.1f} us", flush=True)
dist.destroy_process_group()
