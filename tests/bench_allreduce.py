"""The step's one collective alone: all-reduce (sum) of the 133 MB fp32 gradient arena over NCCL, timed with CUDA events
(max over ranks).  NCCL reads its environment when the communicator is created, so one setting per launch:
   NCCL_ALGO=NVLS python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
       --master-port 29517 tests/bench_allreduce.py"""
import os

import torch
import torch.distributed as dist


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
    n = 33_270_000          # the AV model's gradient arena (floats)
    buf = torch.randn(n, device="cuda")
    for _ in range(5):
        dist.all_reduce(buf)
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters = 20
    e0.record()
    for _ in range(iters):
        dist.all_reduce(buf)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / iters], device="cuda")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        ms = float(ms)
        busbw = n * 4 * 2 * (world - 1) / world / (ms * 1e-3) / 1e9
        print({"world": world, "ms": round(ms, 4), "busbw_GBps": round(busbw, 1),
               "env": {k: v for k, v in os.environ.items() if k.startswith("NCCL_")}})
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
