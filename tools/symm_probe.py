"""Probe of torch's symmetric-memory plumbing on this pool (peer pointers, NVLS multicast), 2+ ranks.
usage: python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/symm_probe.py"""
import os
import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=dev)
t = symm.empty(1 << 20, dtype=torch.float32, device=dev)
t.fill_(rank + 1.0)
hdl = symm.rendezvous(t, dist.group.WORLD)
print(f"rank {rank}: backend={symm.get_backend(dev)} "
      f"multicast_ptr={hdl.multicast_ptr:#x} buffer_ptrs={[hex(p) for p in hdl.buffer_ptrs]} signal_pad_size={hdl.signal_pad_size} "
      f"buffer_size={hdl.buffer_size}", flush=True)
hdl.barrier()
peer = hdl.get_buffer((rank + 1) % world, (16,), torch.float32)
print(f"rank {rank}: peer values {peer[:2].tolist()}", flush=True)
hdl.barrier()
torch.cuda.synchronize()
dist.barrier()
torch.cuda.synchronize()
os._exit(0)
