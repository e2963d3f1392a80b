"""torchrun probe: symmetric memory rendezvous and peer pointers on this box."""
import os, sys
import torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm
rank = int(os.environ['RANK']); local = int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
t = symm.empty(4096, dtype=torch.float64, device=f'cuda:{local}')
hdl = symm.rendezvous(t, dist.group.WORLD)
print(rank, 'world', hdl.world_size, 'ptrs', [hex(p) for p in hdl.buffer_ptrs][:4], 'signal', [hex(p) for p in hdl.signal_pad_ptrs][:2], flush=True)
t.fill_(rank + 1.0)
dist.barrier(); torch.cuda.synchronize()
peer = hdl.get_buffer((rank + 1) % hdl.world_size, (4096,), torch.float64)
print(rank, 'peer value', peer[0].item(), flush=True)
dist.barrier()
dist.destroy_process_group()
