import os, torch, time
import torch.distributed as dist
rank=int(os.environ['RANK']); world=int(os.environ['WORLD_SIZE'])
torch.cuda.set_device(rank)
dist.init_process_group('nccl', device_id=torch.device('cuda', rank))
if rank==0:
    print('can access peer', [[torch.cuda.can_device_access_peer(i,j) for j in range(world) if j!=i] for i in range(world)])
x=torch.zeros(4096*66, dtype=torch.float64, device='cuda'); out=torch.zeros(world*4096*66, dtype=torch.float64, device='cuda')
for _ in range(5): dist.all_gather_into_tensor(out,x)
torch.cuda.synchronize()
for n in (1, 4096*66, 4096*66*16):
    xx=torch.zeros(n, dtype=torch.float64, device='cuda'); oo=torch.zeros(world*n, dtype=torch.float64, device='cuda')
    for _ in range(5): dist.all_gather_into_tensor(oo,xx)
    torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): dist.all_gather_into_tensor(oo,xx)
    e1.record(); torch.cuda.synchronize()
    if rank==0: print('all_gather', n*8, 'B per rank:', e0.elapsed_time(e1)/20*1e3, 'us')
# raw p2p copy
if rank==0 and world>1:
    a=torch.zeros(32*1024*1024//8, dtype=torch.float64, device='cuda:0'); b=torch.zeros_like(a, device='cuda:1')
    for _ in range(3): b.copy_(a)
    torch.cuda.synchronize(); t=time.perf_counter()
    for _ in range(10): b.copy_(a)
    torch.cuda.synchronize(); print('p2p copy GB/s', 10*a.numel()*8/(time.perf_counter()-t)/1e9)
dist.destroy_process_group()
