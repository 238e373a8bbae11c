import torch, time
a = torch.empty(1<<30, dtype=torch.uint8, device="cuda:0")
b = torch.empty(1<<30, dtype=torch.uint8, device="cuda:1")
print("can_access_peer", torch.cuda.can_device_access_peer(0,1))
for _ in range(2):
    torch.cuda.synchronize(0); torch.cuda.synchronize(1)
    t=time.time(); b.copy_(a); torch.cuda.synchronize(0); torch.cuda.synchronize(1); dt=time.time()-t
    print("copy 1 GiB 0->1: %.2f ms = %.1f GB/s" % (dt*1e3, (1<<30)/dt/1e9))
