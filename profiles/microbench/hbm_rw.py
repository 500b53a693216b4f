import torch,time
x=torch.empty(4<<30,dtype=torch.uint8,device='cuda')
y=torch.empty(4<<30,dtype=torch.uint8,device='cuda')
for name,fn,byt in [("memset",lambda:x.zero_(),4<<30),("copy",lambda:y.copy_(x),8<<30),("read(sum int32)",lambda:x.view(torch.int32).sum(),4<<30)]:
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0=torch.cuda.Event(enable_timing=True);e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): fn()
    e1.record();torch.cuda.synchronize()
    ms=e0.elapsed_time(e1)/10
    print(name, "%.3f ms  %.1f GB/s"%(ms, byt/ms/1e6))
