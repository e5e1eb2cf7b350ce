import torch
shapes=[(64,64)]*2+[(64,128),(128,128),(128,256),(256,256),(256,512),(512,512)]+[(512,512)]*4+[(1024,256),(256,256),(512,128),(128,128),(256,64),(64,64),(128,32),(32,32),(32,32),(32,32)]
ts=[torch.randn(27,ci,co,device='cuda') for ci,co in shapes]
def run():
    return [t.view(3,3,3,t.shape[1],t.shape[2]).permute(4,3,0,1,2).contiguous() for t in ts]
for _ in range(3): run()
torch.cuda.synchronize()
e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): out=run()
e1.record(); torch.cuda.synchronize()
n=sum(t.numel() for t in ts)
print("permute+contiguous of all conv dW:", e0.elapsed_time(e1)/10, "ms for", n*4/1e6, "MB")
flat=torch.empty(n,device='cuda')
views=[v.view(o.shape) for v,o in zip(flat.split([o.numel() for o in out]),out)]
e0.record()
for _ in range(10): torch._foreach_copy_(views,out)
e1.record(); torch.cuda.synchronize()
print("foreach_copy:", e0.elapsed_time(e1)/10)
e0.record()
for _ in range(10): c=flat.clone()
e1.record(); torch.cuda.synchronize()
print("clone:", e0.elapsed_time(e1)/10)
e0.record()
for _ in range(10): z=[torch.zeros_like(t) for t in ts]
e1.record(); torch.cuda.synchronize()
print("zeros:", e0.elapsed_time(e1)/10)
