import sys, os
ROOT="/root/repo"; sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tiny-ntt_b200"))
import torch, tntt
Q60=(1<<60)-(1<<14)+1
n,q,psi=2048,Q60,644283108363935541
plan=tntt.get_plan(n,q,psi,True); rows=32768
g=torch.Generator(device="cuda").manual_seed(1)
a=torch.randint(0,q,(rows,n),generator=g,device="cuda",dtype=torch.int64); b=torch.randint(0,q,(rows,n),generator=g,device="cuda",dtype=torch.int64); c=torch.empty_like(a)
for v,d in plan.variants():
    for _ in range(3): tntt.polymul(plan,a,b,out=c,variant=v)
    torch.cuda.synchronize(); e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True); e0.record()
    for _ in range(10): tntt.polymul(plan,a,b,out=c,variant=v)
    e1.record(); torch.cuda.synchronize(); print(d.split()[0], round(rows/(e0.elapsed_time(e1)/10*1e-3)/1e6,2), d.split("regs=")[1])
