import torch, time
d=torch.device('cuda')
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); t0=time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter()-t0)/n*1e6
for mb in (1, 2, 8, 18, 64):
    n=mb<<20
    hp=torch.empty(n,dtype=torch.uint8).pin_memory(); dv=torch.empty(n,dtype=torch.uint8,device=d)
    h2d=t(lambda: dv.copy_(hp,non_blocking=True)); d2h=t(lambda: hp.copy_(dv,non_blocking=True))
    print(f"{mb} MB: H2D {h2d:.0f} us ({n/h2d/1e3:.1f} GB/s)  D2H {d2h:.0f} us ({n/d2h/1e3:.1f} GB/s)")
# both directions at once
n=18<<20
hp1=torch.empty(n,dtype=torch.uint8).pin_memory(); dv1=torch.empty(n,dtype=torch.uint8,device=d)
hp2=torch.empty(n,dtype=torch.uint8).pin_memory(); dv2=torch.empty(n,dtype=torch.uint8,device=d)
s1=torch.cuda.Stream(); s2=torch.cuda.Stream()
def both():
    with torch.cuda.stream(s1): dv1.copy_(hp1,non_blocking=True)
    with torch.cuda.stream(s2): hp2.copy_(dv2,non_blocking=True)
print("bidir 18MB each:", t(both))
