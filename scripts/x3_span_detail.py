import ctypes, os, sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
from diffusionspatialcontrol_b200 import _lib, attention as att
B,L,D,H,S=16,4096,40,8,77
dev=torch.device('cuda')
q=torch.randn(B,L,H*D,device=dev).half(); k=torch.randn(B,S,H*D,device=dev).half(); v=torch.randn(B,S,H*D,device=dev).half()
W=torch.zeros(B,L,S,device=dev); W[:,:L//2,1:3]=0.5; W[:,L//3:,6]=0.7
W=att.padded_region_map(W); compact=att.compact_region_map(W)
view=lambda t:t.view(B,-1,H,D).transpose(1,2)
kv=att.prepare_kv(view(k),view(v),compact[1]); out=torch.empty(B,L,H*D,device=dev,dtype=torch.float16)
flush=torch.empty(512<<20,dtype=torch.uint8,device=dev)
acc=[]
for it in range(12):
    flush.zero_(); flush[:flush.numel()//2].view(torch.int64).sum()
    att.region_attention_prepared(view(q),kv,compact,7.0,out=out); torch.cuda.synchronize()
    cta=np.zeros((2,160,2),dtype=np.uint64); _lib.lib.dsc_debug_x3_cta(cta.ctypes.data_as(ctypes.c_void_p))
    c=cta[:,:148].astype(np.int64); t0=c[0,:,0].min()
    if it>=2: acc.append(np.stack([c[0,:,1]-t0, c[1,:,0]-t0, c[1,:,1]-t0],axis=1))
a=np.mean(np.array(acc),axis=0)/1e3
print("block: p1_end p2_start p2_end (us), mean over 10 calls")
order=np.argsort(a[:,2])
print("fastest 8:", [(int(b), round(a[b,0],1), round(a[b,2],1)) for b in order[:8]])
print("slowest 16:", [(int(b), round(a[b,0],1), round(a[b,2],1)) for b in order[-16:]])
print("p2 dur by block (us):", " ".join(f"{a[b,2]-a[b,1]:.1f}" for b in range(148)))
sd=np.std(np.array(acc)[:,:,2]/1e3,axis=0)
print("per-block std of p2_end over calls: mean", sd.mean().round(2), "max", sd.max().round(2))
