import sys, ctypes, numpy as np, torch
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
from rawnanoporesignalcompression_b200 import synth
from rawnanoporesignalcompression_b200.codec import Codec
R=int(sys.argv[1]) if len(sys.argv)>1 else 20000
codec=int(sys.argv[2]) if len(sys.argv)>2 else 6
dev=torch.device('cuda:0')
cd=Codec(0)
lengths=synth.read_lengths(R,20261019)
sig=synth.generate_reads(lengths,20261019,synth.R9_4_1,device=dev)
off,n,_=synth.chunk_table(lengths)
comp,coff,st=cd.compress_flat(sig.cpu().numpy(),off.astype(np.uint64),n,codec)
d_comp=torch.from_numpy(np.concatenate([comp,np.zeros(64,np.uint8)])).to(dev)
d_coff=torch.from_numpy(coff.astype(np.int64)).to(dev); d_n=torch.from_numpy(n.astype(np.int32)).to(dev); d_off=torch.from_numpy(off.astype(np.int64)).to(dev)
out=torch.empty_like(sig); stt=torch.empty(len(n),dtype=torch.int32,device=dev)
def step(): cd.dev_decode(codec,d_comp,d_coff,d_n,d_off,int(n.max()),out,stt)
for _ in range(5): step()
torch.cuda.synchronize()
cd.set_profiling(True)
for it in range(6):
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record(); step(); e1.record(); torch.cuda.synchronize()
    b=cd.cta_spans()
    t0=min(int(b[k,:,0][b[k,:,2]>0].min()) for k in range(3) if (b[k,:,2]>0).any())
    line=f'step {e0.elapsed_time(e1):.2f} ms:'
    for k,name in ((2,'B'),(1,'A'),(0,'S')):
        m=b[k,:,2]>0
        if not m.any(): continue
        s=(b[k,:,0][m].astype(np.int64)-t0)/1e3; e=(b[k,:,1][m].astype(np.int64)-t0)/1e3
        line+=f' {name}[{m.sum()} ctas start {s.min():.0f}-{s.max():.0f} end {e.min():.0f}/{np.percentile(e,90):.0f}/{e.max():.0f}]'
    print(line)
