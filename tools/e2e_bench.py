import sys, time, os, numpy as np, torch
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0,ROOT); sys.path.insert(0,os.path.join(ROOT,'tests'))
from rawnanoporesignalcompression_b200 import synth
from rawnanoporesignalcompression_b200.codec import Codec, _ptr
R=int(sys.argv[1]) if len(sys.argv)>1 else 30000
dev=torch.device('cuda:0'); cd=Codec(0)
lengths=synth.read_lengths(R,20261018)
off,n,_=synth.chunk_table(lengths)
total=int(lengths.sum())
sig=torch.empty(total,dtype=torch.int16,device=dev)
rs=np.cumsum(lengths)-lengths
for a in range(0,R,4096):
    b=min(R,a+4096); s0=int(rs[a]); s1=int(rs[b-1]+lengths[b-1])
    synth.generate_reads(lengths[a:b],20261018,synth.R10_4_1,first_read=a,device=dev,out=sig[s0:s1])
host=sig.cpu().numpy()
comp,coff,st=cd.compress_flat(host,off.astype(np.uint64),n,6)
h_comp=torch.empty(comp.size+64,dtype=torch.uint8).pin_memory(); h_comp[:comp.size].copy_(torch.from_numpy(comp))
h_out=torch.empty(total,dtype=torch.int16).pin_memory()
hc,ho=h_comp.numpy(),h_out.numpy()
coff=coff.astype(np.uint64); nn=n.astype(np.uint32); soff=off.astype(np.uint64); status=np.zeros(len(n),np.int32)
def step():
    r=cd.lib.rnsc_decompress_batch(cd.ctx,6,_ptr(hc),_ptr(coff),_ptr(nn),len(n),_ptr(ho),_ptr(soff),_ptr(status)); assert r==0
for _ in range(2): step()
t0=time.perf_counter()
for _ in range(3): step()
dt=(time.perf_counter()-t0)/3
assert (ho==host).all()
print(f"slice {os.environ.get('RNSC_SLICE_SAMPLES','default')}: e2e {2e-9*total/dt:.2f} GB/s ({dt*1e3:.1f} ms, D2H {2e-9*total/dt:.1f} GB/s, H2D {1e-9*comp.size/dt:.1f} GB/s)")
