import sys, time, numpy as np, torch
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
from rawnanoporesignalcompression_b200 import synth
from rawnanoporesignalcompression_b200.codec import Codec
R=int(sys.argv[1]) if len(sys.argv)>1 else 6000
codec=int(sys.argv[2]) if len(sys.argv)>2 else 0
dev=torch.device('cuda:0')
cd=Codec(0)
lengths=synth.read_lengths(R,20261019)
sig=synth.generate_reads(lengths,20261019,synth.R9_4_1,device=dev)
off,n,_=synth.chunk_table(lengths)
host=sig.cpu().numpy()
comp,coff,st=cd.compress_flat(host,off.astype(np.uint64),n,codec)
assert not st.any()
d_comp=torch.from_numpy(np.concatenate([comp,np.zeros(64,np.uint8)])).to(dev)
d_coff=torch.from_numpy(coff.astype(np.int64)).to(dev); d_n=torch.from_numpy(n.astype(np.int32)).to(dev); d_off=torch.from_numpy(off.astype(np.int64)).to(dev)
out=torch.empty_like(sig); stt=torch.empty(len(n),dtype=torch.int32,device=dev)
def step(): cd.dev_decode(codec,d_comp,d_coff,d_n,d_off,int(n.max()),out,stt)
for _ in range(3): step()
torch.cuda.synchronize()
assert torch.equal(out,sig) and not stt.any()
if '--phases' in sys.argv:
    cd.set_profiling(True)
    for _ in range(3): step()
    for cls,ph in cd.phase_cycles().items():
        k=max(1,ph['chunks']); print(cls, ph['chunks']//3, {a:int(b/k) for a,b in ph.items() if a!='chunks'})
    cd.set_profiling(False)
N=20
ev=[torch.cuda.Event(enable_timing=True) for _ in range(N+1)]
ev[0].record()
for i in range(N): step(); ev[i+1].record()
torch.cuda.synchronize()
ts=sorted(ev[i].elapsed_time(ev[i+1]) for i in range(N))
ms=ts[N//2]
if '--steps' in sys.argv: print('step ms:', ' '.join(f'{ev[i].elapsed_time(ev[i+1]):.2f}' for i in range(N)))
print(f"codec {codec}: {2e-6*sig.numel()/ms:.1f} GB/s median ({2e-6*sig.numel()/ts[0]:.1f} best), {ms:.3f} ms, routing {cd.decode_class_counts()}, bits/sample {8*coff[-1]/sig.numel():.3f}")
