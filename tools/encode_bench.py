import sys, numpy as np, torch
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
from rawnanoporesignalcompression_b200 import synth
from rawnanoporesignalcompression_b200.codec import Codec
# splitter-stage encode (rnsc_dev_svb_encode: sizes -> scan -> pack) on R reads, device resident; optional oracle check
R=int(sys.argv[1]) if len(sys.argv)>1 else 20000
codecs=[int(c) for c in sys.argv[2].split(',')] if len(sys.argv)>2 else [6,0]
dev=torch.device('cuda:0')
import os
cd=Codec(0,lib_path=os.environ['RNSC_LIB']) if os.environ.get('RNSC_LIB') else Codec(0)
lengths=synth.read_lengths(R,20261018)
sig=synth.generate_reads(lengths,20261018,synth.R10_4_1,device=dev)
off,n,_=synth.chunk_table(lengths)
d_n=torch.from_numpy(n.astype(np.int32)).to(dev); d_off=torch.from_numpy(off.astype(np.int64)).to(dev)
nch=len(n)
for codec in codecs:
    cap=int(sum(cd.lib.rnsc_max_stream_bytes(codec,int(k)) for k in n))
    d_streams=torch.empty(cap,dtype=torch.uint8,device=dev); d_soff=torch.empty(5*nch+1,dtype=torch.int64,device=dev)
    def step(): cd.dev_svb_encode(codec,sig,d_off,d_n,d_streams,d_soff)
    for _ in range(3): step()
    torch.cuda.synchronize()
    if '--check' in sys.argv:
        from oracle_binding import Oracle
        orc=Oracle(); host=sig.cpu().numpy(); soff=d_soff.cpu().numpy(); sh=d_streams.cpu().numpy()
        for i in list(range(0,nch,max(1,nch//200)))+[nch-1]:
            want=orc.streams(codec,host[off[i]:off[i]+n[i]])
            got=[sh[soff[5*i+k]:soff[5*i+k+1]].tobytes() for k in range(5)]
            assert got==want,(codec,i)
        out=torch.zeros_like(sig); st=torch.full((nch,),-1,dtype=torch.int32,device=dev)
        cd.dev_svb_decode(codec,d_streams,d_soff,d_n,d_off,out,st); torch.cuda.synchronize()
        assert not st.any() and torch.equal(out,sig)
        print(f"codec {codec}: streams == oracle on {len(range(0,nch,max(1,nch//200)))+1} chunks, round trip of all {nch} chunks exact")
    N=10
    ev=[torch.cuda.Event(enable_timing=True) for _ in range(N+1)]
    ev[0].record()
    for i in range(N): step(); ev[i+1].record()
    torch.cuda.synchronize()
    ts=sorted(ev[i].elapsed_time(ev[i+1]) for i in range(N)); ms=ts[N//2]
    sb=int(d_soff[-1].item())
    print(f"codec {codec}: {2e-6*sig.numel()/ms:.1f} GB/s of int16 median ({2e-6*sig.numel()/ts[0]:.1f} best), {ms:.3f} ms, {nch} chunks, {sig.numel()} samples, stream bytes/sample {sb/sig.numel():.4f}, roofline frac {(2*sig.numel()+sb)/(ms*1e-3)/1e9/6549.4:.4f}")
    del d_streams,d_soff
