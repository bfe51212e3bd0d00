"""Device-resident whole-chunk decode timing (tuning aid): tools/decode_bench.py READS CODEC [--model R9.4.1|R10.4.1]
[--lib path/to/librnsc.so[,other.so ...]] [--phases] [--steps] [--cache file.npz]. Several libraries (comma-separated) decode
the very same device buffers one after another in one process. The workload (synthetic reads, encoded by the product
encoder) is cached in --cache so that several library variants / settings decode the very same bytes."""
import os, sys, time, argparse
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from rawnanoporesignalcompression_b200 import synth
from rawnanoporesignalcompression_b200.codec import Codec
ap = argparse.ArgumentParser()
ap.add_argument('reads', type=int, nargs='?', default=6000)
ap.add_argument('codec', type=int, nargs='?', default=0)
ap.add_argument('--model', default='R9.4.1')
ap.add_argument('--lib', default=None)
ap.add_argument('--phases', action='store_true')
ap.add_argument('--steps', action='store_true')
ap.add_argument('--cache', default=None)
ap.add_argument('--tag', default='')
a = ap.parse_args()
seed = 20261019 if a.model == 'R9.4.1' else 20261018
dev = torch.device('cuda:0')
libs = a.lib.split(',') if a.lib else [None]
cd = Codec(0, lib_path=libs[0]) if libs[0] else Codec(0)
lengths = synth.read_lengths(a.reads, seed)
off, n, _ = synth.chunk_table(lengths)
if a.cache and os.path.exists(a.cache):
    z = np.load(a.cache); host, comp, coff = z['host'], z['comp'], z['coff']
    sig = torch.from_numpy(host).to(dev)
elif a.reads > 8192:      # the bench's own builder: generated and encoded in shards of 4096 reads
    sys.path.insert(0, ROOT); import bench
    w = bench.build_workload(cd, torch, dev, a.codec, lengths, 0, synth.MODELS[a.model], seed)
    sig, comp, coff = w['d_sig'], w['comp'][:-64], w['coff']
else:
    sig = synth.generate_reads(lengths, seed, synth.MODELS[a.model], device=dev)
    host = sig.cpu().numpy()
    comp, coff, st = cd.compress_flat(host, off.astype(np.uint64), n, a.codec)
    assert not st.any()
    if a.cache: np.savez(a.cache, host=host, comp=comp, coff=coff)
d_comp = torch.from_numpy(np.concatenate([comp, np.zeros(64, np.uint8)])).to(dev)
d_coff = torch.from_numpy(coff.astype(np.int64)).to(dev); d_n = torch.from_numpy(n.astype(np.int32)).to(dev); d_off = torch.from_numpy(off.astype(np.int64)).to(dev)
out = torch.empty_like(sig); stt = torch.empty(len(n), dtype=torch.int32, device=dev)
for li, lp in enumerate(libs):
    if li: cd = Codec(0, lib_path=lp)
    tag = a.tag + (" " + os.path.basename(lp) if lp else "")
    def step(): cd.dev_decode(a.codec, d_comp, d_coff, d_n, d_off, int(n.max()), out, stt)
    for _ in range(3): step()
    torch.cuda.synchronize()
    ok = bool(torch.equal(out, sig)) and not bool(stt.any())
    if a.phases:
        cd.set_profiling(True)
        for _ in range(3): step()
        for cls, ph in cd.phase_cycles().items():
            k = max(1, ph['chunks']); print(tag, cls, ph['chunks'] // 3, {x: int(y / k) for x, y in ph.items() if x != 'chunks'})
        cd.set_profiling(False)
    N = 20
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(N + 1)]
    ev[0].record()
    for i in range(N): step(); ev[i + 1].record()
    torch.cuda.synchronize()
    ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(N))
    ms = ts[N // 2]
    if a.steps: print('step ms:', ' '.join(f'{ev[i].elapsed_time(ev[i+1]):.2f}' for i in range(N)))
    print(f"{tag} codec {a.codec} {a.model} {a.reads} reads: {2e-6*sig.numel()/ms:.1f} GB/s median ({2e-6*sig.numel()/ts[0]:.1f} best), {ms:.3f} ms, bit-exact {ok}, routing {cd.decode_class_counts()}, bits/sample {8*coff[-1]/sig.numel():.3f}", flush=True)
