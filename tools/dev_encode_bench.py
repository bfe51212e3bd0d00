"""Device-resident whole-chunk encode with the GPU entropy stage (rnsc_dev_encode), timing aid: tools/dev_encode_bench.py READS CODEC [--model M] [--lib so]"""
import os, sys, argparse
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rawnanoporesignalcompression_b200 import synth
from rawnanoporesignalcompression_b200.codec import Codec
ap = argparse.ArgumentParser()
ap.add_argument('reads', type=int, nargs='?', default=6000)
ap.add_argument('codec', type=int, nargs='?', default=6)
ap.add_argument('--model', default='R10.4.1')
ap.add_argument('--lib', default=None)
ap.add_argument('--steps', type=int, default=10)
a = ap.parse_args()
seed = 20261019 if a.model == 'R9.4.1' else 20261018
dev = torch.device('cuda:0')
cd = Codec(0, lib_path=a.lib) if a.lib else Codec(0)
lengths = synth.read_lengths(a.reads, seed)
sig = synth.generate_reads(lengths, seed, synth.MODELS[a.model], device=dev)
off, n, _ = synth.chunk_table(lengths)
total = int(lengths.sum()); nch = n.size
d_off = torch.from_numpy(off.astype(np.int64)).to(dev); d_n = torch.from_numpy(n.astype(np.int32)).to(dev)
d_comp = torch.empty(2 * total + total // 4 + 1104 * nch + 64, dtype=torch.uint8, device=dev)
d_coff = torch.empty(nch + 1, dtype=torch.int64, device=dev); d_st = torch.empty(nch, dtype=torch.int32, device=dev)
def step(): cd.dev_encode(a.codec, sig, d_off, d_n, total, d_comp, d_coff, d_st)
for _ in range(2): step()
torch.cuda.synchronize()
assert not d_st.any().item()
d_out = torch.empty_like(sig); d_st2 = torch.empty_like(d_st)
cd.dev_decode(a.codec, d_comp, d_coff, d_n, d_off, int(n.max()), d_out, d_st2); torch.cuda.synchronize()
ok = bool(torch.equal(d_out, sig)) and not bool(d_st2.any())
ev = [torch.cuda.Event(enable_timing=True) for _ in range(a.steps + 1)]
ev[0].record()
for i in range(a.steps): step(); ev[i + 1].record()
torch.cuda.synchronize()
ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(a.steps)); ms = ts[a.steps // 2]
cb = int(d_coff[-1].item())
print(f"dev_encode codec {a.codec} {a.model} {a.reads} reads: {2e-6*total/ms:.1f} GB/s median ({2e-6*total/ts[0]:.1f} best), {ms:.3f} ms, round trip {ok}, bits/sample {8*cb/total:.4f}, routing of the decode {cd.decode_class_counts()}", flush=True)
