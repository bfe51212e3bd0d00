"""End-to-end whole-chunk encode through rnsc_compress_batch with pinned host buffers and the GPU entropy stage
(H2D of the samples, splitter + entropy stage on the device, D2H of the finished chunks): tools/e2e_encode_bench.py [READS] [CODEC]"""
import sys, time, os, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from rawnanoporesignalcompression_b200 import synth
from rawnanoporesignalcompression_b200.codec import Codec, _ptr
R = int(sys.argv[1]) if len(sys.argv) > 1 else 30000
codec = int(sys.argv[2]) if len(sys.argv) > 2 else 6
dev = torch.device('cuda:0'); cd = Codec(0)
lengths = synth.read_lengths(R, 20261018)
off, n, _ = synth.chunk_table(lengths)
total = int(lengths.sum())
sig = torch.empty(total, dtype=torch.int16, device=dev)
rs = np.cumsum(lengths) - lengths
for a in range(0, R, 4096):
    b = min(R, a + 4096); s0 = int(rs[a]); s1 = int(rs[b - 1] + lengths[b - 1])
    synth.generate_reads(lengths[a:b], 20261018, synth.R10_4_1, first_read=a, device=dev, out=sig[s0:s1])
h_sig = torch.empty(total, dtype=torch.int16).pin_memory(); h_sig.copy_(sig.cpu())
cap = 2 * total + total // 4 + 1104 * len(n) + 64
h_comp = torch.empty(cap, dtype=torch.uint8).pin_memory()
hs, hc = h_sig.numpy(), h_comp.numpy()
nn = n.astype(np.uint32); soff = off.astype(np.uint64); coff = np.zeros(len(n) + 1, np.uint64); status = np.zeros(len(n), np.int32)
cd.set_entropy_stage("gpu")
def step():
    r = cd.lib.rnsc_compress_batch(cd.ctx, codec, _ptr(hs), _ptr(soff), _ptr(nn), len(n), _ptr(hc), cap, _ptr(coff), _ptr(status)); assert r == 0, cd.lib.rnsc_last_error()
for _ in range(2): step()
t0 = time.perf_counter()
for _ in range(3): step()
dt = (time.perf_counter() - t0) / 3
assert not status.any()
cb = int(coff[-1])
# round trip through the host-buffer decoder
out = torch.empty(total, dtype=torch.int16).pin_memory(); ho = out.numpy(); st2 = np.zeros(len(n), np.int32)
r = cd.lib.rnsc_decompress_batch(cd.ctx, codec, _ptr(hc), _ptr(coff), _ptr(nn), len(n), _ptr(ho), _ptr(soff), _ptr(st2)); assert r == 0
ok = bool((ho == hs).all()) and not st2.any()
print(f"slice {os.environ.get('RNSC_SLICE_SAMPLES','default')}: e2e encode codec {codec} {2e-9*total/dt:.2f} GB/s of int16 ({dt*1e3:.1f} ms, H2D {2e-9*total/dt:.1f} GB/s, D2H {1e-9*cb/dt:.1f} GB/s), bits/sample {8*cb/total:.4f}, round trip {ok}")
