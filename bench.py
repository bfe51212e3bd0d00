#!/usr/bin/env python
"""bench.py -- headline benchmark of the per-chunk raw-signal codec hot path on B200.

Workload (BASELINE.json configs[1]): synthetic R10.4.1-shaped int16 signal (5 kHz random-walk levels + noise),
100,000 reads cut into POD5 chunks of 102,400 samples, pgnano (variant C5, the reference's checked-in default)
batched whole-chunk DECODE on one B200. A "step" = one pass of the decoder over the whole chunk set.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl rnsc|reference] [--reads R] [--codec C5]

N > 1: launched by torch.distributed.run, one rank per GPU; the signal table is READ-SHARDED (rank r owns reads
[r*R, (r+1)*R): weak scaling), no collective on the data path -- only the per-rank times are gathered.

Printed JSON line (rank 0): metric/value = int16 GB/s of the device-resident decode (inputs already in HBM);
`e2e` = the same decode through the C ABI with pinned HOST buffers (H2D of the compressed chunks and D2H of
the samples inside the timed region); `roofline` for the dominant launch group (k_plan_chunks + k_decode_fast<1024|512|256>)
against the measured HBM copy peak; `cpu_baseline` = the reference's own CPU decoder (oracle/_ref) on the host cores.
`other_legs` carries the other BASELINE configurations, each with its own CPU baseline: config 3 (whole-chunk C5 encode,
bytes compared with the reference encoder's), config 4 (VBZ decode + StreamVByte-stage encode on the R9.4.1-shaped set),
config 5 (the 6.0 G-sample table read-sharded over the ranks: STRONG scaling of decode and encode), the range coder.
`--impl reference` times only the CPU implementation (all host cores, bounded sample per step).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

SEED = 20261018          # R10.4.1-shaped, 100k reads
SEED_R9 = 20261019       # config (4): R9.4.1-shaped, same read count
SEED_5 = 20261020        # config (5): 6.0 G samples, R10.4.1-shaped, read-sharded
METRIC = "int16 signal GB/s, pgnano batched decode (bit-exact)"
UNIT = "GB/s"


def log(*a):
    print("[bench]", *a, file=sys.stderr, flush=True)


# ------------------------------------------------------------------------------------------------------
# CPU arm: the reference's own decoder (oracle/_ref) or the C port (oracle/liboracle.so), one PROCESS per core
# (the reference mutates unsynchronised global counters, C5.hpp:27-48, so threads would share cache lines).
# ------------------------------------------------------------------------------------------------------
def _cpu_worker(kind, codec, mode, comp, coff, sig, off, n, idx, barrier, t_budget, q):
    """mode 'decode': chunk bytes -> samples; mode 'encode': samples -> chunk bytes. The reference's functions
    (oracle/_ref: pgnano::compress_signal_N01 / decompress_signal_N01, pod5::compress_signal / decompress_signal,
    pgnano::Compressor / Decompressor) or the C port, called once per chunk as src/c++/benchmark.cpp:144-146,194-223 does."""
    import ctypes as C
    from oracle_binding import Oracle, RefOracle
    S = C.c_size_t
    nmax = int(n.max()) if len(n) else 0
    out16 = np.empty(nmax + 64, np.int16); o16 = C.c_void_p(out16.ctypes.data)
    out8 = np.empty(4 * nmax + 4096, np.uint8); o8 = C.c_void_p(out8.ctypes.data)
    sz = S(0)
    if kind == "reference":
        lib = RefOracle().lib
        lib.ref_rc_encode.restype = C.c_size_t
        if mode == "decode":
            if codec == 0: fn = lambda p, cl, x, k: lib.ref_vbz_decode(p, S(cl), S(k), o16)
            elif codec == 7: fn = lambda p, cl, x, k: (lib.ref_rc_decode(p, S(cl), o16), 0)[1]
            else: fn = lambda p, cl, x, k: lib.ref_cx_decode(codec, p, S(cl), S(k), o16)
        else:
            if codec == 0: fn = lambda p, cl, x, k: lib.ref_vbz_encode(x, S(k), o8, S(out8.size), C.byref(sz))
            elif codec == 7: fn = lambda p, cl, x, k: 0 if lib.ref_rc_encode(x, S(k), o8, b"R10.4.1") else 1
            else: fn = lambda p, cl, x, k: lib.ref_cx_encode(codec, x, S(k), o8, S(max(2 * k + 26, 1024)), C.byref(sz))
    else:
        lib = Oracle().lib
        lib.orc_rc_encode.restype = C.c_size_t
        if mode == "decode":
            if codec == 7: fn = lambda p, cl, x, k: lib.orc_rc_decode(p, S(cl), S(k), o16)
            else: fn = lambda p, cl, x, k: lib.orc_decode(codec, p, S(cl), S(k), o16, 0)
        else:
            if codec == 7: fn = lambda p, cl, x, k: 0 if lib.orc_rc_encode(x, S(k), o8, C.c_uint8(0)) else 1
            else: fn = lambda p, cl, x, k: lib.orc_encode(codec, x, S(k), o8, S(out8.size), C.byref(sz))
    cbase = comp.ctypes.data if comp is not None else 0
    sbase = sig.ctypes.data
    args = [(C.c_void_p(cbase + int(coff[i])) if comp is not None else None, int(coff[i + 1] - coff[i]) if comp is not None else 0,
             C.c_void_p(sbase + 2 * int(off[i])), int(n[i])) for i in idx]
    barrier.wait()
    t0 = time.perf_counter()
    samples = 0
    bad = 0
    while True:
        for p, cl, x, k in args:
            bad |= fn(p, cl, x, k)
            samples += k
        if time.perf_counter() - t0 >= t_budget or not args:
            break
    q.put((samples, time.perf_counter() - t0, bad))


def cpu_rate(kind, codec, mode, smp, cores, seconds):
    """Aggregate int16 GB/s of the CPU codec over `cores` processes, each looping over its static shard of the
    sample for about `seconds`. Returns (GB/s, samples processed, wall seconds)."""
    import multiprocessing as mp
    ctx = mp.get_context("fork")
    barrier = ctx.Barrier(cores + 1)
    q = ctx.Queue()
    n = smp["n"]
    comp, coff = smp["comp"].get(codec, (None, None))
    order = np.argsort(-n.astype(np.int64), kind="stable")
    shards = [order[w::cores] for w in range(cores)]
    procs = [ctx.Process(target=_cpu_worker, args=(kind, codec, mode, comp, coff, smp["sig"], smp["off"], n, shards[w], barrier, seconds, q)) for w in range(cores)]
    for p in procs:
        p.start()
    barrier.wait()
    t0 = time.perf_counter()
    res = [q.get() for _ in procs]
    wall = time.perf_counter() - t0
    for p in procs:
        p.join()
    if any(r[2] for r in res):
        raise RuntimeError(f"CPU codec reported an error (codec {codec}, {mode})")
    total = sum(r[0] for r in res)
    # every worker ran for its own measured time; aggregate rate = sum of the per-process rates
    rate = sum(2.0 * r[0] / r[1] for r in res) / 1e9
    return rate, total, wall


def cpu_sample(codecs, nreads, seed, model, max_chunks=None):
    """Bounded sample of a workload built WITHOUT the GPU: the first `nreads` reads, encoded by the CPU oracle (the
    reference's own encoder when oracle/_ref is built) for every codec in `codecs`. Keeps the reference bytes per chunk:
    they are also what the product encoder's output is compared with (BASELINE config 3)."""
    from rawnanoporesignalcompression_b200 import synth
    from oracle_binding import Oracle, RefOracle
    lengths = synth.read_lengths(nreads, seed)
    sig = synth.generate_reads(lengths, seed, model).numpy()
    off, n, _ = synth.chunk_table(lengths)
    if max_chunks is not None:
        off, n = off[:max_chunks], n[:max_chunks]
    enc = RefOracle() if RefOracle.available() else Oracle()
    comp = {}
    for codec in codecs:
        parts = []
        for o, c in zip(off, n):
            st, b = enc.encode_status(codec, sig[o : o + c])
            if st:
                raise RuntimeError("oracle encode failed")
            parts.append(b)
        coff = np.zeros(len(parts) + 1, np.int64)
        np.cumsum([len(b) for b in parts], out=coff[1:])
        comp[codec] = (np.frombuffer(b"".join(parts) + b"\0" * 64, np.uint8).copy(), coff)
    return dict(sig=sig, off=off, n=n, comp=comp, samples=int(n.sum()), lengths=lengths)


def cpu_kind():
    from oracle_binding import RefOracle
    return "reference" if RefOracle.available() else "port"


# ------------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index = index; self.proc = None; self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True); self.th.start()
        except OSError:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "samples": len(sm), "reasons": sorted(reasons)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------------
def build_workload(cd, torch, dev, codec, lengths, first_read, model, seed, shard_reads=4096, keep_comp=True):
    """Generates reads [first_read, first_read + len(lengths)) on the GPU (counter-based, reproducible anywhere), encodes
    them with the product encoder (device splitter + host libzstd frame stage, byte-identical to the reference) and leaves
    samples + compressed chunks resident in HBM. t_enc = wall time of the rnsc_compress_batch calls (pageable host in/out)."""
    from rawnanoporesignalcompression_b200 import synth
    nreads = len(lengths)
    off, n, _ = synth.chunk_table(lengths)
    total = int(lengths.sum())
    d_sig = torch.empty(total, dtype=torch.int16, device=dev)
    per_read = (lengths + synth.DEFAULT_CHUNK_SIZE - 1) // synth.DEFAULT_CHUNK_SIZE
    cfirst = np.cumsum(per_read) - per_read
    rstart = np.cumsum(lengths) - lengths
    comp_parts, csize = [], []
    t_gen = t_enc = 0.0
    for a in range(0, nreads, shard_reads):
        b = min(nreads, a + shard_reads)
        s0 = int(rstart[a]); s1 = int(rstart[b - 1] + lengths[b - 1])
        t0 = time.perf_counter()
        synth.generate_reads(lengths[a:b], seed, model, first_read=first_read + a, device=dev, out=d_sig[s0:s1])
        host = d_sig[s0:s1].cpu().numpy()
        t_gen += time.perf_counter() - t0
        c0 = int(cfirst[a]); c1 = int(cfirst[b - 1] + per_read[b - 1])
        t0 = time.perf_counter()
        out, coff, st = cd.compress_flat(host, (off[c0:c1] - s0).astype(np.uint64), n[c0:c1], codec)
        t_enc += time.perf_counter() - t0
        if st.any():
            raise RuntimeError("encoder reported a per-chunk error while building the workload")
        comp_parts.append(out.copy()); csize.append(np.diff(coff.astype(np.int64)))
    csize = np.concatenate(csize) if csize else np.zeros(0, np.int64)
    coff = np.zeros(csize.size + 1, np.int64)
    np.cumsum(csize, out=coff[1:])
    comp = np.concatenate(comp_parts + [np.zeros(64, np.uint8)])
    log(f"workload (seed {seed}, {model.name}, codec {codec}): {nreads} reads from read {first_read}, {n.size} chunks, {total/1e9:.3f} G samples, "
        f"{coff[-1]/1e9:.3f} GB compressed ({8.0*coff[-1]/max(total,1):.4f} bits/sample); generate {t_gen:.1f}s, encode {t_enc:.1f}s "
        f"({2e-9*total/max(t_enc,1e-9):.2f} GB/s incl. PCIe + host zstd)")
    return dict(lengths=lengths, off=off, n=n, total=total, d_sig=d_sig, comp=comp, coff=coff, t_enc=t_enc, encode_gbs=2e-9 * total / max(t_enc, 1e-9))


def table_checksum(torch, x, first_index):
    """Order-sensitive 64-bit checksum of a slice of the signal table: sum over samples of (uint16 sample + 1) x a
    SplitMix64 weight of its GLOBAL sample index, mod 2^64. Sums of slices add up to the checksum of the whole table, so
    the ranks of a read-sharded run can prove together (one all-reduce of 8 bytes) that they decoded the same table."""
    from rawnanoporesignalcompression_b200.synth import _mix
    acc = torch.zeros((), dtype=torch.int64, device=x.device)
    step = 1 << 27
    for a in range(0, x.numel(), step):
        v = x[a : a + step].to(torch.int64) & 0xFFFF
        idx = torch.arange(first_index + a, first_index + a + v.numel(), dtype=torch.int64, device=x.device)
        acc += ((v + 1) * (_mix(idx) | 1)).sum()
    return acc


def device_timed(torch, fn, steps, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def run_rnsc(args):
    import torch
    import torch.distributed as dist
    from rawnanoporesignalcompression_b200 import synth, sharding
    from rawnanoporesignalcompression_b200.codec import Codec, codec_id, _ptr
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    codec = codec_id(args.codec)
    model = synth.MODELS[args.model]
    ncores = os.cpu_count() or 1
    if world > 1:      # each rank keeps to its own slice of the host cores (staging threads, zstd frame stage)
        try:
            per = max(1, ncores // world)
            os.sched_setaffinity(0, set(range(local * per, min(ncores, (local + 1) * per))))
        except (AttributeError, OSError):
            pass

    # ---- CPU baselines first (rank 0, N=1), before CUDA is initialised in this process (fork safety) ----
    cpu = None; cpu_legs = {}; ref_c5 = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        kind = cpu_kind()
        smp = cpu_sample([codec], args.cpu_reads, SEED, model)
        ref_c5 = (smp["comp"][codec], smp["n"])                      # the reference encoder's bytes of the first chunks (config 3 check)
        rate, done, wall = cpu_rate(kind, codec, "decode", smp, ncores, args.cpu_seconds)
        rate1, _, _ = cpu_rate(kind, codec, "decode", smp, 1, min(args.cpu_seconds, 4.0))
        desc = f"first {args.cpu_reads} reads of the workload ({smp['samples']} samples, {len(smp['n'])} chunks)"
        cpu = {"value": round(rate, 3), "unit": UNIT, "cores": ncores, "kind": kind, "one_core_value": round(rate1, 4),
               "sample": f"{desc}, each of {ncores} processes loops over its static shard for ~{args.cpu_seconds:.0f} s; {done} samples decoded in {wall:.1f} s wall"}
        log("cpu_baseline", cpu)
        # the other legs of the path, same sample, a few seconds each (src/c++/benchmark.cpp:144-146 times compress as well)
        t_leg = max(2.0, args.cpu_seconds / 4)
        for name, cdc, mode in ((f"{args.codec}_encode", codec, "encode"),):
            r, _, _ = cpu_rate(kind, cdc, mode, smp, ncores, t_leg)
            cpu_legs[name] = {"value": round(r, 3), "unit": UNIT, "cores": ncores, "kind": kind, "sample": desc}
        # VBZ legs on the R9.4.1-shaped set (config 4), range coder legs (benchmark.cpp:194-223) on a smaller sample (~30 MB/s per core)
        smp9 = cpu_sample([0], max(16, args.cpu_reads // 2), SEED_R9, synth.R9_4_1)
        d9 = f"first {max(16, args.cpu_reads // 2)} reads of the R9.4.1-shaped set (seed {SEED_R9}; {smp9['samples']} samples, {len(smp9['n'])} chunks)"
        for name, mode in (("VBZ_decode_R9.4.1", "decode"), ("VBZ_encode_R9.4.1", "encode")):
            r, _, _ = cpu_rate(kind, 0, mode, smp9, ncores, t_leg)
            cpu_legs[name] = {"value": round(r, 3), "unit": UNIT, "cores": ncores, "kind": kind, "sample": d9}
        smpr = cpu_sample([7], max(8, args.cpu_reads // 8), SEED, model)
        dr = f"first {max(8, args.cpu_reads // 8)} reads of the workload ({smpr['samples']} samples, {len(smpr['n'])} chunks)"
        for name, mode in (("RC_encode", "encode"), ("RC_decode", "decode")):
            r, _, _ = cpu_rate(kind, 7, mode, smpr, ncores, t_leg)
            cpu_legs[name] = {"value": round(r, 3), "unit": UNIT, "cores": ncores, "kind": kind, "sample": dr}
        log("cpu legs", cpu_legs)
        del smp9, smpr

    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: librnsc has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cd = Codec(local)
    if args.host_threads:
        cd.set_host_threads(args.host_threads)
    elif world > 1:
        cd.set_host_threads(max(1, ncores // world))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def allreduce(v, op="sum"):
        t = torch.tensor([float(v)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX if op == "max" else (dist.ReduceOp.MIN if op == "min" else dist.ReduceOp.SUM))
        return float(t.item())

    lengths = synth.read_lengths(args.reads, SEED, first_read=rank * args.reads)
    wl = build_workload(cd, torch, dev, codec, lengths, rank * args.reads, model, SEED)
    n, off, coff, total = wl["n"], wl["off"], wl["coff"], wl["total"]
    nch = n.size
    cbytes = int(coff[-1])
    # BASELINE config 3: the product encoder's bytes against the reference encoder's, chunk by chunk, on the CPU sample
    enc_check = None
    if ref_c5 is not None:
        (rc_comp, rc_off), rn = ref_c5
        k = min(len(rn), nch)
        same = all(int(coff[i + 1] - coff[i]) == int(rc_off[i + 1] - rc_off[i]) and
                   bytes(wl["comp"][int(coff[i]) : int(coff[i + 1])]) == bytes(rc_comp[int(rc_off[i]) : int(rc_off[i + 1])]) for i in range(k))
        enc_check = {"chunks_compared": int(k), "bytes_identical_to_reference": bool(same)}
        log("encode byte parity vs the reference encoder:", enc_check)
    d_comp = torch.from_numpy(wl["comp"]).to(dev)
    d_coff = torch.from_numpy(coff).to(dev)
    d_n = torch.from_numpy(n.astype(np.int32)).to(dev)
    d_off = torch.from_numpy(off.astype(np.int64)).to(dev)
    d_out = torch.empty(total, dtype=torch.int16, device=dev)
    d_status = torch.empty(nch, dtype=torch.int32, device=dev)
    max_n = int(n.max())

    def step():
        cd.dev_decode(codec, d_comp, d_coff, d_n, d_off, max_n, d_out, d_status)

    if args.phases:
        cd.set_profiling(True)
        for _ in range(3):
            step()
        for cls, ph in cd.phase_cycles().items():
            nchk = max(1, ph["chunks"])
            log(f"phases {cls}: chunks {ph['chunks']//3}/step; cycles per chunk: " + ", ".join(f"{k} {v/nchk:.0f}" for k, v in ph.items() if k != "chunks"))
        cd.set_profiling(False)
    sampler = ClockSampler(local); sampler.start()     # clocks / throttle reasons over warm-up + timed region (all under load)
    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    bit_exact = bool(torch.equal(d_out, wl["d_sig"])) and not bool(d_status.any().item())
    classes = cd.decode_class_counts(torch.cuda.current_stream())
    log(f"chunk routing (fast 4/SM, fast 2/SM, fast 1/SM, general): {classes}")
    log(f"decode bit-exact vs generated samples: {bit_exact}")

    barrier()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    l0 = cd.launch_count
    ev[0].record()
    for i in range(args.steps):
        step()
        ev[i + 1].record()
    barrier()
    launches = cd.launch_count - l0
    clocks = sampler.stop()
    ms_total = ev[0].elapsed_time(ev[-1])
    kern_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
    ms_step = allreduce(ms_total, "max") / args.steps
    all_samples = allreduce(total); all_cbytes = allreduce(cbytes)
    value = 2.0 * all_samples / (ms_step * 1e-3) / 1e9

    # ---- roofline of the dominant launch group (this rank's launch: C + 2N algorithmic bytes) ----
    peak, peak_src = measured_peak()
    k_ms = float(np.mean(kern_ms))
    achieved = (cbytes + 2.0 * total) / (k_ms * 1e-3) / 1e9
    roof = {"bound": "hbm", "kernel": "rnsc_dev_decode: k_plan_chunks, then k_decode_fast<1024|512|256> (dominant: <1024>) and k_decode_chunks concurrently (one step = one such launch group)",
            "achieved": round(achieved, 2), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 5),
            "traffic": None, "peak_source": peak_src, "algorithmic_bytes_per_launch": int(cbytes + 2 * total), "avg_launch_ms": round(k_ms, 4)}
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):      # dram__bytes_read+write of the launch group from an `ncu --set full` capture, per algorithmic byte
        try:
            tj = json.load(open(tp))
            roof["traffic"] = int(round(tj["dram_bytes_per_algorithmic_byte"] * (cbytes + 2.0 * total)))
            roof["traffic_source"] = tj.get("source")
        except Exception:
            pass

    # ---- end to end through the C ABI with pinned host buffers ----
    e2e = None
    if not args.no_e2e:
        import psutil
        e_reads = args.e2e_reads or args.reads
        # bound pinned memory to a quarter of the host RAM shared by the ranks of this box
        budget = psutil.virtual_memory().total / 4 / world
        per_read = (cbytes + 2.0 * total) / args.reads
        e_reads = int(max(1, min(e_reads, budget // per_read)))
        per = (wl["lengths"] + synth.DEFAULT_CHUNK_SIZE - 1) // synth.DEFAULT_CHUNK_SIZE
        e_ch = int(per[:e_reads].sum()); e_samples = int(wl["lengths"][:e_reads].sum()); e_cb = int(coff[e_ch])
        h_comp = torch.empty(e_cb + 64, dtype=torch.uint8).pin_memory()
        h_comp[: e_cb + 64].copy_(torch.from_numpy(wl["comp"][: e_cb + 64]))
        h_out = torch.empty(e_samples, dtype=torch.int16).pin_memory()
        hc, ho = h_comp.numpy(), h_out.numpy()
        e_coff = coff[: e_ch + 1].astype(np.uint64); e_n = n[:e_ch].astype(np.uint32)
        e_soff = off[:e_ch].astype(np.uint64)
        status = np.zeros(e_ch, np.int32)

        def e2e_step():
            r = cd.lib.rnsc_decompress_batch(cd.ctx, codec, _ptr(hc), _ptr(e_coff), _ptr(e_n), e_ch, _ptr(ho), _ptr(e_soff), _ptr(status))
            if r or status.any():
                raise RuntimeError("rnsc_decompress_batch failed in the e2e leg")

        for _ in range(max(args.warmup, 3)):
            e2e_step()
        e_steps = max(1, args.steps)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e_steps):
            e2e_step()
        barrier()
        dt = allreduce(time.perf_counter() - t0, "max")
        es = allreduce(e_samples)
        e_ok = bool((h_out == wl["d_sig"][:e_samples].cpu()).all())
        # context for the e2e number: what the PCIe link does with ALL ranks copying at the same time (pinned 1 GiB
        # copies, best of 3, every rank between the same two barriers), per rank and summed over the box
        pcie = {}
        try:
            nb = min(1 << 30, h_out.numel() * 2)
            hb = h_out[: nb // 2]; db = torch.empty(nb // 2, dtype=torch.int16, device=dev)
            for name, (dst_t, src_t) in (("h2d", (db, hb)), ("d2h", (hb, db))):
                best = 0.0
                for _ in range(3):
                    barrier(); t1 = time.perf_counter(); dst_t.copy_(src_t, non_blocking=True); torch.cuda.synchronize()
                    best = max(best, nb / (time.perf_counter() - t1) / 1e9)
                pcie[name + "_GBps"] = round(best, 2)
            # both directions at once (what the pipelined call actually asks of the link): second device buffer, two streams
            db2 = torch.empty(nb // 2, dtype=torch.int16, device=dev); hb2 = torch.empty(nb // 2, dtype=torch.int16).pin_memory()
            sa, sb_ = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
            best = 0.0
            for _ in range(3):
                barrier(); t1 = time.perf_counter()
                with torch.cuda.stream(sa):
                    db.copy_(hb, non_blocking=True)
                with torch.cuda.stream(sb_):
                    hb2.copy_(db2, non_blocking=True)
                torch.cuda.synchronize()
                best = max(best, nb / (time.perf_counter() - t1) / 1e9)
            pcie["duplex_each_way_GBps"] = round(best, 2)
            del db2, hb2, db
        except Exception as ex:  # noqa: BLE001
            pcie = {"error": str(ex)}
        e_val = 2.0 * es * e_steps / dt / 1e9
        if "error" not in pcie:
            agg = {k: round(allreduce(v), 2) for k, v in pcie.items()}
            low = {k: round(allreduce(v, "min"), 2) for k, v in pcie.items()}
            pcie = {"ranks_copying_concurrently": world, "per_rank_min": low, "box_aggregate": agg, "rank0": pcie}
            # the decoder returns 2 bytes per sample over D2H while H2D carries the compressed bytes: with both directions busy
            # the box cannot deliver more int16 GB/s than its aggregate each-way duplex rate
            bound = agg["duplex_each_way_GBps"]
            pcie["e2e_over_duplex_bound"] = round(e_val / bound, 3) if bound else None
            pcie["limiter"] = ("the box's aggregate PCIe / host-memory bandwidth with all ranks copying (e2e is within 15% of the concurrent duplex rate)"
                               if bound and e_val >= 0.85 * bound else "the staging pipeline of rnsc_decompress_batch (e2e is more than 15% below the concurrent duplex rate)")
        e2e = {"value": round(e_val, 3), "unit": UNIT, "h2d_bytes_per_step": int(e_cb + 20 * e_ch + 16),
               "d2h_bytes_per_step": int(2 * e_samples + 4 * e_ch), "steps": e_steps, "reads_per_rank": e_reads, "bit_exact": e_ok, "pcie_link": pcie,
               "call": "rnsc_decompress_batch (pinned host buffers in and out)"}
        # The other direction: whole-chunk ENCODE end to end through rnsc_compress_batch with the GPU entropy stage (pinned host
        # buffers: H2D of the samples, splitter + entropy stage on the device, D2H of the finished chunks). The decode leg's
        # output buffer holds the samples; the result is checked by decoding it on the device.
        try:
            del h_comp
            enc_cap = int(e_cb * 1.03) + 1104 * 64 + 65536
            h_enc = torch.empty(enc_cap, dtype=torch.uint8).pin_memory(); he = h_enc.numpy()
            coff2 = np.zeros(e_ch + 1, np.uint64); status2 = np.zeros(e_ch, np.int32)
            cd.set_entropy_stage("gpu")

            def e2e_enc_step():
                r = cd.lib.rnsc_compress_batch(cd.ctx, codec, _ptr(ho), _ptr(e_soff), _ptr(e_n), e_ch, _ptr(he), enc_cap, _ptr(coff2), _ptr(status2))
                if r or status2.any():
                    raise RuntimeError("rnsc_compress_batch (GPU entropy stage) failed in the e2e leg")

            for _ in range(2):
                e2e_enc_step()
            barrier()
            t0 = time.perf_counter()
            for _ in range(e_steps):
                e2e_enc_step()
            barrier()
            dt2 = allreduce(time.perf_counter() - t0, "max")
            cd.set_entropy_stage("libzstd")
            g_b = int(coff2[-1])
            d_enc = torch.from_numpy(he[: g_b + 64].copy()).to(dev); d_c2 = torch.from_numpy(coff2.astype(np.int64)).to(dev)
            d_o2 = torch.empty(e_samples, dtype=torch.int16, device=dev); d_s2 = torch.empty(e_ch, dtype=torch.int32, device=dev)
            cd.dev_decode(codec, d_enc, d_c2, d_n[:e_ch], d_off[:e_ch], max_n, d_o2, d_s2)
            torch.cuda.synchronize()
            enc_ok = bool(torch.equal(d_o2, wl["d_sig"][:e_samples])) and not bool(d_s2.any().item())
            e2e["encode_gpu_entropy"] = {"value": round(2.0 * es * e_steps / dt2 / 1e9, 3), "unit": UNIT, "h2d_bytes_per_step": int(2 * e_samples + 12 * e_ch + 8),
                                         "d2h_bytes_per_step": int(g_b + 12 * e_ch + 8), "bits_per_sample": round(8.0 * g_b / e_samples, 4),
                                         "roundtrip_bit_exact": enc_ok, "call": "rnsc_compress_batch, RNSC_ENTROPY_GPU (pinned host buffers in and out)"}
            del h_enc, d_enc, d_o2
        except Exception as ex:  # noqa: BLE001
            cd.set_entropy_stage("libzstd")
            e2e["encode_gpu_entropy"] = {"error": str(ex)}
        del h_out

    # ---- other legs of the path (not the headline) -------------------------------------------------------------
    extra = {}
    if not args.no_extra:
        x_reads = min(args.reads, args.extra_reads)
        per = (wl["lengths"] + synth.DEFAULT_CHUNK_SIZE - 1) // synth.DEFAULT_CHUNK_SIZE
        x_ch = int(per[:x_reads].sum()); x_samples = int(wl["lengths"][:x_reads].sum())
        xs = wl["d_sig"][:x_samples]

        def splitter_leg(cid, sig, d_off_, d_n_, n_, nsamp):
            cap = int(sum(cd.lib.rnsc_max_stream_bytes(cid, int(k)) for k in n_))
            d_streams = torch.empty(cap, dtype=torch.uint8, device=dev)
            d_soff = torch.empty(5 * len(n_) + 1, dtype=torch.int64, device=dev)
            ms = device_timed(torch, lambda: cd.dev_svb_encode(cid, sig, d_off_, d_n_, d_streams, d_soff), 3, warm=1)
            sbytes = int(d_soff[-1].item())
            leg = {"value": round(2e-6 * nsamp / ms, 2), "unit": UNIT, "stream_bytes_per_sample": round(sbytes / nsamp, 4),
                   "roofline_frac": round((2.0 * nsamp + sbytes) / (ms * 1e-3) / 1e9 / peak, 4),
                   "what": "rnsc_dev_svb_encode: dense streams (size pass + scan + scatter: the samples are read twice)"}
            del d_streams
            # one pass over the samples into per-chunk slabs (what rnsc_dev_encode runs in front of its entropy stage)
            d_slabs = torch.empty(4 * nsamp + 128 * len(n_) + 64, dtype=torch.uint8, device=dev)
            d_start = torch.empty(5 * len(n_), dtype=torch.int64, device=dev); d_size = torch.empty(5 * len(n_), dtype=torch.int32, device=dev)
            ms1 = device_timed(torch, lambda: cd.dev_svb_encode_slabs(cid, sig, d_off_, d_n_, nsamp, d_slabs, d_start, d_size), 3, warm=1)
            same = int(d_size.sum().item()) == sbytes
            leg["one_pass_slabs"] = {"value": round(2e-6 * nsamp / ms1, 2), "unit": UNIT, "roofline_frac": round((2.0 * nsamp + sbytes) / (ms1 * 1e-3) / 1e9 / peak, 4),
                                     "same_stream_bytes": same, "what": "rnsc_dev_svb_encode_slabs: one read of the samples, stream start + size per stream"}
            return leg

        # BASELINE config 3: pgnano encode. Device-resident splitter stage (int16 -> streams) on a sample, and the whole-chunk
        # encode through rnsc_compress_batch (host buffers; device splitter + host libzstd level-1 frame stage = byte-identical
        # to the reference, see byte_parity), timed while the workload was built.
        extra[f"{args.codec}_splitter_encode"] = splitter_leg(codec, xs, d_off[:x_ch], d_n[:x_ch], n[:x_ch], x_samples)
        enc_t = allreduce(wl["t_enc"], "max")
        extra["config3_encode"] = {"value": round(2e-9 * all_samples / enc_t, 3), "unit": UNIT, "codec": args.codec,
                                   "what": "whole-chunk encode of every read of the workload through rnsc_compress_batch (pageable host buffers in and out, H2D + device "
                                           "splitter + D2H + host libzstd level 1 on this rank's cores); aggregate over ranks, max time",
                                   "byte_parity": enc_check, "bits_per_sample": round(8.0 * all_cbytes / all_samples, 4),
                                   "cpu_baseline": cpu_legs.get(f"{args.codec}_encode")}
        extra["sample"] = f"splitter leg: first {x_reads} reads ({x_samples} samples, {x_ch} chunks)"
        log("config 3", extra["config3_encode"])
        # Device-resident whole-chunk encode with the GPU entropy stage (csrc/zstd_encode.cuh): valid zstd frames written on the
        # device -- decode parity with the reference (its decoder returns the samples), not byte parity with libzstd.
        try:
            g_cap = 2 * total + total // 4 + 1104 * nch + 64
            g_comp = torch.empty(g_cap, dtype=torch.uint8, device=dev); g_coff = torch.empty(nch + 1, dtype=torch.int64, device=dev)
            g_st = torch.empty(nch, dtype=torch.int32, device=dev)
            barrier()
            ms_g = device_timed(torch, lambda: cd.dev_encode(codec, wl["d_sig"], d_off, d_n, total, g_comp, g_coff, g_st), max(3, args.steps), warm=2)
            ms_g = allreduce(ms_g, "max")
            g_bytes = int(g_coff[-1].item())
            g_ok = not bool(g_st.any().item())
            cd.dev_decode(codec, g_comp, g_coff, d_n, d_off, max_n, d_out, d_status)         # this library's decoder on the device-written frames
            torch.cuda.synchronize()
            g_rt = bool(torch.equal(d_out, wl["d_sig"])) and not bool(d_status.any().item())
            ref_checked = 0; ref_ok = None
            if rank == 0:                                                                     # the reference's own decoder on a sample of the chunks
                from oracle_binding import Oracle, RefOracle
                chk = RefOracle() if RefOracle.available() else Oracle()
                k = min(nch, args.cpu_reads)
                hc_ = g_comp[: int(g_coff[k].item()) + 64].cpu().numpy(); hco = g_coff[: k + 1].cpu().numpy()
                hs = wl["d_sig"][: int(off[k - 1] + n[k - 1])].cpu().numpy()
                ref_ok = True
                for i in range(k):
                    r, y = chk.decode_status(codec, hc_[int(hco[i]) : int(hco[i + 1])].tobytes(), int(n[i]))
                    ref_ok = ref_ok and r == 0 and bool((y == hs[off[i] : off[i] + n[i]]).all())
                ref_checked = k
            extra["config3_encode_gpu_entropy"] = {
                "value": round(2e-6 * all_samples / ms_g, 2), "unit": UNIT, "ms_per_step": round(ms_g, 3), "codec": args.codec,
                "what": "device-resident whole-chunk encode of the whole workload: splitter + GPU entropy stage (raw / RLE / Huffman-literal zstd blocks written on the device) "
                        "+ compaction; samples and finished chunks stay in HBM",
                "parity": "decode parity: frames are valid Zstandard and decode to the input samples; bytes differ from libzstd level 1 (use the host-libzstd mode for byte parity)",
                "bits_per_sample": round(8.0 * allreduce(g_bytes) / all_samples, 4), "bits_per_sample_host_libzstd": round(8.0 * all_cbytes / all_samples, 4),
                "all_chunks_ok": g_ok, "roundtrip_bit_exact_all_chunks": g_rt, "reference_decoder_chunks_checked": ref_checked, "reference_decoder_ok": ref_ok,
                "roofline_frac": round((2.0 * total + g_bytes) / (ms_g * 1e-3) / 1e9 / peak, 4), "cpu_baseline": cpu_legs.get(f"{args.codec}_encode")}
            log("config 3 (GPU entropy stage)", extra["config3_encode_gpu_entropy"])
            del g_comp, g_coff, g_st
        except Exception as ex:  # noqa: BLE001
            extra["config3_encode_gpu_entropy"] = {"error": str(ex)}

    # free the headline set before the other configurations are built
    del d_comp, d_out, d_status, d_coff, d_n, d_off
    d_sig0 = wl.pop("d_sig"); del d_sig0
    wl = None
    torch.cuda.empty_cache()

    if not args.no_extra:
        # ---- BASELINE config 4: VBZ batched decode + StreamVByte-stage encode on the R9.4.1-shaped set (seed 20261019) ----
        r9_reads = min(args.reads, args.vbz_reads) if args.vbz_reads else args.reads
        l9 = synth.read_lengths(r9_reads, SEED_R9, first_read=rank * r9_reads)
        w9 = build_workload(cd, torch, dev, 0, l9, rank * r9_reads, synth.R9_4_1, SEED_R9)
        v_comp = torch.from_numpy(w9["comp"]).to(dev); v_coff = torch.from_numpy(w9["coff"]).to(dev)
        v_n = torch.from_numpy(w9["n"].astype(np.int32)).to(dev); v_off = torch.from_numpy(w9["off"].astype(np.int64)).to(dev)
        v_out = torch.empty(w9["total"], dtype=torch.int16, device=dev); v_st = torch.empty(w9["n"].size, dtype=torch.int32, device=dev)
        vmax = int(w9["n"].max())
        barrier()
        ms = device_timed(torch, lambda: cd.dev_decode(0, v_comp, v_coff, v_n, v_off, vmax, v_out, v_st), max(3, args.steps), warm=3)
        ms = allreduce(ms, "max"); v_all = allreduce(w9["total"]); v_cb = allreduce(int(w9["coff"][-1]))
        v_ok = bool(torch.equal(v_out, w9["d_sig"])) and not bool(v_st.any().item())
        extra["config4_VBZ_decode_R9.4.1"] = {"value": round(2e-6 * v_all / ms, 2), "unit": UNIT, "ms_per_step": round(ms, 3), "reads_per_gpu": r9_reads,
                                              "bits_per_sample": round(8.0 * v_cb / v_all, 4), "bit_exact": v_ok, "routing": cd.decode_class_counts(torch.cuda.current_stream()),
                                              "roofline_frac": round((w9["coff"][-1] + 2.0 * w9["total"]) / (ms * 1e-3) / 1e9 / peak, 4),
                                              "cpu_baseline": cpu_legs.get("VBZ_decode_R9.4.1")}
        del v_comp, v_out
        leg = splitter_leg(0, w9["d_sig"], v_off, v_n, w9["n"], w9["total"])
        leg["cpu_baseline"] = cpu_legs.get("VBZ_encode_R9.4.1")
        leg["whole_chunk_encode_host_zstd"] = round(2e-9 * v_all / allreduce(w9["t_enc"], "max"), 3)
        extra["config4_VBZ_splitter_encode_R9.4.1"] = leg
        log("config 4", extra["config4_VBZ_decode_R9.4.1"], leg)
        del w9, v_coff, v_n, v_off, v_st
        torch.cuda.empty_cache()

        # ---- range coder (benchmark.cpp:194-223): device-resident encode + decode of a sample, one warp per chunk ----
        rc_reads = min(args.reads, args.rc_reads)
        lr = synth.read_lengths(rc_reads, SEED, first_read=rank * args.reads)
        r_off, r_n, _ = synth.chunk_table(lr)
        r_sig = synth.generate_reads(lr, SEED, model, first_read=rank * args.reads, device=dev)
        slot = np.zeros(r_n.size + 1, np.int64)
        np.cumsum([(cd.max_compressed_size(7, int(k)) + 24 + 15) & ~15 for k in r_n], out=slot[1:])
        r_comp = torch.empty(int(slot[-1]) + 64, dtype=torch.uint8, device=dev)
        r_slot = torch.from_numpy(slot).to(dev); r_dn = torch.from_numpy(r_n.astype(np.int32)).to(dev); r_doff = torch.from_numpy(r_off.astype(np.int64)).to(dev)
        r_size = torch.empty(r_n.size, dtype=torch.int32, device=dev); r_st = torch.empty(r_n.size, dtype=torch.int32, device=dev)
        ms_e = device_timed(torch, lambda: cd.dev_rc_encode(r_sig, r_doff, r_dn, 0, r_comp, r_slot, r_size, r_st), 2, warm=1)
        ok_e = not bool(r_st.any().item())
        # decode straight from the slots: comp_off[i] = slot[i], the decoder reads the chunk's own length from its header
        r_out = torch.empty_like(r_sig)
        ms_d = device_timed(torch, lambda: cd.dev_rc_decode(r_comp, r_slot, r_dn, r_doff, r_out, r_st), 2, warm=1)
        ok_d = bool(torch.equal(r_out, r_sig)) and not bool(r_st.any().item())
        rs = int(lr.sum())
        extra["RC_encode"] = {"value": round(2e-6 * rs / ms_e, 2), "unit": UNIT, "ok": ok_e, "bits_per_sample": round(8.0 * int(r_size.sum().item()) / rs, 4), "cpu_baseline": cpu_legs.get("RC_encode")}
        extra["RC_decode"] = {"value": round(2e-6 * rs / ms_d, 2), "unit": UNIT, "bit_exact": ok_d, "cpu_baseline": cpu_legs.get("RC_decode"),
                              "sample": f"first {rc_reads} reads per GPU ({rs} samples, {r_n.size} chunks)"}
        log("range coder", extra["RC_encode"], extra["RC_decode"])
        del r_sig, r_comp, r_out
        torch.cuda.empty_cache()

        # ---- BASELINE config 5: the 12 GB table (6.0 G samples, seed 20261020), READ-SHARDED over the ranks: strong scaling ----
        tot5 = int(6.0e9 * args.reads / 100_000)
        l5 = synth.read_lengths(int(args.reads * 1.5) + 64, SEED_5)
        n5 = int(np.searchsorted(np.cumsum(l5), tot5, side="left")) + 1
        l5 = l5[:n5]
        bounds = sharding.partition_reads(l5, world)
        r0, r1 = int(bounds[rank]), int(bounds[rank + 1])
        first_sample = int(l5[:r0].sum())
        w5 = build_workload(cd, torch, dev, codec, l5[r0:r1], r0, model, SEED_5)
        cb5, bb5 = sharding.gather_layout(int(w5["n"].size), int(w5["coff"][-1]))      # 2 integers per rank: chunk and byte base of every shard
        g_coff = sharding.global_offsets(w5["coff"], bb5[rank])                         # this shard's offsets in the global compressed stream
        s_comp = torch.from_numpy(w5["comp"]).to(dev); s_coff = torch.from_numpy(w5["coff"]).to(dev)
        s_n = torch.from_numpy(w5["n"].astype(np.int32)).to(dev); s_off = torch.from_numpy(w5["off"].astype(np.int64)).to(dev)
        s_out = torch.empty(w5["total"], dtype=torch.int16, device=dev); s_st = torch.empty(w5["n"].size, dtype=torch.int32, device=dev)
        smax = int(w5["n"].max())
        barrier()
        ms5 = device_timed(torch, lambda: cd.dev_decode(codec, s_comp, s_coff, s_n, s_off, smax, s_out, s_st), max(3, args.steps), warm=3)
        ms5 = allreduce(ms5, "max")
        ok5 = bool(torch.equal(s_out, w5["d_sig"])) and not bool(s_st.any().item())
        ck_dec = table_checksum(torch, s_out, first_sample); ck_src = table_checksum(torch, w5["d_sig"], first_sample)
        if world > 1:
            dist.all_reduce(ck_dec); dist.all_reduce(ck_src)
        leg5 = splitter_leg(codec, w5["d_sig"], s_off, s_n, w5["n"], w5["total"])
        ms5e = allreduce(2e-6 * w5["total"] / leg5["value"], "max")
        all5 = allreduce(w5["total"]); enc5 = allreduce(w5["t_enc"], "max")
        extra["config5_strong_scaling"] = {
            "scaling": "strong", "table": f"{n5} reads, {int(all5)} samples ({2e-9*all5:.2f} GB of int16), seed {SEED_5}, {model.name}-shaped, codec {args.codec}",
            "sharding": "sharding.partition_reads: contiguous read ranges at the k/N quantiles of the sample prefix sum; sharding.gather_layout all-gathers "
                        "(chunks, bytes) per rank -- no collective on the data path",
            "ranks": world, "chunks_total": int(cb5[-1]), "compressed_bytes_total": int(bb5[-1]), "this_rank_global_byte_range": [int(g_coff[0]), int(g_coff[-1])],
            "decode": {"value": round(2e-6 * all5 / ms5, 2), "unit": UNIT, "ms_per_step": round(ms5, 3), "what": "device-resident whole-chunk decode of every shard, max over ranks"},
            "splitter_encode": {"value": round(2e-6 * all5 / ms5e, 2), "unit": UNIT, "what": "device-resident splitter stage of every shard, max over ranks"},
            "encode_host_zstd": {"value": round(2e-9 * all5 / enc5, 3), "unit": UNIT, "what": "whole-chunk encode through rnsc_compress_batch (host libzstd frame stage: all ranks share the host cores)"},
            "bit_exact": bool(allreduce(0.0 if ok5 else 1.0) == 0.0),
            "table_checksum_decoded": f"{int(ck_dec.item()) & ((1 << 64) - 1):016x}", "table_checksum_source": f"{int(ck_src.item()) & ((1 << 64) - 1):016x}",
            "cpu_baseline": {"decode": cpu, "encode": cpu_legs.get(f"{args.codec}_encode")}}
        log("config 5", extra["config5_strong_scaling"])
        del w5, s_comp, s_out
        torch.cuda.empty_cache()

    # ---- the per-chunk seam (signal_table_writer.cpp:98,107 / signal_table_reader.cpp:133,138, benchmark.cpp:134-190): one call
    #      per chunk through the C++ functions and the C API names, as a drop-in caller would make them ----
    seam_exe = os.path.join(ROOT, "integration", "_build", "seam_bench")
    if not args.no_extra and rank == 0 and os.path.exists(seam_exe):
        try:
            import tempfile
            ls = synth.read_lengths(args.seam_reads, SEED)
            ssig = synth.generate_reads(ls, SEED, model, device=dev).cpu().numpy()
            so, sn, _ = synth.chunk_table(ls)
            with tempfile.TemporaryDirectory() as td:
                pth = os.path.join(td, "chunks.bin")
                synth.write_chunk_file(pth, ssig, so, sn)
                r = subprocess.run([seam_exe, pth, "4"], capture_output=True, text=True, timeout=600)
            seam = json.loads(r.stdout) if r.returncode == 0 else {"error": r.stderr[-400:]}
            seam["what"] = (f"reference's per-chunk benchmark loop (one compress + one decompress call per chunk, host buffers) over the first {args.seam_reads} reads "
                            f"({int(sn.sum())} samples, {len(sn)} chunks): C++ seam (pod5::/pgnano::compress_signal, decompress_signal on arrow buffers) and C API names")
            seam["cpu_one_core_decode_GBps"] = cpu["one_core_value"] if cpu else None
            extra["per_chunk_seam"] = seam
            log("per-chunk seam", seam)
        except Exception as ex:  # noqa: BLE001
            extra["per_chunk_seam"] = {"error": str(ex)}

    if rank == 0:
        out = {"metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
               "ms_per_step": round(ms_step, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int16/u8",
               "data": "synthetic", "impl": "rnsc",
               "config": {"workload": f"synthetic {model.name}-shaped int16 signal (seed {SEED}), {args.reads} reads per GPU, POD5 chunks of 102400 samples, "
                                      f"pgnano {args.codec} whole-chunk batched decode (zstd frames -> streams -> int16)",
                          "codec": args.codec, "reads_per_gpu": args.reads, "chunks_per_gpu": int(nch), "samples_per_gpu": int(total),
                          "compressed_bytes_per_gpu": int(cbytes), "sharding": "reads, contiguous ranges, no data-path collective",
                          "l2": "inputs (compressed set) and outputs far exceed the 126 MB L2; no flush needed"},
               "bits_per_sample": round(8.0 * all_cbytes / all_samples, 4), "bit_exact": bit_exact,
               "gpu_launches": int(launches), "other_legs": extra, "chunk_routing": {"fast_4_per_sm": classes[0], "fast_2_per_sm": classes[1], "fast_1_per_sm": classes[2], "general": classes[3]},
               "clocks": clocks, "roofline": roof, "e2e": e2e, "cpu_baseline": cpu}
        emit(out)
    if world > 1:
        dist.destroy_process_group()
    cd.close()


def run_reference(args):
    """The reference's own CPU implementation of the path (oracle/_ref when built, else the C port) on all host cores."""
    from rawnanoporesignalcompression_b200 import synth
    from rawnanoporesignalcompression_b200.codec import codec_id
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    codec = codec_id(args.codec)
    model = synth.MODELS[args.model]
    cores = os.cpu_count() or 1
    kind = cpu_kind()
    smp = cpu_sample([codec], args.cpu_reads, SEED, model)
    per_step = max(2.0, min(args.cpu_seconds, 120.0 / max(1, args.steps + args.warmup)))
    for _ in range(args.warmup):
        cpu_rate(kind, codec, "decode", smp, cores, min(per_step, 2.0))
    rates, t0 = [], time.perf_counter()
    for _ in range(args.steps):
        r, done, wall = cpu_rate(kind, codec, "decode", smp, cores, per_step)
        rates.append(r)
    elapsed = time.perf_counter() - t0
    value = float(np.mean(rates))
    sample = (f"first {args.cpu_reads} reads of the workload ({smp['samples']} samples, {len(smp['n'])} chunks); each step = {cores} processes looping over "
              f"their static shard for ~{per_step:.1f} s")
    out = {"metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": round(1e3 * elapsed / max(1, args.steps), 2), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "int16/u8", "data": "synthetic", "impl": "reference",
           "config": {"workload": f"synthetic {model.name}-shaped int16 signal (seed {SEED}), pgnano {args.codec} whole-chunk decode on the host CPU "
                                  f"(bounded sample of the {args.reads}-read workload)", "codec": args.codec},
           "gpu_launches": 0,
           "cpu_baseline": {"value": round(value, 3), "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
           "e2e": {"value": round(value, 3), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(out)


_JSON_FD = None


def reserve_stdout():
    """Keeps the process's real stdout for the ONE JSON line: file descriptor 1 is pointed at stderr, so that banners
    that libraries write to stdout (NCCL prints its version there) cannot end up next to it."""
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)


def emit(out):
    line = (json.dumps(out) + "\n").encode()
    sys.stdout.flush()
    if _JSON_FD is None:
        os.write(1, line)
    else:
        os.write(_JSON_FD, line)


def main():
    reserve_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="rnsc", choices=["rnsc", "reference"])
    ap.add_argument("--reads", type=int, default=100_000, help="reads per GPU (BASELINE config: 100000)")
    ap.add_argument("--codec", default="C5")
    ap.add_argument("--model", default="R10.4.1", choices=["R10.4.1", "R9.4.1"])
    ap.add_argument("--cpu-reads", type=int, default=1000, help="reads in the bounded CPU sample (>= 1000 chunks: also the encode byte-parity sample)")
    ap.add_argument("--vbz-reads", type=int, default=0, help="reads per GPU of the R9.4.1 VBZ leg (0 = --reads)")
    ap.add_argument("--rc-reads", type=int, default=16000, help="reads per GPU of the range coder leg")
    ap.add_argument("--seam-reads", type=int, default=200, help="reads of the per-chunk seam leg")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--e2e-reads", type=int, default=0)
    ap.add_argument("--host-threads", type=int, default=0)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the other legs (configs 3, 4, 5, range coder)")
    ap.add_argument("--extra-reads", type=int, default=20000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--phases", action="store_true", help="print per-phase SM cycles of the fast decode kernels (tuning aid) to stderr")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_rnsc(args)


if __name__ == "__main__":
    main()
