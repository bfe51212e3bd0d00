"""Streaming repack at scale: writes a synthetic VBZ POD5 of READS reads (>= 4 GB at the default),
converts it to pgnano C5 with the streaming CLI in a FRESH process (peak RSS of that process = the number to look at), with
the host-libzstd and with the GPU entropy stage, and checks that input and outputs hold the same signal (crc32 of every
decoded sample, decoded group by group). Prints one JSON object."""
import json, os, resource, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
reads = int(sys.argv[1]) if len(sys.argv) > 1 else 105000
work = sys.argv[2] if len(sys.argv) > 2 else "/tmp/rnsc_stream"
os.makedirs(work, exist_ok=True)
import torch
from rawnanoporesignalcompression_b200 import synth, stream_convert as sc
from rawnanoporesignalcompression_b200.codec import Codec
cd = Codec(0)
src = os.path.join(work, "in_vbz.pod5")
t0 = time.perf_counter()
size, total = sc.write_synthetic_pod5(src, cd, reads, 20261018, synth.R10_4_1, "VBZ", template=os.path.join(ROOT, "tests", "golden", "multi_fast5_zip_v3.pod5"),
                                      reads_per_group=4096, device=torch.device("cuda:0"))
t_write = time.perf_counter() - t0
res = {"input": {"reads": reads, "samples": total, "file_bytes": size, "int16_GB": round(2e-9 * total, 3), "write_s": round(t_write, 1)}}
crc_in = sc.signal_checksum(src, "C5", cd)
cd.close(); del cd
for mode in ("gpu", "libzstd"):
    dst = os.path.join(work, f"out_{mode}.pod5")
    prog = ("import resource, sys; from rawnanoporesignalcompression_b200 import cli; rc = cli.main(sys.argv[1:]); "
            "sys.stderr.write('Maximum resident set size (kbytes): %d\\n' % resource.getrusage(resource.RUSAGE_SELF).ru_maxrss); sys.exit(rc)")
    cmd = [sys.executable, "-c", prog, src, dst, "--pgnano", "--entropy", mode]
    t0 = time.perf_counter()
    r = subprocess.run(cmd, capture_output=True, text=True, cwd=ROOT, env=dict(os.environ, RNSC_STREAM_STATS="1"))
    dt = time.perf_counter() - t0
    rss = [ln for ln in r.stderr.splitlines() if "Maximum resident set size" in ln]
    rss_kb = int(rss[0].split(":")[1]) if rss else None
    stats = [ln for ln in r.stderr.splitlines() if ln.startswith("[stream]")]
    res[mode] = {"rc": r.returncode, "wall_s": round(dt, 2), "stdout": r.stdout.strip(), "max_rss_GB": round(rss_kb / 1e6, 3) if rss_kb else None,
                 "int16_GBps_wall_incl_process_start": round(2e-9 * total / dt, 2), "out_bytes": os.path.getsize(dst) if os.path.exists(dst) else None,
                 "stages": stats[0] if stats else None}
    if r.returncode:
        res[mode]["stderr"] = r.stderr[-600:]
cd = Codec(0)
for mode in ("gpu", "libzstd"):
    dst = os.path.join(work, f"out_{mode}.pod5")
    if os.path.exists(dst):
        res[mode]["same_signal_as_input"] = sc.signal_checksum(dst, "C5", cd) == crc_in
res["input"]["crc32_samples_rows"] = list(crc_in)
print(json.dumps(res))
