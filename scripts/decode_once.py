"""Write a slice of config 2 as a BAM and decode it on the GPU a few times (the command ncu captures bgzf.cu's kernels from).
    python scripts/decode_once.py [n_reads] [repeats]"""
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from trueconsense_b200 import bamio, gpu, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
rep = int(sys.argv[2]) if len(sys.argv) > 2 else 3
w = synth.config(1, scale=n / 2_000_000)
batch = synth.generate_reads(w.params, w.ref)
ctx = gpu.Context(0)
with tempfile.TemporaryDirectory() as tmp:
    path = os.path.join(tmp, "x.bam")
    bamio.write_bam(path, batch, "ref", len(w.ref), level=1)
    for i in range(rep):
        t0 = time.perf_counter()
        dev = ctx.bam_file_to_device(path)
        print(i, dev.n_reads, f"{(time.perf_counter() - t0) * 1e3:.2f} ms", {k: round(v * 1e3, 2) for k, v in dev.info.items() if k.startswith("t_")})
