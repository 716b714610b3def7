import sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np
from collections import Counter
os.chdir("/root/repo")
from oracle import pileup, call
from trueconsense_b200 import gpu
from trueconsense_b200.reads import ReadBatch
import importlib.util
spec = importlib.util.spec_from_file_location("tg", "/root/repo/tests/test_gpu_parity.py")
import pytest
tg = importlib.util.module_from_spec(spec); spec.loader.exec_module(tg)
pileup.build()
ctx = gpu.Context(0)
L, col = 400, 200
def check(recs, positions, mode=0):
    b = ReadBatch.from_records(recs)
    p = gpu.extractinserts_params(); p.reserved = mode << 8
    got = ctx.extract_inserts(b, L, positions, p)
    bad = []
    for g in got:
        want = tg._expected_call(pileup, call, b, g["pos"], reserved=mode << 8)
        if (g["string"], g["n_entries"], g["mode_count"]) != want:
            bad.append((g["pos"], (g["string"], g["n_entries"], g["mode_count"]), want))
    return bad
found = 0
for seed in range(400):
    rng = np.random.default_rng(5000 + seed)
    recs = tg._paired_fuzz_records(rng, 8, L, col)
    positions = [col - 3, col, col + 1, col + 2, col + 9, col + 17]
    bad = check(recs, positions)
    if not bad: continue
    # minimise: drop names one at a time
    names = sorted(set(r["qname"] for r in recs))
    cur = recs
    for nm in names:
        trial = [r for r in cur if r["qname"] != nm]
        if trial and check(trial, positions): cur = trial
    bad = check(cur, positions)
    print("seed", seed, "bad", bad)
    for r in cur: print({k: (v if k != "qual" else v) for k, v in r.items()})
    b = ReadBatch.from_records(cur)
    for pos, _, _ in bad:
        print("oracle strings", pos, pileup.pileup_columns(b, region=(pos-1,pos), **pileup.EXTRACTINSERTS))
        print("oracle off    ", pos, pileup.pileup_columns(b, region=(pos-1,pos), reserved=1<<8, **pileup.EXTRACTINSERTS))
    found += 1
    if found >= 3: break
print("done", found)
