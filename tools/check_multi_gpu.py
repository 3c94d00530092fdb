"""In-library gene sharding over all visible GPUs must give records bit-identical to one GPU."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sclane_b200 as sb
from tools import synth
n = sb.device_count()
counts, pt, _ = synth.make_counts(301, 20000, seed=5)      # D = 10,000 abscissae: the X-compressed route
a = sb.testDynamic(counts, pt, n_gpus=1, preds="none")
b = sb.testDynamic(counts, pt, n_gpus=n, preds="none")
off = sb._abi.SclaneRecord.gene_time.offset


def masked(r):
    v = bytearray(bytes(r))
    v[off:off + 8] = bytes(8)          # gene_time is a wall-clock surrogate
    c = sb._abi.SclaneRecord.stage_cycles.offset
    v[c:c + 24] = bytes(24)            # clock64 profile
    return bytes(v)


bad = sum(masked(a.records[i]) != masked(b.records[i]) for i in range(301))
print(f"devices {n}: {bad} records differ between 1 and {n} GPUs; profile {b.profile['ms']}")
assert bad == 0
