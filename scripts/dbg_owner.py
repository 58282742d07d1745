import os, sys, torch, numpy as np
sys.path.insert(0, '/root/repo')
from soapdenovo2_b200 import api, synth, dist as pdist
M = (1 << 64) - 1
def mix64(x):
    x ^= x >> 32; x = (x * 0xD6E8FEB86659FD93) & M; x ^= x >> 32; x = (x * 0xD6E8FEB86659FD93) & M; x ^= x >> 32; return x
def table_hash(ws):
    h = 0
    for w in ws: h = (((h ^ w) * 0x9E3779B97F4A7C15) + (h >> 29)) & M
    return mix64(h)
d = "/tmp/dbgo"; os.makedirs(d, exist_ok=True); synth.scenario_pe_fastq(d)
eng = api.PregraphEngine(K=63, P=8, initG=1, max_rd_len=150, world=2, rank=0)
eng.feed_text(open(d + "/pe_1.fq", "rb").read(), fastq=True, ord_base=0, ord_stride=2)
ptr, ranges, tb = eng.exchange_buffer()
t = torch.as_tensor(pdist.DeviceMemory(ptr, ranges[-1] * tb), device="cuda").cpu().numpy().view(np.uint64).reshape(-1, tb // 8)
print("ranges", ranges, "tuples", t.shape)
bad = 0
for o in range(2):
    seg = t[ranges[o]:ranges[o + 1]]
    owners = np.array([(table_hash([int(r[0]), int(r[1])]) >> 40) % 2 for r in seg[:: max(1, len(seg) // 2000)]])
    print("owner", o, "sampled", len(owners), "wrong", int((owners != o).sum()))
