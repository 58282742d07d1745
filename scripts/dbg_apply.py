import os, sys, torch, numpy as np
sys.path.insert(0, '/root/repo')
from soapdenovo2_b200 import api, synth, dist as pdist
from tests import util
util.build_oracle()
d = "/tmp/dbga"; os.makedirs(d, exist_ok=True); cfg = synth.scenario_pe_fastq(d)
util.run_model(util.MODEL63, cfg, d + "/mod", 63, 8, ("-1", "-a", "1"))
eng = api.PregraphEngine(K=63, P=8, initG=1, max_rd_len=150, world=2, rank=0)
for fn, mate in ((d + "/pe_1.fq", 0), (d + "/pe_2.fq", 1)):
    eng.feed_text(open(fn, "rb").read(), fastq=True, ord_base=mate, ord_stride=2)
    ptr, ranges, tb = eng.exchange_buffer()
    t = torch.as_tensor(pdist.DeviceMemory(ptr, ranges[-1] * tb), device="cuda").clone()   # like a received buffer
    eng.apply_tuples(t.data_ptr(), ranges[-1])
    eng.exchange_clear()
st = eng.finish_pass1(); hist, _, _ = eng.sweeps()
print("distinct", st.distinct, "match", api.kmerfreq_text(hist) == open(d + "/mod.kmerFreq", "rb").read(), hist[1:4], open(d + "/mod.kmerFreq").read().split()[:3])
