import sys, os, collections
sys.path.insert(0, '/root/repo')
from soapdenovo2_b200 import api, synth
from tests import util
util.build_oracle()
d = '/tmp/dbg127'; os.makedirs(d, exist_ok=True)
cfg = synth.scenario_pe_fastq(d)
K = int(sys.argv[1]) if len(sys.argv) > 1 else 127
util.run_model(util.MODEL127, cfg, d + '/mod', K, 3, ("-1", "-T", d + '/mod.table', "-a", "1"))
eng = api.PregraphEngine(K=K, P=3, initG=1, flavour127=1, max_rd_len=150)
eng.feed_text(open(d + '/pe_1.fq', 'rb').read(), fastq=True, ord_base=0, ord_stride=2)
eng.feed_text(open(d + '/pe_2.fq', 'rb').read(), fastq=True, ord_base=1, ord_stride=2)
st = eng.finish_pass1(); eng.sweeps(); eng.build_layout()
got = eng.dump_nodes(); want = open(d + '/mod.table', 'rb').read()
R = 42
G = collections.Counter(got[i:i+32] for i in range(0, len(got), R))
W = {want[i:i+32]: want[i+32:i+R] for i in range(0, len(want), R)}
Gd = {got[i:i+32]: got[i+32:i+R] for i in range(0, len(got), R)}
print('distinct gpu', st.distinct, 'records gpu', len(got)//R, 'model', len(want)//R, 'unique gpu keys', len(G))
dups = [k for k, c in G.items() if c > 1]
print('dup keys in gpu:', len(dups))
missing = [k for k in W if k not in Gd]; extra = [k for k in Gd if k not in W]
print('missing', len(missing), 'extra', len(extra))
for k in extra[:5]:
    print('extra', k.hex(), Gd[k].hex())
for k in missing[:5]:
    print('missing', k.hex(), W[k].hex())
diffv = [k for k in W if k in Gd and W[k] != Gd[k]]
print('value diffs', len(diffv))
for k in diffv[:5]:
    print(k.hex(), 'gpu', Gd[k].hex(), 'model', W[k].hex())
