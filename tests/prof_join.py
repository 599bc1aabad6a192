"""profiling helper: refinement-size joins on the device-resident MSA of one synthetic workload
    python tests/prof_join.py C3|S20 [n_refine]"""
import os
import random
import sys
import time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_tables          # noqa: E402
from muscle_b200 import synth             # noqa: E402
from muscle_b200.engine import Engine     # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "C3"
nref = int(sys.argv[2]) if len(sys.argv) > 2 else 3
seqs = synth.make_config("C5")[:int(cfg[1:])] if cfg.startswith("S") else synth.make_config(cfg)   # S20 = 20 proteins of ~300
n = len(seqs)
e = Engine(0)
e.set_hmm(load_tables())
e.set_seqs(seqs)
e.posteriors_allpairs(want_ea=False)
e.msa_reset()
t0 = time.time()
order = [0]
for k in range(1, n):                      # a chain of joins (cheap: one new sequence each)
	cols, _, _ = e.msa_join(order, [k])
	order.append(k)
print(cfg, "chain of %d joins: %.2f s, %d columns" % (n - 1, time.time() - t0, cols))
rng = random.Random(1)
for it in range(nref):
	pick = [rng.random() < 0.5 for _ in order]
	a = [s for s, f in zip(order, pick) if f]
	b = [s for s, f in zip(order, pick) if not f]
	t0 = time.time()
	cols, score, _ = e.msa_join(a, b)
	print(cfg, "refinement join %d: %d x %d sequences -> %d columns, %.1f ms" % (it, len(a), len(b), cols, 1e3*(time.time() - t0)))
	order = a + b
e.close()
