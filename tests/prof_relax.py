"""profiling helper: posteriors + consistency iterations on one synthetic workload
    python tests/prof_relax.py C2 [iters]"""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_tables          # noqa: E402
from muscle_b200 import synth             # noqa: E402
from muscle_b200.engine import Engine     # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "C2"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 2
seqs = synth.make_config(cfg)
e = Engine(0)
e.set_hmm(load_tables())
e.set_seqs(seqs)
e.posteriors_allpairs(want_ea=False)
print(cfg, "posterior kernel ms", e.stats()["last_kernel_ms"])
for it in range(iters):
	e.consistency_iter()
	print(cfg, "relax iter", it, "kernel ms", e.stats()["last_kernel_ms"], "total ms", e.stats()["last_total_ms"])
e.close()
