"""Ad-hoc timing of the posterior stage (not the bench): python tests/quick_time.py C2"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np
from conftest import load_tables
from muscle_b200 import synth
from muscle_b200.engine import Engine
name = sys.argv[1] if len(sys.argv) > 1 else "C2"
seqs = synth.make_config(name)
e = Engine(0); e.set_hmm(load_tables()); e.set_seqs(seqs)
cells = synth.total_cells(seqs)
for it in range(3):
	t0 = time.time(); ea = e.posteriors_allpairs(); t1 = time.time()
	s = e.stats()
	print("%s iter %d: wall %.3f s, kernel %.1f ms, total %.1f ms -> %.2f Gcells/s (kernel), nnz %d, EA mean %.3f" %
	  (name, it, t1 - t0, s["last_kernel_ms"], s["last_total_ms"], cells/s["last_kernel_ms"]/1e6, e.store_nnz()[1], ea.mean()))
if len(sys.argv) > 2 and sys.argv[2] == "relax":
	for it in range(2):
		t0 = time.time(); e.consistency_iter(); t1 = time.time()
		s = e.stats()
		n = len(seqs)
		triples = n*(n - 1)//2*(n - 2)
		print("%s consistency iter %d: wall %.3f s, kernel %.1f ms -> %.3g (XY,Z) triples/s" % (name, it, t1 - t0, s["last_kernel_ms"], triples/(s["last_kernel_ms"]*1e-3)))
