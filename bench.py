#!/usr/bin/env python
"""bench.py -- MPCFlat all-pairs posterior stage, DP cells/s (BASELINE.json metric).

One "step" = one pass of the hot path (Forward + Backward + posterior + sparsify + EA for every
sequence pair) over one synthetic protein family.  Default workload is BASELINE.json configs[2]
(C3: 1000 proteins, mean length 350), the configuration the north_star's >=50x target is quoted
on; it fits one B200.  With --gpus N the N(N-1)/2 pairs are sharded over the ranks in contiguous,
cell-balanced ranges (strong scaling, no data-path collective in the timed region).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload C3]

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for the definition of every field.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

METRIC = "pair-HMM DP cells/sec (all-pairs Fwd+Bwd+posterior)"
UNIT = "cells/s"
ALGO_BYTES_PER_CELL = 12.0      # 4 B write + 4 B read of Forward-M, 4 B posterior
ISSUE_FLOP_PER_CELL = 243.0


def load_peaks():
	p = os.path.join(ROOT, "MEASURED_PEAKS.json")
	if os.path.exists(p):
		with open(p) as f:
			d = json.load(f)
		return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
	return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(threading.Thread):
	"""nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

	Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
	  "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
	  "clocks_event_reasons.sw_power_cap")

	def __init__(self, index):
		super().__init__(daemon=True)
		self.index = index
		self.rows = []
		self.stop_flag = False
		self.proc = None

	def run(self):
		try:
			self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
			  "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
			for line in self.proc.stdout:
				self.rows.append(line.strip())
				if self.stop_flag:
					break
		except Exception:
			pass

	def finish(self):
		self.stop_flag = True
		if self.proc is not None:
			try:
				self.proc.terminate()
			except Exception:
				pass
		sm, mx, reasons = [], [], set()
		names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
		for r in self.rows:
			f = [x.strip() for x in r.split(",")]
			if len(f) < 7:
				continue
			try:
				sm.append(float(f[0]))
				mx.append(float(f[1]))
			except ValueError:
				continue
			for k, nm in enumerate(names):
				if f[3 + k].lower().startswith("active"):
					reasons.add(nm)
		if not sm:
			return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
		# median over the samples taken under load (upper half of the observed clocks)
		return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
		  "samples": len(sm)}


def shard_ranges(seqs, world):
	"""contiguous ranges of the row-major pair list with ~equal DP cells per rank"""
	L = np.array([len(s) for s in seqs], np.float64)
	n = len(seqs)
	iu, ju = np.triu_indices(n, 1)
	cost = np.cumsum(L[iu]*L[ju])
	total = cost[-1]
	cuts = [0]
	for r in range(1, world):
		cuts.append(int(np.searchsorted(cost, total*r/world)))
	cuts.append(len(iu))
	cells = [float((cost[cuts[r + 1] - 1] - (cost[cuts[r] - 1] if cuts[r] > 0 else 0.0))) for r in range(world)]
	return [(cuts[r], cuts[r + 1]) for r in range(world)], cells, float(total)


def host_threads():
	"""threads this process may run on (cgroup/affinity aware)"""
	try:
		return max(1, len(os.sched_getaffinity(0)))
	except Exception:
		return max(1, os.cpu_count() or 1)


_REF = None


def load_reference():
	"""The compiled reference for the CPU legs: the timing build (oracle/_ref/libmuscle_ref_fast.so,
	-O3 with AVX2+FMA, what upstream ships) when the host CPU supports it, else the strict-IEEE parity
	build.  torchrun exports OMP_NUM_THREADS=1 to its workers: the OpenMP runtime reads it when the
	library is loaded, so it is overridden HERE with the number of threads the process may use."""
	global _REF
	if _REF is not None:
		return _REF
	os.environ["OMP_NUM_THREADS"] = str(host_threads())
	os.environ.pop("OMP_THREAD_LIMIT", None)
	from oracle import pyoracle
	fast = os.path.join(os.path.dirname(pyoracle.REF_SO), "libmuscle_ref_fast.so")
	build = "parity (-O3 -ffp-contract=off)"
	try:
		flags = open("/proc/cpuinfo").read()
		has_v3 = all((" " + f + " ") in flags or (" " + f + "\n") in flags for f in ("avx2", "fma", "bmi2"))
	except Exception:
		has_v3 = False
	if os.path.exists(fast) and has_v3 and not os.environ.get("MB200_REF_PARITY_BUILD"):
		pyoracle.REF_SO = fast
		build = "timing (-O3 -mavx2 -mfma, FMA contraction on)"
	if not os.path.exists(pyoracle.REF_SO):
		_REF = (None, "port")
		return _REF
	R = pyoracle.Ref(threads=host_threads())
	_REF = (R, build)
	return _REF


def cpu_reference_run(seqs, target_cells, threads=0, seconds=None):
	"""Time the compiled reference (oracle/_ref) -- or the C port when it is absent -- on a bounded
	sample (a prefix of the row-major pair list holding ~target_cells DP cells).  With `seconds`
	the sample is sized from a short calibration run so that it takes about that long."""
	if seconds is not None:
		probe = cpu_reference_run(seqs, 2.0e8, threads)
		target_cells = max(2.0e8, probe["value"]*seconds)
	from conftest import load_tables
	from oracle import pyoracle
	L = np.array([len(s) for s in seqs], np.float64)
	n = len(seqs)
	iu, ju = np.triu_indices(n, 1)
	cost = np.cumsum(L[iu]*L[ju])
	hi = int(min(len(iu), max(1, np.searchsorted(cost, target_cells) + 1)))
	cells = float(cost[hi - 1])
	R, build = load_reference()
	if R is not None:
		M = R.mpc(seqs)
		nthr = int(R.threads) if threads <= 0 else int(threads)
		secs = M.posteriors_range(0, hi, nthr)
		M.close()
		kind, cores = "reference", nthr
	else:
		O = pyoracle.Oracle(load_tables())
		nthr = host_threads() if threads <= 0 else int(threads)
		t0 = time.time()
		O.all_pairs(seqs, 0, hi, threads=nthr, want_sparse=False)
		secs = time.time() - t0
		kind, cores = "port", nthr
	return {"value": cells/secs, "unit": UNIT, "cores": int(cores), "kind": kind, "build": build, "seconds": secs,
	  "sample": "first %d of %d pairs (row-major), %.3g cells, OpenMP dynamic schedule, %d threads" %
	  (hi, len(iu), cells, cores)}


def cpu_relax_run(seqs, eng_export, seconds=10.0):
	"""(XY,Z) triples/s of the reference's ConsPair on a bounded sample: the first pairs of a GPU-made
	store imported into the reference MPCFlat.  eng_export(p) -> (off, ent)."""
	R, build = load_reference()
	if R is None:
		return None
	n = len(seqs)
	M = R.mpc(seqs)
	npairs = n*(n - 1)//2
	for p in range(npairs):
		off, ent = eng_export(p)
		M.import_(p, off, ent)
	probe = M.conspairs_range(0, 4, int(R.threads))
	cnt = int(max(8, min(npairs, 4*seconds/max(probe, 1e-6))))
	secs = M.conspairs_range(0, cnt, int(R.threads))
	M.close()
	return {"value": cnt*(n - 2)/secs, "unit": "(XY,Z) triples/s", "cores": int(R.threads), "kind": "reference", "build": build,
	  "sample": "ConsPair on the first %d of %d pairs, %d sequences" % (cnt, npairs, n), "seconds": secs}


def main():
	ap = argparse.ArgumentParser()
	ap.add_argument("--gpus", type=int, default=1)
	ap.add_argument("--steps", type=int, default=3)
	ap.add_argument("--warmup", type=int, default=3)
	ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
	ap.add_argument("--workload", default="C3")
	ap.add_argument("--cpu-seconds", type=float, default=15.0, help="target CPU time of the baseline sample")
	ap.add_argument("--no-cpu-baseline", action="store_true")
	ap.add_argument("--no-pipeline", action="store_true", help="skip the exchange/relax stages after the timed region")
	ap.add_argument("--no-extras", action="store_true", help="skip the secondary records (C4, C2 relax, CLI wall time)")
	args = ap.parse_args()

	rank = int(os.environ.get("RANK", "0"))
	world = int(os.environ.get("WORLD_SIZE", "1"))
	local_rank = int(os.environ.get("LOCAL_RANK", "0"))

	from muscle_b200 import synth
	seqs = synth.make_config(args.workload)
	n = len(seqs)
	lens = [len(s) for s in seqs]
	config = {"workload": "%s: %d synthetic proteins, mean length %.0f (muscle_b200.synth seed %d), all %d pairs" %
	  (args.workload, n, float(np.mean(lens)), synth.CONFIGS[args.workload][3], n*(n - 1)//2),
	  "stage": "MPCFlat::CalcPosteriors (Fwd+Bwd+posterior+sparsify+EA)",
	  "sharding": "contiguous cell-balanced pair ranges, %d rank(s)" % world,
	  "l2": "per-step working set (Forward-M spill + sparse output, >10 GB) exceeds the 126 MB L2"}

	# ------------------------------------------------------------------ reference arm
	if args.impl == "reference":
		if rank != 0:
			return
		probe = cpu_reference_run(seqs, 2.0e8)
		target = max(2.0e8, probe["value"]*args.cpu_seconds)     # each step ~cpu_seconds of host time
		vals = []
		last = None
		for _ in range(args.warmup):
			cpu_reference_run(seqs, target/8)
		t0 = time.time()
		for _ in range(args.steps):
			last = cpu_reference_run(seqs, target)
			vals.append(last["value"])
		wall = time.time() - t0
		v = float(np.mean(vals))
		out = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
		  "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3*wall/max(1, args.steps),
		  "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
		  "config": config,
		  "cpu_baseline": {"value": v, "unit": UNIT, "cores": last["cores"], "kind": last["kind"], "build": last["build"],
		    "sample": last["sample"]},
		  "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
		print(json.dumps(out))
		return

	# ------------------------------------------------------------------ our arm
	# stdout carries exactly ONE line (the JSON record): everything libraries print meanwhile (e.g. NCCL's
	# version banner) is sent to stderr
	sys.stdout.flush()
	saved_stdout = os.dup(1)
	os.dup2(2, 1)
	import torch
	import torch.distributed as dist
	if not torch.cuda.is_available():
		raise SystemExit("bench.py: no CUDA device; libmuscle_b200 has no CPU path")
	torch.cuda.set_device(local_rank)
	if world > 1:
		dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

	from conftest import load_tables
	from muscle_b200.engine import Engine
	from muscle_b200 import dist as mdist
	tables = load_tables()
	eng = Engine(local_rank)
	eng.set_hmm(tables)
	eng.set_seqs(seqs)
	ranges, cells_per_rank, total_cells = shard_ranges(seqs, world)
	p_lo, p_hi = ranges[rank]
	my_cells = cells_per_rank[rank]

	def barrier():
		if world > 1:
			dist.barrier()
		torch.cuda.synchronize()

	def step_resident():
		# inputs (sequences, tables, pair list) already in HBM; EA stays on the device
		eng.posteriors_allpairs(p_lo, p_hi, want_ea=False)
		s = eng.stats()
		return s["last_total_ms"], s["last_kernel_ms"]

	def step_e2e():
		# public API with host buffers: sequences up, EA matrix slice down, every step
		eng.set_seqs(seqs)
		ea = eng.posteriors_allpairs(p_lo, p_hi, want_ea=True)
		s = eng.stats()
		return float(ea[0]), s["h2d_bytes"], s["d2h_bytes"]

	for _ in range(max(3, args.warmup)):
		step_resident()
	launches0 = eng.stats()["kernel_launches"]
	sampler = ClockSampler(local_rank)
	if rank == 0:
		sampler.start()
	barrier()
	t0 = time.time()
	dev_ms = 0.0
	kern_ms = 0.0
	for _ in range(args.steps):
		tot, k = step_resident()
		dev_ms += tot
		kern_ms += k
	barrier()
	wall = time.time() - t0
	clocks = sampler.finish() if rank == 0 else None
	launches = eng.stats()["kernel_launches"] - launches0

	# end to end through the host-facing call
	step_e2e()
	barrier()
	t1 = time.time()
	h2d = d2h = 0
	for _ in range(args.steps):
		_, h2d, d2h = step_e2e()
	barrier()
	wall_e2e = time.time() - t1

	# ---- the rest of MPCFlat's GPU pipeline on the same workload (untimed for `value`; reported in
	# `pipeline`): store all-gather-v (N>1), two sharded consistency iterations, entry exchange (N>1)
	pipe = {"x1_ms": 0.0, "x1_bytes": 0.0, "relax_ms": [0.0, 0.0], "relax_kernel_ms": [0.0, 0.0], "x2_ms": [0.0, 0.0],
	  "x2_bytes": 0.0, "prep_ms": 0.0}
	if not args.no_pipeline and n >= 3:
		barrier()
		if world > 1:
			nb, secs = mdist.gather_store(eng, None, have_store=p_hi > p_lo)
			pipe["x1_ms"], pipe["x1_bytes"] = secs*1e3, float(nb)
		eranges = None
		for it in range(2):
			barrier()
			eng.consistency_iter(p_lo, p_hi)
			st = eng.stats()
			pipe["relax_ms"][it] = st["last_total_ms"]
			pipe["relax_kernel_ms"][it] = st["last_kernel_ms"]
			if it == 0:
				pipe["prep_ms"] = st["last_total_ms"] - st["last_kernel_ms"]
			if world > 1:
				if eranges is None:
					nnz, _ = eng.store_nnz()
					base = np.concatenate([[0], np.cumsum(nnz.astype(np.int64))])
					eranges = [(int(base[a]), int(base[b])) for (a, b) in ranges]
				nb, secs = mdist.gather_values(eng, eranges, rank, None)
				pipe["x2_ms"][it], pipe["x2_bytes"] = secs*1e3, float(nb)
		barrier()

	# max over ranks (device time of the timed region, wall of both loops, pipeline stages)
	vec = [dev_ms, kern_ms, wall, wall_e2e, pipe["x1_ms"], pipe["relax_ms"][0], pipe["relax_ms"][1],
	  pipe["relax_kernel_ms"][0], pipe["relax_kernel_ms"][1], pipe["x2_ms"][0], pipe["x2_ms"][1], pipe["prep_ms"]]
	tv = torch.tensor(vec, dtype=torch.float64, device="cuda")
	if world > 1:
		dist.all_reduce(tv, op=dist.ReduceOp.MAX)
	(dev_ms_max, kern_ms_max, wall_max, wall_e2e_max, x1_ms, r_ms0, r_ms1, rk_ms0, rk_ms1, x2_ms0, x2_ms1, prep_ms) = \
	  [float(x) for x in tv.tolist()]

	out = None
	if rank == 0:
		steps = args.steps
		# value: whole-job cells over the max-over-ranks DEVICE time of the K timed steps (CUDA events on
		# the library's launching stream, first launch to last completion of every step); the
		# barrier-to-barrier wall clock of the same region is reported beside it
		value = total_cells*steps/(dev_ms_max*1e-3)
		e2e_value = total_cells*steps/wall_e2e_max
		peak, peak_src = load_peaks()
		# dominant kernel k_posterior_sm (one launch per shared-memory size class): live CUDA-event time of
		# the launches of this rank, algorithmic bytes = 12 B/cell x cells of this rank
		achieved = ALGO_BYTES_PER_CELL*my_cells*steps/(kern_ms*1e-3)/1e9
		# DRAM traffic and instruction count per cell come from the committed ncu capture of the SHIPPED
		# kernel (profiles/traffic_bytes_per_cell.json names the capture); per step like `achieved`
		traffic = None
		winst_per_cell = None
		traffic_src = None
		tp = os.path.join(ROOT, "profiles", "traffic_bytes_per_cell.json")
		if os.path.exists(tp):
			try:
				with open(tp) as f:
					prof = json.load(f)
				traffic = prof.get("dram_bytes_per_cell", None)
				winst_per_cell = prof.get("warp_inst_per_cell", None)
				traffic_src = prof.get("source", None)
				if traffic is not None:
					traffic = traffic*my_cells
			except Exception:
				traffic = None
		npairs_all = n*(n - 1)//2
		triples = float(npairs_all)*(n - 2)
		pipeline = None
		if not args.no_pipeline and n >= 3:
			pipeline = {
			  "what": "rest of the MPCFlat GPU pipeline on the same workload, per-stage device/host time, max over ranks",
			  "exchange1_store_allgather": None if world == 1 else {"bytes_received_per_gpu": x1_bytes_f(pipe), "ms": x1_ms,
			    "gbps_per_gpu": x1_bytes_f(pipe)/max(x1_ms, 1e-9)/1e6, "how": "one ncclBroadcast per source rank straight into the final store buffers"},
			  "relax": {"triples": triples, "ms_per_iter": [r_ms0, r_ms1], "kernel_ms_per_iter": [rk_ms0, rk_ms1],
			    "triples_per_s": triples/max(rk_ms1*1e-3, 1e-12), "one_time_prep_ms": prep_ms,
			    "note": "iteration 1 total includes building the transposed store and the column masks (one_time_prep_ms)"},
			  "exchange2_entries_allgather": None if world == 1 else {"bytes_received_per_gpu": pipe["x2_bytes"], "ms_per_iter": [x2_ms0, x2_ms1],
			    "gbps_per_gpu": pipe["x2_bytes"]/max(x2_ms1, 1e-9)/1e6}}
		out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": max(3, args.warmup),
		  "ms_per_step": dev_ms_max/steps, "wall_ms_per_step": 1e3*wall_max/steps, "higher_is_better": True,
		  "scaling": "strong", "vs_baseline": None,
		  "dtype": "f32", "data": "synthetic", "config": config,
		  "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
		    "what": "mb200_set_seqs(host bytes) + mb200_posteriors_allpairs -> EA on host; sparse store stays in HBM for the next stage"},
		  "gpu_launches": int(launches),
		  "clocks": clocks,
		  "roofline": {"bound": "hbm", "kernel": "k_posterior_sm", "achieved": achieved, "peak": peak, "unit": "GB/s",
		    "frac": achieved/peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
		    "algorithmic_bytes_per_cell": ALGO_BYTES_PER_CELL,
		    "kernel_ms_per_step": kern_ms/steps,
		    "note": "kernel is fp32-issue bound, not HBM bound (SURVEY.md 8d): see issue_gflops",
		    "issue_gflops": ISSUE_FLOP_PER_CELL*my_cells*steps/(kern_ms*1e-3)/1e9,
		    "issue": None if not (winst_per_cell and clocks and clocks.get("sm_mhz")) else {
		      "what": "warp-instructions issued per second vs 148 SMs x 4 schedulers x SM clock (1 issue/clk each)",
		      "warp_inst_per_cell": winst_per_cell,
		      "achieved_ginst": winst_per_cell*my_cells*steps/(kern_ms*1e-3)/1e9,
		      "peak_ginst": 148*4*clocks["sm_mhz"]*1e6/1e9,
		      "frac": winst_per_cell*my_cells*steps/(kern_ms*1e-3)/(148*4*clocks["sm_mhz"]*1e6)}},
		  "pipeline": pipeline,
		  "device_ms_per_step": dev_ms_max/steps}
	if world > 1:
		dist.barrier()
		dist.destroy_process_group()
	eng.close()
	if rank == 0:
		if not args.no_cpu_baseline and world == 1:
			out["cpu_baseline"] = cpu_reference_run(seqs, 0, seconds=args.cpu_seconds)
		if world == 1 and not args.no_extras:
			try:
				out["extra"] = extras(tables, args)
			except Exception as e:           # the headline line must not die on a secondary measurement
				out["extra"] = {"error": repr(e)}
		sys.stdout.flush()
		os.dup2(saved_stdout, 1)
		print(json.dumps(out))
		sys.stdout.flush()


def x1_bytes_f(pipe):
	return float(pipe["x1_bytes"])


def extras(tables, args):
	"""Secondary records of SURVEY.md section 8d, 1 GPU: C4 (long sequences) cells/s, relax triples/s
	on C2 with the reference's ConsPair timed beside it, and `muscle_b200 -align` wall time."""
	from muscle_b200 import synth
	from muscle_b200.engine import Engine
	ex = {}
	# ---- C4: 128 proteins of 1500-3000 residues (multi-strip path)
	seqs4 = synth.make_config("C4")
	e = Engine(0)
	e.set_hmm(tables)
	e.set_seqs(seqs4)
	for _ in range(2):
		e.posteriors_allpairs(want_ea=False)
	ms = 0.0
	for _ in range(2):
		e.posteriors_allpairs(want_ea=False)
		ms += e.stats()["last_total_ms"]
	ex["C4_posterior"] = {"value": synth.total_cells(seqs4)*2/(ms*1e-3), "unit": UNIT, "ms_per_step": ms/2,
	  "workload": "C4: %d proteins, lengths 1500-3000, all pairs" % len(seqs4)}
	e.close()
	# ---- relax on C2, GPU and the reference's ConsPair
	seqs2 = synth.make_config("C2")
	n2 = len(seqs2)
	e = Engine(0)
	e.set_hmm(tables)
	e.set_seqs(seqs2)
	e.posteriors_allpairs(want_ea=False)
	offs, ents = e.export_all()
	kms = []
	for _ in range(2):
		e.consistency_iter()
		kms.append(e.stats()["last_kernel_ms"])
	tri2 = float(n2*(n2 - 1)//2)*(n2 - 2)
	ex["C2_relax"] = {"value": tri2/(kms[1]*1e-3), "unit": "(XY,Z) triples/s", "kernel_ms_per_iter": kms,
	  "workload": "C2: %d proteins, %d pairs" % (n2, n2*(n2 - 1)//2)}
	e.close()
	if not args.no_cpu_baseline:
		cb = cpu_relax_run(seqs2, lambda p: (offs[p], ents[p]), seconds=8.0)
		if cb is not None:
			ex["C2_relax"]["cpu_baseline"] = cb
	# ---- the drop-in binary end to end (reference CLI + GPU engine), if it was built
	cli = os.path.join(ROOT, "integration", "_build", "muscle_b200")
	refcli = os.path.join(ROOT, "oracle", "_ref", "muscle")
	if os.path.exists(cli):
		import tempfile
		rec = {}
		with tempfile.TemporaryDirectory() as td:
			for name in ("C1", "C2"):
				sq = synth.make_config(name)
				fa = os.path.join(td, name + ".fa")
				with open(fa, "w") as f:
					for i, q in enumerate(sq):
						f.write(">s%d\n%s\n" % (i, q))
				t0 = time.time()
				r = subprocess.run([cli, "-align", fa, "-output", os.path.join(td, name + ".gpu.afa")], capture_output=True, text=True)
				rec[name + "_gpu_wall_s"] = time.time() - t0
				rec[name + "_gpu_ok"] = r.returncode == 0
				if name == "C1" and os.path.exists(refcli) and not args.no_cpu_baseline:
					t0 = time.time()
					r2 = subprocess.run([refcli, "-align", fa, "-output", os.path.join(td, name + ".cpu.afa"), "-threads", str(host_threads())],
					  capture_output=True, text=True)
					rec[name + "_cpu_wall_s"] = time.time() - t0
					rec[name + "_cpu_threads"] = host_threads()
					if r.returncode == 0 and r2.returncode == 0:
						rec[name + "_msa_identical"] = open(os.path.join(td, name + ".gpu.afa")).read() == open(os.path.join(td, name + ".cpu.afa")).read()
		rec["what"] = "wall time of `muscle_b200 -align` (reference CLI + libmuscle_b200) incl. process start and CUDA context; CPU = unmodified reference CLI (parity build)"
		ex["align_cli"] = rec
	return ex


if __name__ == "__main__":
	main()
