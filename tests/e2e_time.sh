#!/bin/bash
# wall time of `muscle -align` end to end: GPU-engine build vs the unmodified CPU reference (all host threads)
# usage: tests/e2e_time.sh <config name e.g. C2> [skipcpu]
set -e
cd "$(dirname "$0")/.."
python - "$1" <<'PY'
import sys
sys.path.insert(0, '.')
from muscle_b200 import synth
seqs = synth.make_config(sys.argv[1])
open('/tmp/e2e_%s.fa' % sys.argv[1], 'w').write("".join(">s%d\n%s\n" % (i, s) for i, s in enumerate(seqs)))
PY
t0=$(date +%s.%N); integration/_build/muscle_b200 -align /tmp/e2e_$1.fa -output /tmp/e2e_$1.gpu.afa -quiet; t1=$(date +%s.%N); echo "gpu-engine muscle -align $1: $(python -c "print(round($t1 - $t0, 2))") s wall"
if [ "$2" != "skipcpu" ]; then
t0=$(date +%s.%N); oracle/_ref/muscle -align /tmp/e2e_$1.fa -output /tmp/e2e_$1.cpu.afa -quiet; t1=$(date +%s.%N); echo "cpu reference muscle -align $1 ($(nproc) cores): $(python -c "print(round($t1 - $t0, 2))") s wall"
cmp -s /tmp/e2e_$1.gpu.afa /tmp/e2e_$1.cpu.afa && echo "MSA IDENTICAL" || echo "MSA DIFFERENT"
fi
