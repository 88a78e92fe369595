#!/bin/bash
# Runs the book's unmodified ch4/Main.cpp on the engine (bin/main_ch4) for a few seeds and writes the observables the
# statistics test compares (tests/golden/make_ch4_statistics.py::summarise) to gpurun_out/ch4/.
ROOT=$(cd "$(dirname "$0")/.." && pwd)
OUT=$ROOT/gpurun_out/ch4
mkdir -p $OUT
for s in "$@"; do
    d=$(mktemp -d)
    mkdir -p $d/results
    t0=$SECONDS
    ( cd $d && env ESPIC_SEED=$s timeout 200 $ROOT/plasma-simulations-by-example_b200/bin/main_ch4 > run.log 2>&1 )
    echo "seed $s: $((SECONDS - t0)) s"; tail -2 $d/run.log
    cp $d/runtime_diags.csv $OUT/runtime_diags_$s.csv
    python - "$d" "$OUT/gpu_$s.json" "$ROOT" <<'PY'
import json, sys
sys.path.insert(0, sys.argv[3] + "/tests/golden")
from make_ch4_statistics import summarise
json.dump(summarise(sys.argv[1]), open(sys.argv[2], "w"), indent=1)
print("wrote", sys.argv[2])
PY
done
