#!/bin/bash
# One gpurun call that collects a round's evidence into gpurun_out/ (every stage under its own timeout, later stages
# still run if an earlier one fails).  Usage on the GPU box:
#   bash tools/collect_round.sh [tag] [stages]      stages: any of  tests bench train probe wmsa ncu   (default: all)
# e.g.  gpurun --timeout 900 -- 'bash tools/collect_round.sh r02a "tests bench train"'
tag=${1:-rXX}
stages=${2:-"tests bench train probe wmsa ncu"}
out=gpurun_out
mkdir -p $out
has() { [[ " $stages " == *" $1 "* ]]; }
if has tests; then
  timeout 400 python -m pytest tests -m gpu -v -s -p no:cacheprovider > $out/${tag}_gpu_tests.log 2>&1
  echo "pytest rc=$?" >> $out/${tag}_gpu_tests.log
  grep -E "passed|failed|error" $out/${tag}_gpu_tests.log | tail -3
fi
if has bench; then
  timeout 300 python bench.py > $out/${tag}_bench_fwd.json 2> $out/${tag}_bench_fwd.err; echo "bench fwd rc=$?"
  head -c 600 $out/${tag}_bench_fwd.json; echo
fi
if has train; then
  timeout 200 python bench.py --mode train --steps 5 --warmup 3 > $out/${tag}_bench_train.json 2> $out/${tag}_bench_train.err; echo "bench train rc=$?"
  head -c 400 $out/${tag}_bench_train.json; echo
fi
if has probe; then
  timeout 150 python tools/train_probe.py 8 > $out/${tag}_train_probe.json 2> $out/${tag}_train_probe.err; echo "probe rc=$?"
  head -c 300 $out/${tag}_train_probe.json; echo
fi
if has wmsa; then
  timeout 200 python tools/wmsa_microbench.py > $out/${tag}_wmsa_microbench.json 2> $out/${tag}_wmsa_microbench.err; echo "wmsa rc=$?"
fi
if has ncu; then
  timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file $out/${tag}_launches.csv python tools/forward_once.py > $out/${tag}_ncu.log 2>&1; echo "ncu rc=$?"
fi
ls -la $out | tail -12
