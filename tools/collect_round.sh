#!/bin/bash
# One gpurun call that collects a round's evidence into gpurun_out/ (every stage under its own timeout, later stages
# still run if an earlier one fails).  Usage on the GPU box:
#   bash tools/collect_round.sh [tag] [stages]      stages: any of  tests smoke bench b512 wmsa ncu full   (default: all)
tag=${1:-rXX}
stages=${2:-"tests smoke bench b512 wmsa ncu full"}
out=gpurun_out
mkdir -p $out
has() { [[ " $stages " == *" $1 "* ]]; }
if has tests; then      # the WHOLE gpu suite, no -x
  timeout 500 python -m pytest tests -m gpu -v -s -p no:cacheprovider > $out/${tag}_gpu_tests.log 2>&1
  echo "pytest rc=$?" >> $out/${tag}_gpu_tests.log
  grep -E "passed|failed|error" $out/${tag}_gpu_tests.log | tail -3
fi
if has smoke; then      # what the driver runs before the bench
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $out/${tag}_smoke.log
fi
if has bench; then
  timeout 400 python bench.py > $out/${tag}_bench_fwd.json 2> $out/${tag}_bench_fwd.err; echo "bench fwd rc=$?"
  head -c 700 $out/${tag}_bench_fwd.json; echo
  timeout 200 python bench.py --residual bf16 --no-cpu-baseline --steps 10 > $out/${tag}_bench_fwd_bf16resid.json 2> /dev/null; echo "bench bf16-resid rc=$?"
fi
if has b512; then
  timeout 200 python bench.py --size 512 --no-cpu-baseline > $out/${tag}_bench_512.json 2> $out/${tag}_bench_512.err; echo "bench 512 rc=$?"
  head -c 300 $out/${tag}_bench_512.json; echo
fi
if has wmsa; then
  timeout 300 python tools/wmsa_microbench.py > $out/${tag}_wmsa_microbench.json 2> $out/${tag}_wmsa_microbench.err; echo "wmsa rc=$?"
fi
if has ncu; then        # launch list + DRAM bytes + tensor-pipe % of every kernel of one forward
  LABELS_OUT=$out/${tag}_labels.json timeout 600 ncu --profile-from-start off --clock-control none --csv \
    --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct \
    --log-file $out/${tag}_fwd_metrics.csv python tools/forward_once.py > $out/${tag}_ncu.log 2>&1; echo "ncu rc=$?"
fi
if has full; then       # one --set full capture with source correlation of the two dominant kernels
  timeout 300 ncu --set full --import-source on --clock-control none -k regex:leff_fused -s 2 -c 1 -o $out/${tag}_leff_fused_c256 \
    python tools/leff_fused_probe.py 256 64 32 > $out/${tag}_ncu_full1.log 2>&1; echo "ncu full leff rc=$?"
  timeout 300 ncu --set full --import-source on --clock-control none -k regex:wmsa --profile-from-start off -c 12 -o $out/${tag}_wmsa_all \
    python tools/forward_once.py > $out/${tag}_ncu_full2.log 2>&1; echo "ncu full wmsa rc=$?"
fi
ls -la $out | tail -14
