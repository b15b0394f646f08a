#!/bin/bash
# One parametrised GPU-box session: tools/gpu_session.sh <tag> <step> [<step> ...]; every step's output lands in gpurun_out/<tag>_*.
# steps: tests | tests:<pytest -k expr> | smoke | bench[:extra args] | benchn:<N>[:extra args] | ref | py:<script and args>
tag=$1; shift
mkdir -p gpurun_out
for step in "$@"; do
  kind=${step%%:*}; arg=""; [[ "$step" == *:* ]] && arg=${step#*:}
  case $kind in
    tests) if [ -n "$arg" ]; then timeout 1500 python -m pytest tests -m gpu -q -x -k "$arg" > gpurun_out/${tag}_pytest.log 2>&1; else timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest.log 2>&1; fi; tail -5 gpurun_out/${tag}_pytest.log ;;
    smoke) timeout 600 python __graft_entry__.py smoke > gpurun_out/${tag}_smoke.log 2>&1; tail -3 gpurun_out/${tag}_smoke.log ;;
    bench) timeout 1500 python bench.py $arg > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; tail -c 1500 gpurun_out/${tag}_bench.json; tail -5 gpurun_out/${tag}_bench.err; cp gpurun_out/bench_full_n1.json gpurun_out/${tag}_bench_full.json 2>/dev/null ;;
    benchn) n=${arg%%:*}; extra=""; [[ "$arg" == *:* ]] && extra=${arg#*:}
       timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n $extra > gpurun_out/${tag}_bench_n$n.json 2> gpurun_out/${tag}_bench_n$n.err
       tail -c 1500 gpurun_out/${tag}_bench_n$n.json; tail -5 gpurun_out/${tag}_bench_n$n.err; cp gpurun_out/bench_full_n$n.json gpurun_out/${tag}_bench_full_n$n.json 2>/dev/null ;;
    ref) timeout 900 python bench.py --impl reference $arg > gpurun_out/${tag}_bench_ref.json 2> gpurun_out/${tag}_bench_ref.err; tail -c 600 gpurun_out/${tag}_bench_ref.json ;;
    py) timeout 1500 python $arg > gpurun_out/${tag}_py.log 2>&1; tail -30 gpurun_out/${tag}_py.log ;;
  esac
  echo "== $step done rc=$?"
done
