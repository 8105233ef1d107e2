#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "$TMO" "$@" > gpurun_out/$name.log 2>&1; echo "exit=$?" >> gpurun_out/$name.log; tail -n ${TAILN:-3} gpurun_out/$name.log | cut -c1-400; }
TMO=1500 TAILN=12 run r2h_tests python -m pytest tests -m gpu -q -p no:cacheprovider -rs
TMO=300 TAILN=2 run r2h_smoke python __graft_entry__.py smoke
