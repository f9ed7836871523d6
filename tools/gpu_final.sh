#!/bin/bash
# one gpurun call: the final evidence set of a round (GPU test suite, smoke, every bench line, reference arms, micro-benchmarks)
tag=${1:-final}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -W ignore 2>&1 | tail -5 | tee gpurun_out/${tag}_pytest.txt
python __graft_entry__.py --smoke 2>&1 | tail -1 | tee gpurun_out/${tag}_smoke.txt
one() { name=$1; shift; timeout 200 python bench.py "$@" 2>gpurun_out/${tag}_${name}.err | tail -1 > gpurun_out/${tag}_${name}.json; }
one D20 --steps 20 --warmup 3
one D2000 --steps 2000
one D2000_wide --steps 2000 --e2e-wide
for c in P B C E; do one $c --config $c --steps 2000; done
one E20 --config E --steps 20 --warmup 3
one ref_D --impl reference --steps 8 --warmup 2
one ref_E --impl reference --config E --steps 8 --warmup 2
one refcuda_D --impl reference-cuda --steps 50
one refcuda_E --impl reference-cuda --config E --steps 50
timeout 400 python tools/bench_ops.py > gpurun_out/${tag}_ops_microbench.jsonl 2> gpurun_out/${tag}_ops_microbench.err
ls gpurun_out/${tag}_* | head -40
