#!/usr/bin/env bash
# Everything that has to be green on a B200 before a kernel change is kept (run from the repo root, e.g. through
#   gpurun --timeout 1500 -- 'bash tools/run_gpu_checks.sh'        # one GPU
#   gpurun --gpus 2 --timeout 900 -- 'bash tools/run_gpu_checks.sh multi'
# ).  Each step is wrapped in its own timeout so a hung kernel cannot take the box down.
set -u
fail=0
step() { echo "=== $*"; "$@" || { echo "!!! FAILED: $*"; fail=1; }; }
if [ "${1:-}" = "multi" ]; then
    n=$(nvidia-smi -L | wc -l)
    step timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$n" --master-addr 127.0.0.1 \
        --master-port 29561 tools/mgpu_check.py
    step timeout 600 python -m pytest tests/test_gpu_multi.py -x -q
else
    step timeout 300 tools/bin/gemm_test 0           # GEMM / conv3d kernel variants vs a naive reference
    step timeout 300 tools/bin/attn_test 0           # the three attention kernels vs a naive reference
    step timeout 900 python -m pytest tests -m gpu -x -q
    step timeout 120 python -c "import __graft_entry__ as g; g.smoke()"
    step timeout 400 python bench.py --steps 4 --warmup 3 --no-cpu-baseline
fi
exit $fail
