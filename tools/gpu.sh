#!/bin/bash
# Dev-container side: refresh every in-tree binary, then run a command on the GPU box.
# usage: tools/gpu.sh [--gpus N] <timeout-seconds> '<command>'
cd "$(dirname "$0")/.."
GP=()
if [ "$1" = "--gpus" ]; then GP=(--gpus "$2"); shift 2; fi
python -c "import __graft_entry__ as g; g.build()" || exit 1
exec /usr/local/graft/bin/gpurun "${GP[@]}" --timeout "$1" -- "$2"
