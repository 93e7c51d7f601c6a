#!/bin/bash
# build the library in-tree (a stale .so would travel to the box), then run a command on the B200 box
# usage: tools/gpu.sh [--gpus N] <timeout seconds> '<command>'
set -e
cd "$(dirname "$0")/.."
GP=""
if [ "$1" == "--gpus" ]; then GP="--gpus $2"; shift 2; fi
make -C tf-1d-2d-segmentation-end2endpipelines_b200 -j8 > /dev/null
python -c "import __graft_entry__ as g; g.build()"
exec /usr/local/graft/bin/gpurun $GP --timeout "$1" -- "$2"
