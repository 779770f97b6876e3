#!/bin/bash
# Rebuild the library if any source changed, then run a command on a B200 through gpurun.
#   tools/gpu.sh [--timeout S] -- '<command>'
set -e
cd "$(dirname "$0")/.."
python -m ldmae_b200.build > /dev/null
exec /usr/local/graft/bin/gpurun "$@"
