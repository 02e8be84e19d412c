#!/bin/bash
# builds the extension locally (so the snapshot carries a fresh .so), then runs the given command on a B200 box
# usage: tools/gpu.sh <timeout_s> '<command>'   (output -> /tmp/gpurun_last.log)
set -e
cd /root/repo
python -c "import sys; sys.path.insert(0,'acl-gan_b200'); import aclgan_native as N; N.build(); L=N.lib(); assert not [s for s in N.exported_symbols() if not hasattr(L,s)]"
exec /usr/local/graft/bin/gpurun --timeout "$1" -- "$2"
