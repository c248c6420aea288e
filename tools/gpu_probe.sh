#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 300 tools/bin/gl_probe > gpurun_out/probe.log 2>&1
echo "probe exit $?"
cat gpurun_out/probe.log
