#!/bin/bash
cd "$GRAFT_REPO_ROOT"
python examples/stage_flow.py 12 32 2>&1 | tail -16
python examples/stage_flow.py 20 128 2>&1 | tail -16
