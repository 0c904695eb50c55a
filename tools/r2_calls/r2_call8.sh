#!/bin/bash
cd ${GRAFT_REPO_ROOT:-.}
timeout 200 python tools/dcb_tail_trace.py 136 240 384 384 384 2>&1 | tail -52
