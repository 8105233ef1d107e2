#!/bin/bash
mkdir -p gpurun_out
echo "=== mma kernel phases"; timeout 120 python scripts/attn_phases.py 256 8 32 64 2>&1 | grep phases
echo "=== scalar kernel phases"; HQ_DEBUG=1 HQ_ATTN_SCALAR=1 timeout 120 python scripts/attn_phases.py 256 8 32 64 2>&1 | grep phases
echo "=== mma stages=2 (8 CTAs/SM)"; HQ_DEBUG=1 HQ_ATTM_STAGES=2 timeout 120 python scripts/attn_phases.py 256 32 64 2>&1 | grep phases
echo "=== mma stages=3"; HQ_DEBUG=1 HQ_ATTM_STAGES=3 timeout 120 python scripts/attn_phases.py 256 32 2>&1 | grep phases
