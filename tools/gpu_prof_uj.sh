#!/bin/bash
# wait accounting of the uniform-jobs stage kernel (who waits for whom), then the plain timing
mkdir -p gpurun_out
MW_TILE_VARIANT=3 MW_STAGE_PROF=1 timeout 300 python tools/probe_dycore.py 2>&1 | tail -8
