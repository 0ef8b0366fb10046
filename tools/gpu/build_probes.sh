#!/bin/sh
# builds the standalone probe binaries next to their sources (they travel to the GPU box with the snapshot)
set -e
cd "$(dirname "$0")"
mkdir -p _bin
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xptxas -v imad_probes.cu -o _bin/imad_probes 2> _bin/imad_probes.ptxas.log
echo built _bin/imad_probes
