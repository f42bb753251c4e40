#!/bin/bash
# build here (no GPU needed), run under gpurun:  gpurun -- scripts/ubench/ubench
set -e
cd "$(dirname "$0")"
nvcc -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -o ubench ubench.cu
