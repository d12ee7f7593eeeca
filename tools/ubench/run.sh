#!/bin/bash
cd "$(dirname "$0")" && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -o /tmp/ubench ubench.cu && /tmp/ubench
