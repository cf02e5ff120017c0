#!/usr/bin/env bash
# helper: run on the GPU box via gpurun
set -x
mkdir -p gpurun_out
nvidia-smi -L
python -m pytest tests -m gpu -x -q 2>&1 | tail -40
