#!/bin/bash
echo "== torch-bundled cuBLAS (12.8)"; python profiles/time_model_pass.py 2>&1 | tail -3
PRE=/usr/local/cuda/lib64/libcublasLt.so.12:/usr/local/cuda/lib64/libcublas.so.12
echo "== system cuBLAS 12.9 preloaded, no emulation"; LD_PRELOAD=$PRE python profiles/time_model_pass.py 2>&1 | tail -3
echo "== system cuBLAS 12.9 preloaded, CUBLAS_EMULATE_SINGLE_PRECISION=1"; LD_PRELOAD=$PRE CUBLAS_EMULATE_SINGLE_PRECISION=1 python profiles/time_model_pass.py 2>&1 | tail -3
echo "== ... + CUBLAS_EMULATION_STRATEGY=eager"; LD_PRELOAD=$PRE CUBLAS_EMULATE_SINGLE_PRECISION=1 CUBLAS_EMULATION_STRATEGY=eager python profiles/time_model_pass.py 2>&1 | tail -3
