#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/diag_train_grad.py 2>&1 | grep native
timeout 600 python tools/prof_pretrain.py 2>&1 | tail -60
