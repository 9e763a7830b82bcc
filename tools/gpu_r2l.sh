#!/bin/bash
timeout 600 python tools/diag_train_ops.py 2>&1 | tail -16
for i in 1 2 3; do timeout 600 python tools/diag_train_grad.py 2>&1 | grep "sap native_linear=True scale=1024"; done
