#!/bin/bash
timeout 900 python tools/diag_train_grad.py 2>&1 | grep "|"
