#!/bin/bash
mkdir -p gpurun_out
timeout 100 python tools/diag_dropin.py default 2>&1 | grep -v "^out_rgb\|Warning\|torch.cross\|linalg.cross\|left_normal\|default value" | tail -12
GSB200_LIB=$PWD/gsgen_b200/_variants/lib_noellipse.so timeout 100 python tools/diag_dropin.py noellipse 2>&1 | grep -v "^out_rgb\|Warning\|torch.cross\|linalg.cross\|left_normal\|default value" | tail -12
