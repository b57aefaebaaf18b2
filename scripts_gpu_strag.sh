#!/bin/bash
export ANM_B200_LIB=$PWD/gym_anm_b200/lib/libanm_b200_diag.so
python tools/tail_probe.py 1.0
ANM_FORCE_DENSE=1 python tools/tail_probe.py 1.0
