#!/bin/bash
MMDIT_COMM_UNROLL=4 bash tools/r02_multi.sh 8 cfg2 0 _unroll4
MMDIT_COMM_UNROLL=2 MMDIT_COMM_CTAS=64 bash tools/r02_multi.sh 8 cfg2 0 _unroll2_ctas64
