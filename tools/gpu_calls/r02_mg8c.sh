#!/bin/bash
MMDIT_COMM_CTAS=32 bash tools/r02_multi.sh 8 cfg2 0 _ctas32
MMDIT_COMM_CTAS=64 bash tools/r02_multi.sh 8 cfg2 0 _ctas64
MMDIT_COMM_UNROLL=2 bash tools/r02_multi.sh 8 cfg2 0 _unroll2
