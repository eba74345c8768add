#!/bin/bash
bash tools/r02_multi.sh 8 cfg2 0 _b
bash tools/r02_multi.sh 8 cfg3 64 _b
bash tools/r02_multi.sh 1 cfg2 0 _b8
