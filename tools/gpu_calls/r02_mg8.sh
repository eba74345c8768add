#!/bin/bash
bash tools/r02_multi.sh 8 cfg2
bash tools/r02_multi.sh 8 cfg3 64
bash tools/r02_multi.sh 8 cfg4 16
