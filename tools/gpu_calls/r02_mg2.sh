#!/bin/bash
bash tools/r02_multi.sh 2 cfg2
