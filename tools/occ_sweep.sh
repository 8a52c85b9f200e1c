#!/bin/bash
for k in 2 3 4 5 6 7 8; do echo "grid per SM $k"; PB2_GRID_PER_SM=$k python tools/tune_trace.py 12 8 0; done
