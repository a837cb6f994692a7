#!/bin/bash
# what the set-up + epilogue cost: march time as a function of the number of time steps
for t in 2 1024; do python tools/variant_probe.py 1024 $t 32768 237 1237 | sed 's/regs.*kernel//; s/maxdiff.*//'; done
for t in 2 512; do python tools/variant_probe.py 512 $t 32768 138 | sed 's/regs.*kernel//; s/maxdiff.*//'; done
