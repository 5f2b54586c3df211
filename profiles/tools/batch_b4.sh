#!/bin/bash
bash profiles/tools/run_variants_wl.sh robertson_4M --steps 3 --warmup 3 -- default rob256 rob512
bash profiles/tools/run_variants_wl.sh pendulum3_1M_tend2 --steps 3 --warmup 3 -- default b128x2
